// Streaming inference with cached per-vertex intermediates: the device-side pieces of the reference's
// inference_optimized.py:144-301.  A request is a handful of new edges; the handler touches only the vertices around them
// (row sets of tens, not thousands), so these kernels are sized for latency: fp32 throughout (the reference serves in fp32 on the
// CPU), no tensor-core path, no workspace.
//
//   ogl_graph_row_degrees / ogl_graph_gather_rows   g.out_degrees(v) / g.in_edges(v) / g.out_edges(v)   (:185, :194-196, :205)
//   ogl_infer_rows_linear                           relu(fc_pool(h)) and fc_self(h) + fc_neigh(neigh) on a row set (:258-260, :273-276)
//   ogl_infer_induced_mean                          subgraph(S).update_all(copy_src, mean) restricted to edges inside S (:265-268)
#include "graph.cuh"

namespace ogl {

namespace {

__global__ void k_row_degrees(GraphView g, const int64_t* __restrict__ v, int64_t n, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t x = v[i];
  out[i] = (x >= 0 && x < g.n_vertices) ? g.deg[x] : 0;
}

// one warp per requested vertex: its adjacency row (ascending edge id) -> out[offsets[i] ...]
__global__ void k_gather_adj_rows(GraphView g, const int64_t* __restrict__ v, int64_t n, const int64_t* __restrict__ offsets,
                                  int64_t* __restrict__ out_src) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const int64_t x = v[w];
  if (x < 0 || x >= g.n_vertices) return;
  const int64_t base = g.row_start[x];
  const int d = g.deg[x];
  int64_t* dst = out_src + offsets[w];
  for (int j = lane; j < d; j += 32) dst[j] = (int64_t)(g.adj[base + j] & 0xFFFFFFFFull);
}

// out[out_ids[i], o] = act( x1[ids1[i], :k1] . w1[o, :k1] + b1[o] + (x2 ? x2[ids2[i], :k2] . w2[o, :k2] + b2[o] : 0) )
// one CTA per row (the row(s) staged in shared memory), one warp per output column (coalesced weight rows, shuffle reduction)
__global__ void __launch_bounds__(256) k_rows_linear(const float* __restrict__ x1, int ld1, const int64_t* __restrict__ ids1,
                                                     const float* __restrict__ w1, int k1, const float* __restrict__ b1,
                                                     const float* __restrict__ x2, int ld2, const int64_t* __restrict__ ids2,
                                                     const float* __restrict__ w2, int k2, const float* __restrict__ b2, int relu,
                                                     float* __restrict__ out, int ldo, const int64_t* __restrict__ out_ids, int n_out) {
  extern __shared__ float row[];                    // [k1 + k2]
  const int i = blockIdx.x;
  const float* r1 = x1 + (int64_t)ids1[i] * ld1;
  for (int c = threadIdx.x; c < k1; c += blockDim.x) row[c] = r1[c];
  if (x2) {
    const float* r2 = x2 + (int64_t)ids2[i] * ld2;
    for (int c = threadIdx.x; c < k2; c += blockDim.x) row[k1 + c] = r2[c];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  float* orow = out + (int64_t)out_ids[i] * ldo;
  for (int o = warp; o < n_out; o += n_warps) {
    float acc = 0.f;
    const float* wr = w1 + (int64_t)o * k1;
    for (int c = lane; c < k1; c += 32) acc = fmaf(row[c], wr[c], acc);
    if (x2) {
      const float* wr2 = w2 + (int64_t)o * k2;
      for (int c = lane; c < k2; c += 32) acc = fmaf(row[k1 + c], wr2[c], acc);
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, sft);
    if (lane == 0) {
      float y = acc + (b1 ? b1[o] : 0.f) + ((x2 && b2) ? b2[o] : 0.f);
      if (relu) y = fmaxf(y, 0.f);
      orow[o] = y;
    }
  }
}

// neigh[v, :] = mean over in-edges (u -> v) with member[u] != 0 of proj[u, :], 0 when there is none; one warp per vertex of the
// set, edges in ascending edge id (the message order of DGL's mailbox), lanes over 16-byte column strips
__global__ void __launch_bounds__(256) k_induced_mean(GraphView g, const uint8_t* __restrict__ member, const int64_t* __restrict__ nodes,
                                                      int64_t n, const float* __restrict__ proj, int ldp, int F, float* __restrict__ out,
                                                      int ldo) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const int64_t v = nodes[w];
  const int64_t base = g.row_start[v];
  const int d = g.deg[v];
  for (int c0 = lane; c0 < F; c0 += 32) {
    float acc = 0.f;
    int cnt = 0;
    for (int j = 0; j < d; ++j) {
      const int64_t u = (int64_t)(g.adj[base + j] & 0xFFFFFFFFull);
      if (member[u]) { acc += proj[u * ldp + c0]; ++cnt; }
    }
    out[v * ldo + c0] = cnt ? acc / (float)cnt : 0.f;
  }
}

// The whole neighbourhood query of one request in ONE launch and one D2H copy.  Row x of the view holds the sources of x's in-edges,
// row x + v_off the targets of x's out-edges (the serving graph stores every edge in both).  For the request's vertices v[0 .. n):
//   out[0] = 1 if a section overflowed (nothing is copied then), out[1] / out[2] = total in / out entries,
//   out[3 + i] = out-degree of v[i];  in-rows and out-rows are copied only for vertices with out-degree < th (inference_optimized.py:186-192):
//   in offsets [n + 1] | out offsets [n + 1] | in sources [cap_in] | out targets [cap_out] | out-degree of every target [cap_out]
__global__ void __launch_bounds__(256) k_infer_query(GraphView g, const int64_t* __restrict__ v, int n, int64_t v_off, int th, int cap_in,
                                                     int cap_out, int64_t* __restrict__ out) {
  int64_t* deg_out = out + 3;
  int64_t* off_in = out + 3 + n;
  int64_t* off_out = off_in + n + 1;
  int64_t* in_src = off_out + n + 1;
  int64_t* out_dst = in_src + cap_in;
  int64_t* out_deg_dst = out_dst + cap_out;
  __shared__ int overflow;
  if (threadIdx.x == 0) {
    int64_t ti = 0, to = 0;
    for (int i = 0; i < n; ++i) {
      const int64_t x = v[i];
      const int dgo = g.deg[x + v_off];
      deg_out[i] = dgo;
      off_in[i] = ti;
      off_out[i] = to;
      if (dgo < th) { ti += g.deg[x]; to += dgo; }
    }
    off_in[n] = ti;
    off_out[n] = to;
    out[1] = ti;
    out[2] = to;
    overflow = (ti > cap_in || to > cap_out) ? 1 : 0;
    out[0] = overflow;
  }
  __syncthreads();
  if (overflow) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < n; i += 8) {
    const int64_t x = v[i];
    const int n_in = (int)(off_in[i + 1] - off_in[i]), n_out = (int)(off_out[i + 1] - off_out[i]);
    const int64_t bi = g.row_start[x], bo = g.row_start[x + v_off];
    for (int j = lane; j < n_in; j += 32) in_src[off_in[i] + j] = (int64_t)(g.adj[bi + j] & 0xFFFFFFFFull);
    for (int j = lane; j < n_out; j += 32) {
      const int64_t t = (int64_t)(g.adj[bo + j] & 0xFFFFFFFFull);
      out_dst[off_out[i] + j] = t;
      out_deg_dst[off_out[i] + j] = g.deg[t + v_off];
    }
  }
}

}  // namespace

}  // namespace ogl

using namespace ogl;

extern "C" int ogl_infer_query(ogl_graph* g, const int64_t* v_dev, int n, int64_t v_off, int th, int cap_in, int cap_out, int64_t* out_dev,
                               void* stream) {
  OGL_ARG(g && v_dev && out_dev && n > 0 && v_off >= 0 && cap_in > 0 && cap_out > 0, "ogl_infer_query: bad arguments");
  OGL_LAUNCH(k_infer_query, 1, 256, 0, stream, graph_view(g), v_dev, n, v_off, th, cap_in, cap_out, out_dev);
  return OGL_OK;
}

extern "C" int ogl_graph_row_degrees(ogl_graph* g, const int64_t* v_dev, int64_t n, int64_t* deg_out_dev, void* stream) {
  OGL_ARG(g && (n == 0 || (v_dev && deg_out_dev)), "ogl_graph_row_degrees: null");
  if (n == 0) return OGL_OK;
  OGL_LAUNCH(k_row_degrees, (unsigned)ceil_div(n, 256), 256, 0, stream, graph_view(g), v_dev, n, deg_out_dev);
  return OGL_OK;
}

extern "C" int ogl_graph_gather_rows(ogl_graph* g, const int64_t* v_dev, int64_t n, const int64_t* offsets_dev, int64_t* out_src_dev,
                                     void* stream) {
  OGL_ARG(g && (n == 0 || (v_dev && offsets_dev && out_src_dev)), "ogl_graph_gather_rows: null");
  if (n == 0) return OGL_OK;
  OGL_LAUNCH(k_gather_adj_rows, (unsigned)ceil_div(n * 32, 256), 256, 0, stream, graph_view(g), v_dev, n, offsets_dev, out_src_dev);
  return OGL_OK;
}

extern "C" int ogl_infer_rows_linear(const float* x1_dev, int ld1, const int64_t* ids1_dev, const float* w1_dev, int k1, const float* b1_dev,
                                     const float* x2_dev, int ld2, const int64_t* ids2_dev, const float* w2_dev, int k2,
                                     const float* b2_dev, int relu, float* out_dev, int ldo, const int64_t* out_ids_dev, int64_t n_rows,
                                     int n_out, void* stream) {
  OGL_TRY(require_device());
  OGL_ARG(x1_dev && ids1_dev && w1_dev && out_dev && out_ids_dev && k1 > 0 && n_out > 0 && n_rows >= 0, "ogl_infer_rows_linear: bad arguments");
  OGL_ARG(!x2_dev || (ids2_dev && w2_dev && k2 > 0), "ogl_infer_rows_linear: incomplete second segment");
  if (n_rows == 0) return OGL_OK;
  const size_t smem = sizeof(float) * (size_t)(k1 + (x2_dev ? k2 : 0));
  OGL_ARG(smem <= 48 * 1024, "ogl_infer_rows_linear: rows of %d + %d floats exceed the 48 KB staging buffer", k1, k2);
  OGL_LAUNCH(k_rows_linear, (unsigned)n_rows, 256, smem, stream, x1_dev, ld1, ids1_dev, w1_dev, k1, b1_dev, x2_dev, ld2, ids2_dev, w2_dev,
             x2_dev ? k2 : 0, b2_dev, relu, out_dev, ldo, out_ids_dev, n_out);
  return OGL_OK;
}

extern "C" int ogl_infer_induced_mean(ogl_graph* g, const uint8_t* member_dev, const int64_t* nodes_dev, int64_t n, const float* proj_dev,
                                      int ldp, int n_feats, float* out_dev, int ldo, void* stream) {
  OGL_ARG(g && (n == 0 || (member_dev && nodes_dev && proj_dev && out_dev)) && n_feats > 0, "ogl_infer_induced_mean: bad arguments");
  if (n == 0) return OGL_OK;
  OGL_LAUNCH(k_induced_mean, (unsigned)ceil_div(n * 32, 256), 256, 0, stream, graph_view(g), member_dev, nodes_dev, n, proj_dev, ldp,
             n_feats, out_dev, ldo);
  return OGL_OK;
}
