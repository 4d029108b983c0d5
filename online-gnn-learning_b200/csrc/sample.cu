// Uniform k-neighbour sampler (Philox) + to_block frontier compaction + RBR subset draw.
//
// Replaces dgl.sampling.MultiLayerNeighborSampler([s]*2, replace=True, return_eids=True) and
// NodeDataLoader's to_block as driven from train/graphsage/pytorch/model.py:44-47,128-131.
// Device twin of oracle/sampler.py (bit-exact index lists under the same Philox stream).
#include "sample.cuh"

namespace ogl {

constexpr int kBlock = 256;

// one thread per (row, quad of picks): 1 Philox call -> 4 picks
__global__ void __launch_bounds__(kBlock) k_sample(GraphView g, const int32_t* __restrict__ dst_nodes, const int32_t* __restrict__ n_dst_dev,
                                                   int n_dst_max, int fanout, uint2 key, const uint32_t* __restrict__ step_dev,
                                                   uint32_t step_imm, uint32_t hop, int32_t* __restrict__ out_src,
                                                   int64_t* __restrict__ out_eid) {
  const int Q = (fanout + 3) >> 2;
  const int n_dst = n_dst_dev ? min(*n_dst_dev, n_dst_max) : n_dst_max;
  const uint32_t step = step_dev ? *step_dev : step_imm;
  const int64_t total = (int64_t)n_dst * Q;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(t / Q), q = (int)(t % Q);
    const int v = dst_nodes[row];
    const int deg = g.deg[v];
    const int64_t start = g.row_start[v];
    const uint4 w = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)row, hop, step), key);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = q * 4 + i;
      if (j < fanout) {
        const int64_t o = (int64_t)row * fanout + j;
        if (deg > 0) {
          const uint32_t r = __umulhi(pick4(w, i), (uint32_t)deg);
          const unsigned long long ent = __ldg(g.adj + start + r);      // one 8-byte entry: source + edge id in one sector
          out_src[o] = (int32_t)(uint32_t)ent;
          if (out_eid) out_eid[o] = (int64_t)(ent >> 32);
        } else {
          out_src[o] = -1;
          if (out_eid) out_eid[o] = -1;
        }
      }
    }
  }
}

// seed ids -> int32 rows.  An id outside [0, n_vertices) would index deg / row_start out of bounds: it is replaced by vertex 0
// and reported through *err_flag (read back by ogl_plan_error_flags; DGL raises on such seeds)
__global__ void __launch_bounds__(kBlock) k_cast_nodes(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t n, int64_t n_vertices,
                                                       uint32_t* __restrict__ err_flag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t v = in[i];
    if (v < 0 || v >= n_vertices) {
      if (err_flag) atomicOr(err_flag, 1u);
      v = 0;
    }
    out[i] = (int32_t)v;
  }
}

// ---- to_block: direct-address first-appearance table ----------------------------------------
// first[g] = min over appearances of (position in dst list | n_dst + edge position)
__global__ void __launch_bounds__(kBlock) k_tb_mark(const int32_t* __restrict__ dst_nodes, const int32_t* __restrict__ n_dst_dev,
                                                    int n_dst_max, int fanout, const int32_t* __restrict__ picked, int32_t* __restrict__ first) {
  const int n_dst = min(*n_dst_dev, n_dst_max);
  const int64_t ne = (int64_t)n_dst * fanout;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_dst + ne; t += (int64_t)gridDim.x * blockDim.x) {
    if (t < n_dst) {
      atomicMin(&first[dst_nodes[t]], (int32_t)t);
    } else {
      const int g = picked[t - n_dst];
      if (g >= 0) atomicMin(&first[g], (int32_t)t);
    }
  }
}

__global__ void __launch_bounds__(kBlock) k_tb_flags(const int32_t* __restrict__ n_dst_dev, int n_dst_max, int fanout,
                                                     const int32_t* __restrict__ picked, const int32_t* __restrict__ first,
                                                     int32_t* __restrict__ flags, int64_t ne_max) {
  const int n_dst = min(*n_dst_dev, n_dst_max);
  const int64_t ne = (int64_t)n_dst * fanout;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < ne_max; p += (int64_t)gridDim.x * blockDim.x) {
    int f = 0;
    if (p < ne) {
      const int g = picked[p];
      f = (g >= 0 && first[g] == (int32_t)(n_dst + p)) ? 1 : 0;
    }
    flags[p] = f;
  }
}

__global__ void __launch_bounds__(kBlock) k_tb_emit(const int32_t* __restrict__ dst_nodes, const int32_t* __restrict__ n_dst_dev,
                                                    int n_dst_max, int fanout, const int32_t* __restrict__ picked,
                                                    const int32_t* __restrict__ first, const int32_t* __restrict__ flags,
                                                    const int32_t* __restrict__ pos, const int32_t* __restrict__ n_new_dev,
                                                    int32_t* __restrict__ src_nodes, int32_t* __restrict__ n_src_dev, int n_src_max,
                                                    int32_t* __restrict__ edge_lid, int64_t ne_max) {
  const int n_dst = min(*n_dst_dev, n_dst_max);
  const int64_t ne = (int64_t)n_dst * fanout;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_src_dev = min(n_dst + *n_new_dev, n_src_max);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_dst_max + ne_max; t += (int64_t)gridDim.x * blockDim.x) {
    if (t < n_dst_max) {
      if (t < n_dst) src_nodes[t] = dst_nodes[t];
    } else {
      const int64_t p = t - n_dst_max;
      if (p < ne) {
        const int g = picked[p];
        int lid = -1;
        if (g >= 0) {
          const int m = first[g];
          lid = m < n_dst ? m : n_dst + pos[m - n_dst];
          if (flags[p] && n_dst + pos[p] < n_src_max) src_nodes[n_dst + pos[p]] = g;      // (the plan sizes n_src_max so that this always holds)
        }
        edge_lid[p] = lid;
      } else {
        edge_lid[p] = -1;
      }
    }
  }
}

__global__ void __launch_bounds__(kBlock) k_tb_reset(const int32_t* __restrict__ src_nodes, const int32_t* __restrict__ n_src_dev,
                                                     int n_src_max, int32_t* __restrict__ first) {
  const int n = min(*n_src_dev, n_src_max);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) first[src_nodes[t]] = 0x7fffffff;
}

__global__ void __launch_bounds__(kBlock) k_fill_i32(int32_t* p, int32_t v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- reverse edge lists of a sampled block: for every source row the slots (d * fanout + j) that picked it -------
__global__ void __launch_bounds__(kBlock) k_rev_count(const int32_t* __restrict__ edge_lid, const int32_t* __restrict__ n_dst_dev, int n_dst_max,
                                                      int fanout, int32_t* __restrict__ cnt) {
  const int64_t ne = (int64_t)min(*n_dst_dev, n_dst_max) * fanout;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < ne; p += (int64_t)gridDim.x * blockDim.x) {
    const int lid = edge_lid[p];
    if (lid >= 0) atomicAdd(&cnt[lid], 1);
  }
}
__global__ void __launch_bounds__(kBlock) k_rev_fill(const int32_t* __restrict__ edge_lid, const int32_t* __restrict__ n_dst_dev, int n_dst_max,
                                                     int fanout, const int32_t* __restrict__ rev_ptr, int32_t* __restrict__ cursor,
                                                     int32_t* __restrict__ rev_edge) {
  const int64_t ne = (int64_t)min(*n_dst_dev, n_dst_max) * fanout;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < ne; p += (int64_t)gridDim.x * blockDim.x) {
    const int lid = edge_lid[p];
    // entry = (destination row << 8) | slot: the consumer (k_pool_bwd) then needs no division per column strip
    if (lid >= 0) rev_edge[rev_ptr[lid] + atomicAdd(&cursor[lid], 1)] = (int32_t)(((p / fanout) << 8) | (p % fanout));
  }
}

// The fill above claims slots with atomics, so a row's entries land in arbitrary order -- and k_pool_bwd would sum a source row's
// gradient contributions in a different order from run to run.  Every row is therefore put into ascending entry order
// ((destination << 8) | slot = sampling order): one thread per row (most rows hold 1-3 entries and are finished by their thread),
// rows of up to kRevWarpMax entries by their warp (rank by counting through shared memory), longer ones by a whole CTA.
constexpr int kRevThreadMax = 8, kRevWarpMax = 1024;
__global__ void __launch_bounds__(kBlock) k_rev_sort(const int32_t* __restrict__ rev_ptr, int32_t* __restrict__ rev_edge,
                                                     const int32_t* __restrict__ n_src_dev, int n_src_max, int32_t* __restrict__ long_rows) {
  __shared__ int32_t buf[kBlock / 32][kRevWarpMax];
  const int n = n_src_dev ? min(*n_src_dev, n_src_max) : n_src_max;      // (rows beyond the live count are empty)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nthreads = gridDim.x * blockDim.x;
  for (int r0 = blockIdx.x * blockDim.x + threadIdx.x - lane; r0 < n; r0 += nthreads) {
    const int r = r0 + lane;
    int a = 0, len = 0;
    if (r < n) {
      a = rev_ptr[r];
      len = rev_ptr[r + 1] - a;
      if (len > 1 && len <= kRevThreadMax) {                 // insertion sort in registers
        int32_t e[kRevThreadMax];
#pragma unroll
        for (int i = 0; i < kRevThreadMax; ++i) e[i] = i < len ? rev_edge[a + i] : 0x7fffffff;
#pragma unroll
        for (int i = 1; i < kRevThreadMax; ++i)
#pragma unroll
          for (int j = i; j > 0; --j)
            if (e[j] < e[j - 1]) { const int32_t t = e[j]; e[j] = e[j - 1]; e[j - 1] = t; }
#pragma unroll
        for (int i = 0; i < kRevThreadMax; ++i)
          if (i < len) rev_edge[a + i] = e[i];
      } else if (len > kRevWarpMax) {
        long_rows[1 + atomicAdd(long_rows, 1)] = r;
      }
    }
    unsigned todo = __ballot_sync(0xffffffffu, len > kRevThreadMax && len <= kRevWarpMax);
    while (todo) {
      const int src_lane = __ffs(todo) - 1;
      todo &= todo - 1;
      const int aw = __shfl_sync(0xffffffffu, a, src_lane), lw = __shfl_sync(0xffffffffu, len, src_lane);
      for (int i = lane; i < lw; i += 32) buf[w][i] = rev_edge[aw + i];
      __syncwarp();
      for (int i = lane; i < lw; i += 32) {
        const int32_t e = buf[w][i];
        int rank = 0;
        for (int j = 0; j < lw; ++j) rank += buf[w][j] < e ? 1 : 0;      // entries of a row are distinct
        rev_edge[aw + rank] = e;
      }
      __syncwarp();
    }
  }
}
// rows beyond kRevWarpMax entries (one source picked > 1024 times inside one block: extreme hubs only): one CTA per row
__global__ void __launch_bounds__(1024) k_rev_sort_long(const int32_t* __restrict__ rev_ptr, int32_t* __restrict__ rev_edge,
                                                        int32_t* __restrict__ long_rows, int32_t* __restrict__ copy) {
  const int nl = long_rows[0];
  for (int q = blockIdx.x; q < nl; q += gridDim.x) {
    const int r = long_rows[1 + q];
    const int a = rev_ptr[r], len = rev_ptr[r + 1] - a;
    for (int i = threadIdx.x; i < len; i += blockDim.x) copy[a + i] = rev_edge[a + i];
    __syncthreads();
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      const int32_t e = copy[a + i];
      int rank = 0;
      for (int j = 0; j < len; ++j) rank += copy[a + j] < e ? 1 : 0;
      rev_edge[a + rank] = e;
    }
    __syncthreads();
  }
}

int reverse_edges(ToBlockWs* ws, const int32_t* edge_lid, const int32_t* n_dst_dev, int n_dst_max, int fanout, int n_src_max,
                  int32_t* rev_ptr, int32_t* rev_edge, cudaStream_t s) {
  const int64_t ne_max = (int64_t)n_dst_max * fanout;
  OGL_ARG(n_src_max + 1 <= ws->rev_cap, "reverse_edges: workspace too small");
  OGL_CUDA(cudaMemsetAsync(ws->rev_cnt, 0, sizeof(int32_t) * 2 * (size_t)ws->rev_cap, s));     // counters + cursors
  OGL_LAUNCH(k_rev_count, grid_for(ne_max, kBlock), kBlock, 0, s, edge_lid, n_dst_dev, n_dst_max, fanout, ws->rev_cnt);
  OGL_TRY(exclusive_scan_i32(ws->rev_cnt, rev_ptr, (int64_t)n_src_max + 1, ws->rev_scan_scratch, nullptr, s));   // own scratch: may run beside to_block
  OGL_LAUNCH(k_rev_fill, grid_for(ne_max, kBlock), kBlock, 0, s, edge_lid, n_dst_dev, n_dst_max, fanout, rev_ptr, ws->rev_cnt + ws->rev_cap,
             rev_edge);
  // canonical (ascending) order inside every row: the backward pass is then bit-reproducible.  Rows beyond the live source count
  // are empty (their counters are zero), so the pass simply covers all n_src_max rows
  OGL_CUDA(cudaMemsetAsync(ws->rev_sort_scratch, 0, sizeof(int32_t), s));
  OGL_LAUNCH(k_rev_sort, grid_for(n_src_max, kBlock, 16), kBlock, 0, s, rev_ptr, rev_edge, (const int32_t*)nullptr, n_src_max,
             ws->rev_sort_scratch);
  OGL_LAUNCH(k_rev_sort_long, 32, 1024, 0, s, rev_ptr, rev_edge, ws->rev_sort_scratch, ws->rev_sort_scratch + (ws->ne_max / kRevWarpMax + 8));
  return OGL_OK;
}

int sample_hop(const GraphView& g, const int32_t* dst_nodes, const int32_t* n_dst_dev, int n_dst_max, int fanout, uint64_t seed,
               const uint32_t* step_dev, uint32_t step_imm, uint32_t hop, int32_t* out_src, int64_t* out_eid, cudaStream_t s) {
  const int Q = (fanout + 3) / 4;
  const uint2 key = make_uint2((uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
  OGL_LAUNCH(k_sample, grid_for((int64_t)n_dst_max * Q, kBlock), kBlock, 0, s, g, dst_nodes, n_dst_dev, n_dst_max, fanout, key,
             step_dev, step_imm, hop, out_src, out_eid);
  return OGL_OK;
}

int to_block_init(ToBlockWs* ws, int64_t v_cap, int64_t ne_max, int64_t rows_max) {
  ws->v_cap = v_cap;
  ws->ne_max = ne_max;
  OGL_CUDA(cudaMalloc(&ws->first, sizeof(int32_t) * v_cap));
  OGL_CUDA(cudaMalloc(&ws->flags, sizeof(int32_t) * (ne_max + 1)));
  OGL_CUDA(cudaMalloc(&ws->pos, sizeof(int32_t) * (ne_max + 1)));
  const int64_t scan_n = (ne_max > v_cap ? ne_max : v_cap) + 2;
  OGL_CUDA(cudaMalloc(&ws->scan_scratch, sizeof(int32_t) * scan_scratch_elems(scan_n)));
  ws->rev_cap = (rows_max > v_cap ? rows_max : v_cap) + 1;      // a frontier may exceed v_cap by the duplicates among its seeds
  OGL_CUDA(cudaMalloc(&ws->rev_cnt, sizeof(int32_t) * 2 * (size_t)ws->rev_cap));
  OGL_CUDA(cudaMalloc(&ws->rev_scan_scratch, sizeof(int32_t) * scan_scratch_elems(ws->rev_cap + 1)));
  OGL_CUDA(cudaMalloc(&ws->rev_sort_scratch, sizeof(int32_t) * (size_t)(ne_max / kRevWarpMax + 8 + ne_max + 8)));
  OGL_CUDA(cudaMalloc(&ws->n_new, sizeof(int32_t)));
  OGL_LAUNCH(k_fill_i32, grid_for(v_cap, kBlock), kBlock, 0, 0, ws->first, 0x7fffffff, v_cap);
  OGL_CUDA(cudaDeviceSynchronize());
  return OGL_OK;
}

void to_block_free(ToBlockWs* ws) {
  cudaFree(ws->first); cudaFree(ws->flags); cudaFree(ws->pos); cudaFree(ws->scan_scratch); cudaFree(ws->n_new); cudaFree(ws->rev_cnt); cudaFree(ws->rev_scan_scratch); cudaFree(ws->rev_sort_scratch);
  *ws = ToBlockWs();
}

int to_block(ToBlockWs* ws, const int32_t* dst_nodes, const int32_t* n_dst_dev, int n_dst_max, int fanout, const int32_t* picked,
             int32_t* src_nodes, int32_t* n_src_dev, int n_src_max, int32_t* edge_lid, cudaStream_t s) {
  const int64_t ne_max = (int64_t)n_dst_max * fanout;
  OGL_ARG(ne_max <= ws->ne_max, "to_block: workspace too small");
  OGL_LAUNCH(k_tb_mark, grid_for(n_dst_max + ne_max, kBlock), kBlock, 0, s, dst_nodes, n_dst_dev, n_dst_max, fanout, picked, ws->first);
  OGL_LAUNCH(k_tb_flags, grid_for(ne_max, kBlock), kBlock, 0, s, n_dst_dev, n_dst_max, fanout, picked, ws->first, ws->flags, ne_max);
  OGL_TRY(exclusive_scan_i32(ws->flags, ws->pos, ne_max, ws->scan_scratch, ws->n_new, s));
  OGL_LAUNCH(k_tb_emit, grid_for(n_dst_max + ne_max, kBlock), kBlock, 0, s, dst_nodes, n_dst_dev, n_dst_max, fanout, picked, ws->first,
             ws->flags, ws->pos, ws->n_new, src_nodes, n_src_dev, n_src_max, edge_lid, ne_max);
  OGL_LAUNCH(k_tb_reset, grid_for(n_src_max, kBlock), kBlock, 0, s, src_nodes, n_src_dev, n_src_max, ws->first);
  return OGL_OK;
}

int cast_nodes(const int64_t* in, int32_t* out, int64_t n, int64_t n_vertices, uint32_t* err_flag, cudaStream_t s) {
  if (n > 0) OGL_LAUNCH(k_cast_nodes, grid_for(n, kBlock), kBlock, 0, s, in, out, n, n_vertices, err_flag);
  return OGL_OK;
}

// ---- RBR: keyed Feistel permutation, first n hits below n_pop (oracle/sampler.py:draw_uniform_subset) ----
__device__ __forceinline__ uint32_t feistel(uint32_t x, int bits, uint2 key, uint32_t counter) {
  const int half = bits >> 1;
  const uint32_t mask = (1u << half) - 1u;
  uint32_t l = x >> half, r = x & mask;
#pragma unroll
  for (uint32_t rnd = 0; rnd < 4; ++rnd) {
    const uint32_t f = philox4x32_10(make_uint4(r, rnd, counter, 0x5EEDu), key).x & mask;
    const uint32_t nl = r;
    r = l ^ f;
    l = nl;
  }
  return (l << half) | r;
}

// single CTA: candidates i = 0,1,2,... in rounds of blockDim.x; ordered compaction of hits
__global__ void __launch_bounds__(1024) k_draw_uniform(int64_t n_pop, int64_t n, int bits, uint2 key, uint32_t counter, int64_t* __restrict__ out) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t have = 0;
  const uint64_t domain = 1ull << bits;
  for (uint64_t base = 0; have < n && base < domain; base += blockDim.x) {
    const uint64_t i = base + threadIdx.x;
    uint32_t c = 0;
    bool hit = false;
    if (i < domain) {
      c = feistel((uint32_t)i, bits, key, counter);
      hit = (int64_t)c < n_pop;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
      int v = s_warp[lane];
      int inc = v;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      s_warp[lane] = inc - v;
      if (lane == 31) s_total = inc;
    }
    __syncthreads();
    if (hit) {
      const int64_t at = have + s_warp[warp] + __popc(m & ((1u << lane) - 1));
      if (at < n) out[at] = (int64_t)c;
    }
    have += s_total;
    __syncthreads();
  }
}

}  // namespace ogl

using namespace ogl;

extern "C" int ogl_draw_uniform(int64_t n_pop, int64_t n, uint64_t seed, uint32_t counter, int64_t* out_idx_dev, void* stream) {
  OGL_TRY(require_device());
  OGL_ARG(n_pop >= 0 && n >= 0 && n <= n_pop && out_idx_dev && n_pop < (1LL << 31), "ogl_draw_uniform: bad arguments");
  if (n == 0) return OGL_OK;
  int bits = 2;
  while ((1LL << bits) < n_pop) ++bits;
  bits += bits & 1;
  const uint2 key = make_uint2((uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
  OGL_LAUNCH(k_draw_uniform, 1, 1024, 0, stream, n_pop, n, bits, key, counter, out_idx_dev);
  return OGL_OK;
}

extern "C" int ogl_sample_neighbors(ogl_graph* g, const int64_t* dst_dev, int64_t n, int fanout, uint64_t seed, uint32_t step,
                                    uint32_t hop, int32_t* out_src_dev, int64_t* out_eid_dev, void* stream) {
  OGL_TRY(require_device());
  OGL_ARG(g && dst_dev && out_src_dev && n >= 0 && n < (1LL << 31) && fanout > 0, "ogl_sample_neighbors: bad arguments");
  if (n == 0) return OGL_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int32_t* tmp = nullptr;
  {
    // the int32 row list is stream-ordered scratch: keep the default pool from handing its memory back to the driver at every
    // synchronisation (release threshold 0), which made single calls cost tens of milliseconds now and then
    static bool pool_kept = false;
    if (!pool_kept) {
      int dev = 0;
      cudaMemPool_t pool = nullptr;
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
      pool_kept = true;
    }
  }
  OGL_CUDA(cudaMallocAsync(&tmp, sizeof(int32_t) * n, s));
  int r = cast_nodes(dst_dev, tmp, n, graph_view(g).n_vertices, nullptr, s);
  if (r == OGL_OK) r = sample_hop(graph_view(g), tmp, nullptr, (int)n, fanout, seed, nullptr, step, hop, out_src_dev, out_eid_dev, s);
  cudaFreeAsync(tmp, s);
  return r;
}
