// fp64 sum-tree for prioritized rehearsal (PBR).
//
// Device restatement of train/prioritized_replay/segment_tree.py:69-79 (leaf write + parent
// recompute, parent = left + right), :30-67 (reduce with the reference's association order),
// :118-125 (find_prefixsum_idx) and replay_buffer.py:110-130,164-181,219-244 (priority
// transform + stratified proportional draw).  All tree arithmetic is IEEE fp64 with explicit
// round-to-nearest intrinsics (no FMA contraction), so it is bit-exact with the Python floats.
#include "common.cuh"

namespace ogl {

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) k_tree_write(double* __restrict__ value, int64_t cap, const int64_t* __restrict__ idx,
                                                       const double* __restrict__ val, int64_t n) {
  // (indices outside [0, cap) are skipped in all three update kernels: the host mirrors raise before the call, as the reference's
  // SegmentTree.__setitem__ asserts; a raw C-ABI caller gets no out-of-bounds write either)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if ((uint64_t)idx[i] < (uint64_t)cap) value[cap + idx[i]] = val[i];
}

// one level up: parent of every written leaf at height `level` (duplicates write identical values)
__global__ void __launch_bounds__(kBlock) k_tree_level(double* value, int64_t cap, const int64_t* __restrict__ idx, int64_t n, int level) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if ((uint64_t)idx[i] >= (uint64_t)cap) continue;
    const int64_t node = (cap + idx[i]) >> level;
    value[node] = __dadd_rn(value[2 * node], value[2 * node + 1]);
  }
}

// small batches: one CTA walks all levels (block barrier + fence between levels)
__global__ void __launch_bounds__(1024) k_tree_update_small(double* value, int64_t cap, const int64_t* __restrict__ idx,
                                                            const double* __restrict__ val, int64_t n, int levels) {
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
    if ((uint64_t)idx[i] < (uint64_t)cap) value[cap + idx[i]] = val[i];
  __threadfence_block();
  __syncthreads();
  for (int level = 1; level <= levels; ++level) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      if ((uint64_t)idx[i] >= (uint64_t)cap) continue;
      const int64_t node = (cap + idx[i]) >> level;
      const double s = __dadd_rn(((volatile double*)value)[2 * node], ((volatile double*)value)[2 * node + 1]);
      ((volatile double*)value)[node] = s;
    }
    __threadfence_block();
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kBlock) k_tree_find(const double* __restrict__ value, int64_t cap, const double* __restrict__ mass,
                                                      int64_t n, int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double m = mass[i];
    int64_t node = 1;
    while (node < cap) {
      const double left = value[2 * node];
      if (left > m) {
        node = 2 * node;
      } else {
        m = __dsub_rn(m, left);
        node = 2 * node + 1;
      }
    }
    out[i] = node - cap;
  }
}

// segment_tree.py:30-45, end inclusive; same association order of the additions
__device__ double tree_reduce(const double* value, int64_t start, int64_t end, int64_t node, int64_t ns, int64_t ne) {
  if (start == ns && end == ne) return value[node];
  const int64_t mid = (ns + ne) / 2;
  if (end <= mid) return tree_reduce(value, start, end, 2 * node, ns, mid);
  if (mid + 1 <= start) return tree_reduce(value, start, end, 2 * node + 1, mid + 1, ne);
  const double a = tree_reduce(value, start, mid, 2 * node, ns, mid);
  const double b = tree_reduce(value, mid + 1, end, 2 * node + 1, mid + 1, ne);
  return __dadd_rn(a, b);
}

__global__ void k_tree_sum(const double* value, int64_t cap, int64_t lo, int64_t hi, double* out) {
  // [lo, hi) -> inclusive end like SegmentTree.reduce (:62-67)
  *out = (hi - 1 < lo) ? 0.0 : tree_reduce(value, lo, hi - 1, 1, 0, cap - 1);
}

__global__ void __launch_bounds__(kBlock) k_tree_stratified(const double* __restrict__ value, int64_t cap, const double* __restrict__ u,
                                                            int64_t n, const double* __restrict__ p_total, int64_t* __restrict__ out) {
  const double every = __ddiv_rn(*p_total, (double)n);          // replay_buffer.py:170
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double m = __dadd_rn(__dmul_rn(u[i], every), __dmul_rn((double)i, every));   // :178
    int64_t node = 1;
    while (node < cap) {
      const double left = value[2 * node];
      if (left > m) {
        node = 2 * node;
      } else {
        m = __dsub_rn(m, left);
        node = 2 * node + 1;
      }
    }
    out[i] = node - cap;
  }
}

// ---- loss -> leaf transform (replay_buffer.py:110-130 + :219-244) ------------------------------
// phase 1 (single CTA, deterministic): clip, log, fold batch min/max into the running state
__global__ void __launch_bounds__(1024) k_prio_minmax(const float* __restrict__ loss, int64_t n, double clip_lo, double clip_hi,
                                                      double* __restrict__ mm /* min_val,max_val,min_log,max_log */) {
  __shared__ double s[4][32];
  double mn = 1e300, mx = -1e300;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    double p = (double)loss[i];
    p = fmin(fmax(p, clip_lo), clip_hi);
    mn = fmin(mn, p);
    mx = fmax(mx, p);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s[0][warp] = mn; s[1][warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0 && n > 0) {
    for (int w = 1; w < 32; ++w) { mn = fmin(mn, s[0][w]); mx = fmax(mx, s[1][w]); }
    // log is monotone: extreme of the logs = log of the extremes
    if (mx > mm[1]) mm[1] = mx;
    if (mn < mm[0]) mm[0] = mn;
    const double lmn = log(mn), lmx = log(mx);
    if (lmx > mm[3]) mm[3] = lmx;
    if (lmn < mm[2]) mm[2] = lmn;
  }
}

__global__ void __launch_bounds__(kBlock) k_prio_leaf(const float* __restrict__ loss, int64_t n, double clip_lo, double clip_hi, double eps,
                                                      double alpha, const double* __restrict__ mm, double* __restrict__ leaf) {
  const double lmin = mm[2], scale = __dsub_rn(mm[3], mm[2]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double p = (double)loss[i];
    p = fmin(fmax(p, clip_lo), clip_hi);
    const double l = log(p);
    double v = __dsub_rn(l, lmin);
    if (scale > 0) v = __ddiv_rn(v, scale);
    v = __dadd_rn(v, eps);
    leaf[i] = pow(v, alpha);
  }
}

}  // namespace ogl

using namespace ogl;

struct ogl_sumtree {
  int64_t cap = 0;
  int levels = 0;
  double* value = nullptr;     // [2*cap]
  double* scalar = nullptr;    // p_total scratch
  double* leaf_tmp = nullptr;
  int64_t leaf_tmp_n = 0;
};

extern "C" int ogl_sumtree_create(ogl_sumtree** out, int64_t capacity) {
  OGL_TRY(require_device());
  OGL_ARG(out && capacity > 0 && (capacity & (capacity - 1)) == 0, "ogl_sumtree_create: capacity must be a power of two");
  ogl_sumtree* t = new ogl_sumtree();
  t->cap = capacity;
  while ((1LL << t->levels) < capacity) t->levels++;
  OGL_CUDA(cudaMalloc(&t->value, sizeof(double) * 2 * capacity));
  OGL_CUDA(cudaMemset(t->value, 0, sizeof(double) * 2 * capacity));
  OGL_CUDA(cudaMalloc(&t->scalar, sizeof(double) * 4));
  *out = t;
  return OGL_OK;
}

extern "C" int ogl_sumtree_destroy(ogl_sumtree* t) {
  if (!t) return OGL_OK;
  cudaFree(t->value); cudaFree(t->scalar); cudaFree(t->leaf_tmp);
  delete t;
  return OGL_OK;
}

extern "C" int ogl_sumtree_set(ogl_sumtree* t, const int64_t* idx_dev, const double* val_dev, int64_t n, void* stream) {
  OGL_ARG(t && n >= 0 && (n == 0 || (idx_dev && val_dev)), "ogl_sumtree_set: bad arguments");
  if (n == 0) return OGL_OK;
  if (n <= 16384) {
    OGL_LAUNCH(k_tree_update_small, 1, 1024, 0, stream, t->value, t->cap, idx_dev, val_dev, n, t->levels);
  } else {
    OGL_LAUNCH(k_tree_write, grid_for(n, kBlock), kBlock, 0, stream, t->value, t->cap, idx_dev, val_dev, n);
    for (int level = 1; level <= t->levels; ++level)
      OGL_LAUNCH(k_tree_level, grid_for(n, kBlock), kBlock, 0, stream, t->value, t->cap, idx_dev, n, level);
  }
  return OGL_OK;
}

extern "C" int ogl_sumtree_set_from_loss(ogl_sumtree* t, const int64_t* idx_dev, const float* loss_dev, int64_t n, double clip_lo,
                                         double clip_hi, double eps, double alpha, double* minmax_io_dev, void* stream) {
  OGL_ARG(t && n >= 0 && minmax_io_dev && (n == 0 || (idx_dev && loss_dev)), "ogl_sumtree_set_from_loss: bad arguments");
  if (n == 0) return OGL_OK;
  if (t->leaf_tmp_n < n) {
    if (t->leaf_tmp) cudaFree(t->leaf_tmp);
    OGL_CUDA(cudaMalloc(&t->leaf_tmp, sizeof(double) * n));
    t->leaf_tmp_n = n;
  }
  OGL_LAUNCH(k_prio_minmax, 1, 1024, 0, stream, loss_dev, n, clip_lo, clip_hi, minmax_io_dev);
  OGL_LAUNCH(k_prio_leaf, grid_for(n, kBlock), kBlock, 0, stream, loss_dev, n, clip_lo, clip_hi, eps, alpha, minmax_io_dev, t->leaf_tmp);
  return ogl_sumtree_set(t, idx_dev, t->leaf_tmp, n, stream);
}

extern "C" int ogl_sumtree_sum(ogl_sumtree* t, int64_t lo, int64_t hi, double* out_dev, void* stream) {
  OGL_ARG(t && out_dev && lo >= 0 && hi <= t->cap, "ogl_sumtree_sum: bad arguments");
  OGL_LAUNCH(k_tree_sum, 1, 1, 0, stream, t->value, t->cap, lo, hi, out_dev);
  return OGL_OK;
}

extern "C" int ogl_sumtree_find(ogl_sumtree* t, const double* mass_dev, int64_t n, int64_t* out_idx_dev, void* stream) {
  OGL_ARG(t && n >= 0 && (n == 0 || (mass_dev && out_idx_dev)), "ogl_sumtree_find: bad arguments");
  if (n == 0) return OGL_OK;
  OGL_LAUNCH(k_tree_find, grid_for(n, kBlock), kBlock, 0, stream, t->value, t->cap, mass_dev, n, out_idx_dev);
  return OGL_OK;
}

extern "C" int ogl_sumtree_sample_stratified(ogl_sumtree* t, const double* uniforms_dev, int64_t n, int64_t n_items, int64_t* out_idx_dev,
                                             void* stream) {
  OGL_ARG(t && n > 0 && n_items >= 1 && n_items <= t->cap && uniforms_dev && out_idx_dev, "ogl_sumtree_sample_stratified: bad arguments");
  // p_total = sum(0, n_items - 1): the LAST leaf is excluded, as in replay_buffer.py:169
  OGL_LAUNCH(k_tree_sum, 1, 1, 0, stream, t->value, t->cap, (int64_t)0, n_items - 1, t->scalar);
  OGL_LAUNCH(k_tree_stratified, grid_for(n, kBlock), kBlock, 0, stream, t->value, t->cap, uniforms_dev, n, t->scalar, out_idx_dev);
  return OGL_OK;
}

extern "C" int ogl_sumtree_values(ogl_sumtree* t, const double** value_dev, int64_t* capacity) {
  OGL_ARG(t && value_dev && capacity, "null");
  *value_dev = t->value;
  *capacity = t->cap;
  return OGL_OK;
}
