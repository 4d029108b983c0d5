// Dense GEMM entry points of the GraphSAGE-pool path (SIMT fp32-accumulate + tcgen05 bf16 / fp16 / tf32).
#pragma once
#include "common.cuh"

namespace ogl {

// C[m, n] = act( sum_seg sum_k A_seg[m, k] * B_seg[n, k] + bias[n] ) , optional relu / mask.
// A, B row-major with the contraction index contiguous ("NT").  Up to two (A, B, K) segments
// accumulate into the same output (used for [h_self | neigh] x [W_self | W_neigh]^T).
struct GemmNT {
  const void* a[2] = {nullptr, nullptr};   // element type = in_bf16 ? bf16 : f32
  int lda[2] = {0, 0};
  const void* b[2] = {nullptr, nullptr};
  int ldb[2] = {0, 0};
  int k[2] = {0, 0};
  int n_seg = 1;
  const int32_t* a_rows_dev[2] = {nullptr, nullptr};  // rows of segment s valid for m < *a_rows_dev[s] (nullptr: all)
  int a_rows_max[2] = {0, 0};      // rows that exist in segment s's buffer (0: m_max); reads beyond are zero-filled (TMA)
  int force_cg = 0;                // tcgen05 path: 0 = auto, 1 = single CTA, 2 = CTA pairs (cta_group::2)
  const float* bias = nullptr;     // [n] fp32 (nullable)
  const float* bias2 = nullptr;    // second bias added too (fc_self.bias + fc_neigh.bias)
  int relu = 0;
  float alpha = 1.f;               // scales the accumulated product before the bias (feat_drop's 1 / (1 - p) in the backward pass)
  const void* mask = nullptr;      // same shape/type as A-typed [m, ldmask]: out = mask>0 ? out : 0
  int ldmask = 0;
  void* c = nullptr;               // out_bf16 ? bf16 : f32
  int ldc = 0;
  int m_max = 0;                   // static row bound (grid size)
  const int32_t* m_dev = nullptr;  // dynamic row count on device (nullptr: m_max)
  int n = 0;
  int in_bf16 = 0, out_bf16 = 0;   // 16-bit operands / output ...
  int f16 = 0;                     // ... which are fp16 instead of bf16 (mode OGL_FP16)
  int tf32 = 0;                    // mode OGL_TF32: operands are fp32 holding TF32-rounded values (tcgen05 kind::tf32 / exact on the SIMT path)
  int out_tf32 = 0;                // ... and the fp32 output is rounded to TF32 too (it is a later GEMM's operand)
  int zero_tail = 1;               // write zeros to rows [m, round_up(m, 128)) (contraction padding for later TN GEMMs)
};

// C[n, k] = sum_{m < *m_dev} A[m, n] * B[m, k]   (fp32 out; weight gradients)
struct GemmTN {
  const void* a = nullptr;   // [m, lda], n contiguous
  int lda = 0;
  const void* b = nullptr;   // [m, ldb], k contiguous
  int ldb = 0;
  float* c = nullptr;        // [n, ldc]
  int ldc = 0;
  int n = 0, k = 0;
  int m_max = 0;
  const int32_t* m_dev = nullptr;
  int in_bf16 = 0;           // 16-bit operands ...
  int f16 = 0;               // ... which are fp16 instead of bf16 (mode OGL_FP16)
  int tf32 = 0;              // operands are fp32 holding TF32-rounded values
  float alpha = 1.f;         // C = alpha * A^T B (mode OGL_FP16: undoes the loss scale carried by the activation gradients)
  float* partial = nullptr;  // split workspace [splits, n, k]
  int64_t partial_elems = 0;
};

int gemm_nt_simt(const GemmNT& g, cudaStream_t s);
int gemm_tn_simt(const GemmTN& g, cudaStream_t s);

// tcgen05 / TMA / TMEM versions (bf16 or tf32 operands, fp32 accumulate); same contracts
int gemm_nt_tc(const GemmNT& g, cudaStream_t s);
int gemm_tn_tc(const GemmTN& g, cudaStream_t s);
// up to 4 weight-gradient GEMMs contracting over the same rows (same m_dev / m_max) in one launch; g[0].partial is the workspace
int gemm_tn_tc_group(const GemmTN* g, int count, cudaStream_t s);

// deterministic reduction of the split partials of up to 4 problems in one launch: c = sum over z (ascending) of partial[z]
struct ReduceGroup {
  struct Problem { const float* partial; float* c; int n, k, ldp, ldc; float alpha; } pr[4];
  int count, splits;
};
int reduce_splits_group(const ReduceGroup& g, cudaStream_t s);
bool gemm_tc_available();

}  // namespace ogl
