// Peer-memory gradient exchange fused with Adam (peer.cu), called from plan.cu.
#pragma once
#include "sage_kernels.cuh"

struct ogl_peer;

namespace ogl {

struct PeerAdamArgs {
  int mode = 0;                      // OGL_F32 | OGL_BF16 | OGL_TF32: element type of the weight shadows
  float* params = nullptr;
  const float* grads = nullptr;      // the plan's gradient buffer: must be the peer group's local buffer
  float *m = nullptr, *v = nullptr;
  float lr = 0.f, b1 = 0.f, b2 = 0.f, eps = 0.f;
  const uint32_t* t_dev = nullptr;   // Adam step counter (the kernel uses *t_dev + 1; the caller bumps it after the last bucket)
  const ShadowSeg* segs = nullptr;
  int n_segs = 0;
  float* reduced_out = nullptr;      // optional: the summed gradients of the range (tests)
};

int peer_sum_adam(ogl_peer* p, const PeerAdamArgs& a, int64_t lo, int64_t hi, cudaStream_t s);
// enqueue: wait until every peer has finished reading this rank's gradients of the previous exchange
int peer_wait_readers(ogl_peer* p, cudaStream_t s);

}  // namespace ogl
