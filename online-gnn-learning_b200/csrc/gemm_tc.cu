// tcgen05 / TMA / TMEM GEMMs (bf16 operands, fp32 accumulation in tensor memory).
#include "gemm.cuh"
namespace ogl {
bool gemm_tc_available() { return false; }
int gemm_nt_tc(const GemmNT&, cudaStream_t) { set_error("gemm_nt_tc: not built"); return OGL_ERR_ARG; }
int gemm_tn_tc(const GemmTN&, cudaStream_t) { set_error("gemm_tn_tc: not built"); return OGL_ERR_ARG; }
}  // namespace ogl
extern "C" int ogl_gemm_bf16_nt(const void*, int, const void*, int, float*, int, int, int, int, void*) {
  ogl::set_error("ogl_gemm_bf16_nt: not built");
  return OGL_ERR_ARG;
}
