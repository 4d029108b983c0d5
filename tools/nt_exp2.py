"""MMA-only (OGL_GEMM_DBG=3) timing of the NT tcgen05 kernel against tile width N, contraction depth K and cta_group."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ogl_b200 import native

dbg = os.environ.get("OGL_GEMM_DBG", "0")


def bench(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


m = 74 * 256 * 8        # exactly 8 row blocks of 256 per CTA pair
for k in (608, 1216):
    for n in (64, 96, 128, 160, 192, 256):
        a = torch.randn(m, k, device="cuda").half()
        b = (torch.randn(n, k, device="cuda") * 0.05).half()
        bias = torch.randn(n, device="cuda")
        for cg in (1, 2):
            ms = bench(lambda: native.gemm_f16_nt_ex(a, b, k=k, out_f16=True, bias=bias, relu=True, cg=cg))
            tiles_per_unit = 8 if cg == 2 else (m // 128 + 147) // 148
            clk = ms * 1e-3 * 1.965e9
            n_mma = k // 16
            print("dbg=%s cg=%d n=%3d k=%4d  %.4f ms  %.0f TFLOP/s  clk/tile %.0f  clk/mma %.1f (ideal %.1f)" % (
                dbg, cg, n, k, ms, 2.0 * m * n * k / ms / 1e9, clk / tiles_per_unit, clk / tiles_per_unit / n_mma, n / 2.0), flush=True)
