"""Ablations of the NT tcgen05 kernel at the fc_pool shape of the Reddit step (fp16 operands, fp16 TMA-store epilogue).
OGL_GEMM_DBG: 1 = epilogue drains the accumulator without storing, 2 = no operand loads (MMA issue on stale shared memory), 3 = both."""
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


for (m, n, k) in ((89000, 602, 602), (18432, 600, 600), (89000, 512, 602), (89000, 256, 602)):
    ldk = (k + 7) // 8 * 8
    a = torch.randn(m, ldk, device="cuda").half()
    b = (torch.randn(n, ldk, device="cuda") * 0.05).half()
    bias = torch.randn(n, device="cuda")
    fl = 2.0 * m * n * k
    for cg in (2,):
        ms = bench(lambda: native.gemm_f16_nt_ex(a, b, k=k, out_f16=True, bias=bias, relu=True, cg=cg))
        print("NT fp16 m=%d n=%d k=%d cg=%d dbg=%s tag=%s  %.4f ms  %.0f TFLOP/s" % (m, n, k, cg, dbg, os.environ.get("TAG", ""), ms, fl / ms / 1e9), flush=True)
