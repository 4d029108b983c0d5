"""Perf experiments on the NT tcgen05 kernel (bf16 output + bias + ReLU = the fc_pool GEMM of the Reddit-shaped step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ogl_b200 import native

m, n, k = 89000, int(os.environ.get("EXP_N", "602")), int(os.environ.get("EXP_K", "602"))
ldk = (k + 7) // 8 * 8
a = torch.randn(m, ldk, device="cuda").bfloat16()
b = (torch.randn(n, ldk, device="cuda") * 0.05).bfloat16()
bias = torch.randn(n, device="cuda")
ref = torch.relu(a[:, :k].float() @ b[:, :k].float().t() + bias)


def bench(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for cg in (1, 2):
    out = native.gemm_bf16_nt_ex(a, b, k=k, out_bf16=True, bias=bias, relu=True, cg=cg)
    err = (out.float() - ref).abs().max().item()
    ms = bench(lambda: native.gemm_bf16_nt_ex(a, b, k=k, out_bf16=True, bias=bias, relu=True, cg=cg))
    print("cg=%d dbg=%s  %.4f ms  %.0f TFLOP/s  max_err %.3g" % (cg, os.environ.get("OGL_GEMM_DBG", "0"), ms, 2.0 * m * n * k / ms / 1e9, err), flush=True)
