"""Probe of the tcgen05 GEMM kernels on a B200: escalating shapes, prints the error structure of each case
(run under `timeout`; a hang or fault here must not take the test suite with it)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ogl_b200
from ogl_b200 import native

torch.manual_seed(0)


def report(name, got, ref):
    err = (got - ref).abs()
    scale = ref.abs().max().item()
    bad = err > 2e-3 * scale + 1e-3 * ref.abs()
    print("%-44s max_err %.3e scale %.3e bad %d/%d %s" % (name, err.max().item(), scale, int(bad.sum()), bad.numel(),
                                                         "OK" if not bad.any() else "MISMATCH"), flush=True)
    if bad.any():
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        print("   bad rows: n=%d first %s last %s | bad cols: n=%d first %s last %s" % (
            len(rows), rows[:6].tolist(), rows[-3:].tolist(), len(cols), cols[:6].tolist(), cols[-3:].tolist()), flush=True)
    return not bad.any()


def nt(m, n, k, pad=True):
    ld = (k + 7) // 8 * 8
    a = torch.zeros(m, ld, dtype=torch.bfloat16, device="cuda")
    b = torch.zeros(n, ld, dtype=torch.bfloat16, device="cuda")
    a[:, :k] = torch.randn(m, k, device="cuda")
    b[:, :k] = torch.randn(n, k, device="cuda")
    if ld > k:                                   # poison the pad columns: the kernel must not read them
        a[:, k:] = 1000.0
        b[:, k:] = 1000.0
    got = native.gemm_bf16_nt(a, b, k=k)
    torch.cuda.synchronize()
    ref = a[:, :k].float() @ b[:, :k].float().t()
    return report("NT m=%d n=%d k=%d" % (m, n, k), got, ref)


def tn(m, n, k, ws=1 << 24):
    ldn, ldk = (n + 7) // 8 * 8, (k + 7) // 8 * 8
    a = torch.zeros(m, ldn, dtype=torch.bfloat16, device="cuda")
    b = torch.zeros(m, ldk, dtype=torch.bfloat16, device="cuda")
    a[:, :n] = torch.randn(m, n, device="cuda")
    b[:, :k] = torch.randn(m, k, device="cuda")
    got = native.gemm_bf16_tn(a, b, n=n, k=k, workspace_elems=ws)
    torch.cuda.synchronize()
    ref = a[:, :n].float().t() @ b[:, :k].float()
    return report("TN m=%d n=%d k=%d ws=%d" % (m, n, k, ws), got, ref)


def bench(fn, flops, name, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%-44s %.3f ms  %.1f TFLOP/s" % (name, ms, flops / ms / 1e9), flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
ok = True
if which in ("all", "nt"):
    for shp in [(128, 16, 16), (128, 64, 64), (128, 64, 128), (128, 256, 512), (128, 48, 41), (100, 24, 50), (1000, 602, 602),
                (2500, 600, 1204), (384, 300, 64), (26000, 41, 600)]:
        ok &= nt(*shp)
if which in ("all", "tn"):
    for shp in [(64, 64, 64), (64, 128, 64), (128, 128, 256), (512, 128, 256), (1000, 602, 602), (100, 41, 600), (5000, 600, 41),
                (30000, 602, 602), (200, 24, 50)]:
        ok &= tn(*shp)
    ok &= tn(1000, 602, 602, ws=0)
if which in ("all", "perf"):
    m, n, k = 150000, 602, 602
    a = torch.randn(m, 608, device="cuda").bfloat16()
    b = torch.randn(n, 608, device="cuda").bfloat16()
    bench(lambda: native.gemm_bf16_nt(a, b, k=k), 2.0 * m * n * k, "NT perf m=%d n=%d k=%d" % (m, n, k))
    bench(lambda: a[:, :k].float() if False else torch.matmul(a, b.t()), 2.0 * m * n * 608, "torch bf16 matmul same shape (cuBLAS)")
    bench(lambda: native.gemm_bf16_tn(a, a, n=n, k=k), 2.0 * m * n * k, "TN perf m=%d n=%d k=%d" % (m, n, k))
    bench(lambda: torch.matmul(a.t(), a), 2.0 * m * 608 * 608, "torch bf16 A^T A (cuBLAS)")
print("PROBE", "PASS" if ok else "FAIL")
