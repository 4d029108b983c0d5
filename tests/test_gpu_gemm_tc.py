"""tcgen05 / TMA / TMEM GEMM kernels vs a plain PyTorch fp32 reference of the same op (bf16 operands, fp32
accumulation: |err| <= 2e-3 * max|ref| + 1e-3 * |ref|), through the C ABI.  Shapes cover single / multiple
k-stages (ring wrap), ragged K / N / M tails, the narrower last N tile, pad columns that must not be read,
the split-over-rows TN path with and without workspace."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(got, ref, what):
    err = (got - ref).abs()
    bound = 2e-3 * ref.abs().max() + 1e-3 * ref.abs()
    assert bool((err <= bound).all()), "%s: max err %.3e (scale %.3e)" % (what, err.max().item(), ref.abs().max().item())


@pytest.mark.parametrize("m,n,k", [(128, 16, 16), (128, 64, 64), (128, 256, 512), (128, 48, 41), (100, 24, 50), (1000, 602, 602),
                                   (2500, 600, 1204), (384, 300, 64), (26000, 41, 600), (1, 8, 8), (129, 257, 65)])
def test_gemm_nt_tcgen05(m, n, k):
    from ogl_b200 import native
    torch.manual_seed(m + n + k)
    ld = (k + 7) // 8 * 8
    a = torch.full((m, ld), 1000.0, dtype=torch.bfloat16, device="cuda")      # poisoned pad columns
    b = torch.full((n, ld), 1000.0, dtype=torch.bfloat16, device="cuda")
    a[:, :k] = torch.randn(m, k, device="cuda")
    b[:, :k] = torch.randn(n, k, device="cuda")
    got = native.gemm_bf16_nt(a, b, k=k)
    _check(got, a[:, :k].float() @ b[:, :k].float().t(), "NT %dx%dx%d" % (m, n, k))


@pytest.mark.parametrize("m,n,k,ws", [(64, 64, 64, 1 << 22), (512, 128, 256, 1 << 22), (1000, 602, 602, 1 << 24), (100, 41, 600, 1 << 22),
                                      (5000, 600, 41, 1 << 22), (30000, 602, 602, 1 << 24), (200, 24, 50, 1 << 22), (1000, 602, 602, 0),
                                      (1, 8, 8, 0), (4097, 166, 166, 1 << 22)])
def test_gemm_tn_tcgen05(m, n, k, ws):
    from ogl_b200 import native
    torch.manual_seed(m + n + k)
    ldn, ldk = (n + 7) // 8 * 8, (k + 7) // 8 * 8
    a = torch.full((m, ldn), 1000.0, dtype=torch.bfloat16, device="cuda")
    b = torch.full((m, ldk), 1000.0, dtype=torch.bfloat16, device="cuda")
    a[:, :n] = torch.randn(m, n, device="cuda")
    b[:, :k] = torch.randn(m, k, device="cuda")
    got = native.gemm_bf16_tn(a, b, n=n, k=k, workspace_elems=ws)
    _check(got, a[:, :n].float().t() @ b[:, :k].float(), "TN %dx%dx%d" % (m, n, k))


def test_tn_is_deterministic():
    from ogl_b200 import native
    torch.manual_seed(0)
    a = torch.randn(20000, 608, device="cuda").bfloat16()
    c1 = native.gemm_bf16_tn(a, a, n=602, k=602)
    c2 = native.gemm_bf16_tn(a, a, n=602, k=602)
    assert torch.equal(c1, c2)
