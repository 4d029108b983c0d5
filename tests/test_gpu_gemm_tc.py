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


# ---- tcgen05 kind::tf32 (mode OGL_TF32): operands are fp32 values already rounded to TF32, so the tensor core reads them exactly
# and the result differs from the fp64 product of the same operands only by the fp32 accumulation: |err| <= 1e-6 * (sqrt(K) + 8) * scale
def _check_tf32(got, ref64, k, what):
    err = (got.double() - ref64).abs()
    scale = ref64.abs().max()
    assert bool((err <= 1e-6 * (k ** 0.5 + 8) * scale + 1e-30).all()), "%s: max err %.3e (scale %.3e)" % (what, err.max().item(), scale.item())


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("m,n,k", [(128, 16, 8), (128, 64, 32), (128, 256, 512), (128, 48, 41), (100, 24, 50), (1000, 602, 602),
                                   (2500, 600, 1204), (384, 300, 64), (26000, 41, 600), (1, 8, 8), (129, 257, 65), (40000, 602, 602)])
def test_gemm_nt_tcgen05_tf32(m, n, k, cg):
    from ogl_b200 import native
    torch.manual_seed(m + n + k)
    ld = (k + 7) // 8 * 8
    a = torch.full((m, ld), 1000.0, dtype=torch.float32, device="cuda")      # poisoned pad columns
    b = torch.full((n, ld), 1000.0, dtype=torch.float32, device="cuda")
    a[:, :k] = native.round_tf32(torch.randn(m, k, device="cuda"))
    b[:, :k] = native.round_tf32(torch.randn(n, k, device="cuda"))
    bias = torch.randn(n, device="cuda")
    ref = a[:, :k].double() @ b[:, :k].double().t()
    got = native.gemm_tf32_nt_ex(a, b, k=k, tma_out=False, cg=cg)          # direct fp32 stores (the logits epilogue)
    _check_tf32(got, ref, k, "NT tf32 %dx%dx%d cg%d" % (m, n, k, cg))
    # activation epilogue: bias + ReLU + mask, TF32-rounded, TMA stores
    mask = torch.randn(m, (n + 7) // 8 * 8, device="cuda")
    got2 = native.gemm_tf32_nt_ex(a, b, k=k, tma_out=True, bias=bias, relu=True, mask=mask, cg=cg)
    ref2 = torch.relu(ref + bias.double()) * (mask[:, :n] > 0)
    err = (got2.double() - ref2).abs()
    assert bool((err <= 2 ** -11 * ref2.abs() + 1e-6 * (k ** 0.5 + 8) * ref.abs().max()).all()), "tf32 activation epilogue: max err %.3e" % err.max().item()
    assert torch.equal(got2, native.round_tf32(got2)), "activation output is not TF32-rounded"


@pytest.mark.parametrize("m,n,k,ws", [(64, 64, 64, 1 << 22), (512, 128, 256, 1 << 22), (1000, 602, 602, 1 << 24), (100, 41, 600, 1 << 22),
                                      (5000, 600, 41, 1 << 22), (30000, 602, 602, 1 << 24), (200, 24, 50, 1 << 22), (1000, 602, 602, 0),
                                      (1, 8, 8, 0), (4097, 166, 166, 1 << 22)])
def test_gemm_tn_tcgen05_tf32(m, n, k, ws):
    from ogl_b200 import native
    torch.manual_seed(m + n + k)
    ldn, ldk = (n + 7) // 8 * 8, (k + 7) // 8 * 8
    a = torch.full((m, ldn), 1000.0, dtype=torch.float32, device="cuda")
    b = torch.full((m, ldk), 1000.0, dtype=torch.float32, device="cuda")
    a[:, :n] = native.round_tf32(torch.randn(m, n, device="cuda"))
    b[:, :k] = native.round_tf32(torch.randn(m, k, device="cuda"))
    got = native.gemm_tf32_tn(a, b, n=n, k=k, workspace_elems=ws)
    _check_tf32(got, a[:, :n].double().t() @ b[:, :k].double(), m, "TN tf32 %dx%dx%d" % (m, n, k))


def test_nt_persistent_loop_many_tiles_per_cta():
    """the bench shape of fc_pool: ~1,040 output tiles over 74 CTA pairs (14 per pair) exercise the persistent tile loop, the TMEM
    accumulator ping-pong and the ring wrap together with the bias + ReLU + bf16 TMA-store epilogue"""
    from ogl_b200 import native
    torch.manual_seed(7)
    m, n, k = 89000, 602, 602
    a = torch.zeros(m, 608, dtype=torch.bfloat16, device="cuda")
    b = torch.zeros(n, 608, dtype=torch.bfloat16, device="cuda")
    a[:, :k] = torch.randn(m, k, device="cuda")
    b[:, :k] = torch.randn(n, k, device="cuda") * 0.05
    bias = torch.randn(n, device="cuda")
    got = native.gemm_bf16_nt_ex(a, b, k=k, out_bf16=True, bias=bias, relu=True).float()
    ref = torch.relu(a[:, :k].float() @ b[:, :k].float().t() + bias)
    err = (got - ref).abs()
    assert bool((err <= 2 ** -8 * ref.abs() + 1e-3 * ref.abs().max()).all()), "max err %.3e" % err.max().item()


# ---- tcgen05 kind::f16 with fp16 operands (mode OGL_FP16): products of fp16 values are exact in fp32, so -- like the tf32 flavour -- the
# result differs from the fp64 product of the same operands only by the fp32 accumulation
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("m,n,k", [(128, 16, 16), (128, 256, 512), (100, 24, 50), (1000, 602, 602), (2500, 600, 1204), (26000, 41, 600),
                                   (1, 8, 8), (129, 257, 65), (40000, 602, 602), (77000, 300, 640), (60001, 1000, 90)])
def test_gemm_nt_tcgen05_fp16(m, n, k, cg):
    from ogl_b200 import native
    torch.manual_seed(m + n + k)
    ld = (k + 7) // 8 * 8
    a = torch.full((m, ld), 1000.0, dtype=torch.float16, device="cuda")      # poisoned pad columns
    b = torch.full((n, ld), 1000.0, dtype=torch.float16, device="cuda")
    a[:, :k] = torch.randn(m, k, device="cuda")
    b[:, :k] = torch.randn(n, k, device="cuda")
    bias = torch.randn(n, device="cuda")
    ref = a[:, :k].double() @ b[:, :k].double().t()
    got = native.gemm_f16_nt_ex(a, b, k=k, out_f16=False, cg=cg)           # fp32 output (the logits epilogue)
    _check_tf32(got, ref, k, "NT fp16 %dx%dx%d cg%d" % (m, n, k, cg))
    # activation epilogue: bias + ReLU + fp16 mask (positive / zero / negative / -0 entries), fp16 output through TMA stores
    mask = torch.randn(m, (n + 7) // 8 * 8, device="cuda")
    mask[::3] = 0.0
    mask[1::7] = -0.0
    mask = mask.half()
    got2 = native.gemm_f16_nt_ex(a, b, k=k, out_f16=True, bias=bias, relu=True, mask=mask, cg=cg)
    assert got2.dtype == torch.float16
    ref2 = torch.relu(ref + bias.double()) * (mask[:, :n].float() > 0)
    err = (got2.double() - ref2).abs()
    assert bool((err <= 2 ** -11 * ref2.abs() + 1e-6 * (k ** 0.5 + 8) * ref.abs().max() + 2 ** -25).all()), "fp16 activation epilogue: max err %.3e" % err.max().item()
    # the same without a mask (the forward fc_pool epilogue: bias + ReLU, 16-bit TMA stores)
    got3 = native.gemm_f16_nt_ex(a, b, k=k, out_f16=True, bias=bias, relu=True, cg=cg)
    ref3 = torch.relu(ref + bias.double())
    err = (got3.double() - ref3).abs()
    assert bool((err <= 2 ** -11 * ref3.abs() + 1e-6 * (k ** 0.5 + 8) * ref.abs().max() + 2 ** -25).all()), "fp16 bias + ReLU epilogue: max err %.3e" % err.max().item()


@pytest.mark.parametrize("m,n,k,ws", [(64, 64, 64, 1 << 22), (1000, 602, 602, 1 << 24), (100, 41, 600, 1 << 22), (5000, 600, 41, 1 << 22),
                                      (30000, 602, 602, 1 << 24), (200, 24, 50, 1 << 22), (1000, 602, 602, 0), (1, 8, 8, 0)])
def test_gemm_tn_tcgen05_fp16(m, n, k, ws):
    """the weight-gradient shape with the unscaling factor of the loss-scaled backward pass (alpha = 2^-7: exact)"""
    from ogl_b200 import native
    torch.manual_seed(m + n + k)
    ldn, ldk = (n + 7) // 8 * 8, (k + 7) // 8 * 8
    a = torch.full((m, ldn), 1000.0, dtype=torch.float16, device="cuda")
    b = torch.full((m, ldk), 1000.0, dtype=torch.float16, device="cuda")
    a[:, :n] = torch.randn(m, n, device="cuda")
    b[:, :k] = torch.randn(m, k, device="cuda")
    got = native.gemm_f16_tn(a, b, n=n, k=k, alpha=2.0 ** -7, workspace_elems=ws)
    _check_tf32(got, (a[:, :n].double().t() @ b[:, :k].double()) * 2.0 ** -7, m, "TN fp16 %dx%dx%d" % (m, n, k))
