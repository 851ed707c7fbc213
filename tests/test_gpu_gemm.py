"""tcgen05 GEMM + fused epilogues vs fp32 torch matmul (reference test: apex/contrib/test/fused_dense/test_fused_dense.py, M=1536,
K=1024, N=3072, atol=rtol=1e-3 on its scale) and the modules built on it."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [(128, 256, 64), (1536, 3072, 1024), (200, 328, 136), (8, 8, 8), (1000, 1000, 1000), (4096, 512, 4096), (136, 4104, 72)]


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_layouts(cuda_dev, M, N, K, dtype, a_mn, b_mn):
    from apex_b200.ops import gemm as G
    torch.manual_seed(0)
    A = torch.randn(M, K, device=cuda_dev).to(dtype)
    B = torch.randn(N, K, device=cuda_dev).to(dtype)
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    d = G.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    assert d is not None, "native GEMM refused an aligned problem"
    ref = A.float() @ B.float().t()
    assert _rel(d, ref) < 2e-3, _rel(d, ref)
    assert G.stats["native"] > 0


def test_gemm_epilogues(cuda_dev):
    from apex_b200.ops import gemm as G
    torch.manual_seed(0)
    M, N, K = 520, 776, 264
    x = torch.randn(M, K, device=cuda_dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=cuda_dev, dtype=torch.bfloat16) * 0.05
    bias = torch.randn(N, device=cuda_dev, dtype=torch.bfloat16)
    ref = x.float() @ w.float().t() + bias.float()
    y = G.gemm(x, w, epi=G.EPI_BIAS, bias=bias)
    assert _rel(y, ref) < 1e-2
    aux = torch.empty(M, N, device=cuda_dev, dtype=torch.bfloat16)
    h = G.gemm(x, w, epi=G.EPI_BIAS_GELU, bias=bias, aux=aux)
    assert _rel(aux, ref) < 1e-2 and _rel(h, F.gelu(aux.float())) < 1e-2
    r = G.gemm(x, w, epi=G.EPI_BIAS_RELU, bias=bias)
    assert _rel(r, torch.relu(ref)) < 1e-2
    s = G.gemm(x, w, epi=G.EPI_BIAS_SIGMOID, bias=bias)
    assert _rel(s, torch.sigmoid(ref)) < 1e-2
    # dgelu: (dy @ W2) * gelu'(aux)
    dy = torch.randn(M, 192, device=cuda_dev, dtype=torch.bfloat16)
    w2 = torch.randn(192, N, device=cuda_dev, dtype=torch.bfloat16) * 0.05
    a = aux.float().requires_grad_(True)
    (F.gelu(a) * (dy.float() @ w2.float())).sum().backward()
    dg = G.gemm(dy, w2, b_mn=True, epi=G.EPI_DGELU, aux=aux)
    assert _rel(dg, a.grad) < 1e-2
    # beta=1 accumulate into fp32 and bf16 main grads
    main = torch.randn(N, K, device=cuda_dev)
    dyy = torch.randn(M, N, device=cuda_dev, dtype=torch.bfloat16)
    exp = main + dyy.float().t() @ x.float()
    from apex_b200.transformer.functional import wgrad_gemm_accum_fp16, wgrad_gemm_accum_fp32
    wgrad_gemm_accum_fp32(x, dyy, main)
    assert _rel(main, exp) < 2e-3
    main16 = torch.randn(N, K, device=cuda_dev, dtype=torch.bfloat16)
    exp16 = main16.float() + dyy.float().t() @ x.float()
    wgrad_gemm_accum_fp16(x.view(4, 130, K), dyy.view(4, 130, N), main16)
    assert _rel(main16, exp16) < 1e-2
    cs = G.colsum(dyy)
    assert _rel(cs, dyy.float().sum(0)) < 1e-2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_fused_dense_modules(cuda_dev, dtype):
    from apex_b200.fused_dense import FusedDense, FusedDenseGeluDense
    torch.manual_seed(0)
    tol = 2e-2 if dtype != torch.float32 else 1e-4
    fd = FusedDense(1024, 3072).to(cuda_dev, dtype)
    ref = torch.nn.Linear(1024, 3072).to(cuda_dev)
    with torch.no_grad():
        ref.weight.copy_(fd.weight.float())
        ref.bias.copy_(fd.bias.float())
    x = torch.randn(3, 512, 1024, device=cuda_dev, dtype=dtype, requires_grad=True)
    xr = x.detach().float().requires_grad_(True)
    dy = torch.randn(3, 512, 3072, device=cuda_dev, dtype=dtype)
    fd(x).backward(dy)
    ref(xr).backward(dy.float())
    assert _rel(x.grad, xr.grad) < tol and _rel(fd.weight.grad, ref.weight.grad) < tol and _rel(fd.bias.grad, ref.bias.grad) < tol
    g = FusedDenseGeluDense(1024, 4096, 1024).to(cuda_dev, dtype)
    x2 = torch.randn(1536, 1024, device=cuda_dev, dtype=dtype, requires_grad=True)
    x2r = x2.detach().float().requires_grad_(True)
    ps = {n: p.detach().float().requires_grad_(True) for n, p in g.named_parameters()}
    out_ref = F.linear(F.gelu(F.linear(x2r, ps["weight1"], ps["bias1"])), ps["weight2"], ps["bias2"])
    out = g(x2)
    assert _rel(out, out_ref) < tol
    dy2 = torch.randn_like(out)
    out.backward(dy2)
    out_ref.backward(dy2.float())
    assert _rel(x2.grad, x2r.grad) < tol
    for n, p in g.named_parameters():
        assert _rel(p.grad, ps[n].grad) < tol, n


@pytest.mark.parametrize("activation", ["none", "relu", "sigmoid"])
def test_mlp_module(cuda_dev, activation):
    from apex_b200.mlp import MLP
    torch.manual_seed(0)
    sizes = [480, 1024, 1024, 512, 256, 8]
    m = MLP(sizes, activation=activation).to(cuda_dev, torch.bfloat16)
    layers = []
    for i in range(len(sizes) - 1):
        lin = torch.nn.Linear(sizes[i], sizes[i + 1]).to(cuda_dev)
        with torch.no_grad():
            lin.weight.copy_(m.weights[i].float())
            lin.bias.copy_(m.biases[i].float())
        layers.append(lin)
        if activation == "relu":
            layers.append(torch.nn.ReLU())
        elif activation == "sigmoid":
            layers.append(torch.nn.Sigmoid())
    ref = torch.nn.Sequential(*layers)
    x = torch.randn(1024, 480, device=cuda_dev, dtype=torch.bfloat16, requires_grad=True)
    xr = x.detach().float().requires_grad_(True)
    y, yr = m(x), ref(xr)
    assert _rel(y, yr) < 3e-2
    y.float().mean().backward()
    yr.mean().backward()
    # five bf16 layers vs an fp32 oracle: ReLU masks flip on near-zero pre-activations, so the bound is loose
    assert _rel(x.grad, xr.grad) < 0.15
    assert _rel(m.weights[0].grad, layers[0].weight.grad) < 0.15


@pytest.mark.parametrize("M,N,K", [(256, 512, 256), (300, 264, 1040), (128, 128, 64), (1024, 2048, 4096)])
@pytest.mark.parametrize("fmt", ["e4m3", "e5m2"])
def test_gemm_fp8(cuda_dev, M, N, K, fmt):
    """tcgen05.mma.kind::f8f6f4 GEMM vs the fp32 product of the dequantised operands (exact up to fp32 accumulation order)."""
    from apex_b200.ops import gemm as G
    dt = torch.float8_e4m3fn if fmt == "e4m3" else torch.float8_e5m2
    torch.manual_seed(0)
    a = torch.randn(M, K, device=cuda_dev)
    b = torch.randn(N, K, device=cuda_dev)
    a8, sa = G.quantize_fp8(a, dt)
    b8, sb = G.quantize_fp8(b, dt)
    alpha = float((sa * sb).item())
    ref = (a8.float() @ b8.float().t()) * alpha
    out = G.gemm_fp8(a8, b8, alpha, out_dtype=torch.float32)
    assert out is not None
    torch.testing.assert_close(out, ref, atol=2e-3 * K ** 0.5, rtol=1e-3)
    bias = torch.randn(N, device=cuda_dev, dtype=torch.bfloat16)
    out2 = G.gemm_fp8(a8, b8, alpha, out_dtype=torch.bfloat16, epi=G.EPI_BIAS, bias=bias)
    torch.testing.assert_close(out2.float(), ref + bias.float(), atol=0.05 * K ** 0.5 / 8 + 0.05, rtol=2e-2)


def test_fused_dense_fp8_close_to_bf16(cuda_dev):
    from apex_b200.fused_dense import fused_dense_fp8_function
    torch.manual_seed(0)
    x = torch.randn(512, 1024, device=cuda_dev, dtype=torch.bfloat16, requires_grad=True)
    w = (torch.randn(768, 1024, device=cuda_dev) * 0.03).to(torch.bfloat16).requires_grad_(True)
    b = torch.randn(768, device=cuda_dev, dtype=torch.bfloat16, requires_grad=True)
    y = fused_dense_fp8_function(x, w, b)
    ref = torch.nn.functional.linear(x.float(), w.float(), b.float())
    rel = (y.float() - ref).norm() / ref.norm()
    assert rel < 0.06, rel  # two e4m3 operands: ~2^-4 relative element noise averaged over K
    y.float().pow(2).mean().backward()
    assert torch.isfinite(x.grad.float()).all() and torch.isfinite(w.grad.float()).all() and b.grad is not None


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_fp8_quantize_kernel(cuda_dev, dtype):
    """csrc/fp8_quant.cu vs the torch formulation; scales stay on the device and feed the GEMM epilogue."""
    from apex_b200.ops import gemm as G
    torch.manual_seed(0)
    x = (torch.randn(1000, 1040, device=cuda_dev) * 3).to(dtype)
    q, inv = G.quantize_fp8(x)
    amax = x.float().abs().max()
    torch.testing.assert_close(inv, (amax / 448.0).reshape(1), rtol=1e-6, atol=0)
    ref = (x.float() * (448.0 / amax)).to(torch.float8_e4m3fn)
    assert (q.float() - ref.float()).abs().max() <= 32.0 and (q.view(torch.uint8) != ref.view(torch.uint8)).float().mean() < 0.02
    w = torch.randn(264, 1040, device=cuda_dev).to(dtype)
    w8, winv = G.quantize_fp8(w)
    out = G.gemm_fp8(q, w8, 1.0, scale_a=inv, scale_b=winv, out_dtype=torch.float32)
    assert out is not None
    want = (q.float() @ w8.float().t()) * inv * winv
    torch.testing.assert_close(out, want, atol=2e-2, rtol=1e-3)


@pytest.mark.parametrize("M,N,K", [(512, 1024, 256), (300, 200, 136), (8192, 4096, 1024)])
def test_dgrad_epilogue_accumulates_the_bias_gradient(cuda_dev, M, N, K):
    """EPI_DGELU / EPI_NONE + colsum: the column sums of the stored tile come out of the epilogue registers (fused BGRAD)."""
    from apex_b200.ops import gemm as G
    torch.manual_seed(0)
    dy = torch.randn(M, K, device=cuda_dev, dtype=torch.bfloat16)
    w = (torch.randn(K, N, device=cuda_dev) * 0.05).to(torch.bfloat16)
    aux = torch.randn(M, N, device=cuda_dev, dtype=torch.bfloat16)
    for use_aux in (False, True):
        dx, cs = G.linear_dgrad(dy, w, dgelu_aux=aux if use_aux else None, want_colsum=True)
        ref = dy.float() @ w.float()
        if use_aux:
            a = aux.float()
            ref = ref * (0.5 * (1 + torch.erf(a * 0.7071067811865476)) + a * torch.exp(-0.5 * a * a) * 0.3989422804014327)
        torch.testing.assert_close(dx.float(), ref, atol=3e-2, rtol=3e-2)
        want = ref.sum(0)
        err = (cs.float() - want).abs().max().item()
        assert err <= 2e-2 * max(1.0, want.abs().max().item()), (use_aux, err)


@pytest.mark.parametrize("R,C", [(256, 512), (130, 70), (64, 4096)])
@pytest.mark.parametrize("fmt", ["e4m3", "e5m2"])
def test_fp8_quantize_dual_matches_the_plain_kernel_and_its_transpose(cuda_dev, R, C, fmt):
    from apex_b200.ops import gemm as G
    dt = torch.float8_e4m3fn if fmt == "e4m3" else torch.float8_e5m2
    torch.manual_seed(0)
    x = (torch.randn(R, C, device=cuda_dev) * 3).bfloat16()
    q0, s0 = G.quantize_fp8(x, dt)
    q, qt, s = G.quantize_fp8_dual(x, dt)
    torch.testing.assert_close(s, s0)
    assert torch.equal(q.view(torch.uint8), q0.view(torch.uint8))
    assert torch.equal(qt.view(torch.uint8), q0.view(torch.uint8).t().contiguous())


def test_mixed_format_fp8_gemm_e5m2_times_e4m3(cuda_dev):
    from apex_b200.ops import gemm as G
    torch.manual_seed(0)
    a = torch.randn(256, 512, device=cuda_dev).bfloat16()
    b = torch.randn(384, 512, device=cuda_dev).bfloat16()
    a8, sa = G.quantize_fp8(a, torch.float8_e5m2)
    b8, sb = G.quantize_fp8(b, torch.float8_e4m3fn)
    out = G.gemm_fp8(a8, b8, 1.0, scale_a=sa, scale_b=sb, out_dtype=torch.float32)
    ref = (a8.float() * sa) @ (b8.float() * sb).t()
    torch.testing.assert_close(out, ref, rtol=1e-3, atol=1e-2)      # exact products of the dequantised operands, fp32 accumulation order aside


def test_fused_dense_fp8_backward_close_to_bf16(cuda_dev):
    """fp8_backward=True: dgrad (E5M2 dy x E4M3 W^T) and wgrad (E5M2 dy^T x E4M3 x^T) on the fp8 GEMM vs the 16-bit backward."""
    from apex_b200.fused_dense import fused_dense_fp8_function, fused_dense_function
    torch.manual_seed(0)
    x = torch.randn(512, 1024, device=cuda_dev).bfloat16().requires_grad_()
    w = (torch.randn(768, 1024, device=cuda_dev) * 0.03).bfloat16().requires_grad_()
    b = torch.randn(768, device=cuda_dev).bfloat16().requires_grad_()
    dy = torch.randn(512, 768, device=cuda_dev).bfloat16()
    y8 = fused_dense_fp8_function(x, w, b, fp8_backward=True)
    g8 = torch.autograd.grad(y8, (x, w, b), dy)
    y16 = fused_dense_function(x, w, b)
    g16 = torch.autograd.grad(y16, (x, w, b), dy)
    for a_, r_, name in zip(g8, g16, ("dx", "dw", "db")):
        err = (a_.float() - r_.float()).norm() / r_.float().norm()
        assert err < (0.08 if name != "db" else 1e-2), (name, float(err))


def test_ffn_block_fp8_backward_close_to_bf16(cuda_dev):
    from apex_b200.fused_dense import fused_dense_gelu_dense_fp8_function, fused_dense_gelu_dense_function
    torch.manual_seed(0)
    x = torch.randn(256, 512, device=cuda_dev).bfloat16().requires_grad_()
    w1 = (torch.randn(1024, 512, device=cuda_dev) * 0.04).bfloat16().requires_grad_()
    b1 = (torch.randn(1024, device=cuda_dev) * 0.1).bfloat16().requires_grad_()
    w2 = (torch.randn(512, 1024, device=cuda_dev) * 0.03).bfloat16().requires_grad_()
    b2 = (torch.randn(512, device=cuda_dev) * 0.1).bfloat16().requires_grad_()
    dy = torch.randn(256, 512, device=cuda_dev).bfloat16()
    y8 = fused_dense_gelu_dense_fp8_function(x, w1, b1, w2, b2, fp8_backward=True)
    g8 = torch.autograd.grad(y8, (x, w1, b1, w2, b2), dy)
    y16 = fused_dense_gelu_dense_function(x, w1, b1, w2, b2)
    g16 = torch.autograd.grad(y16, (x, w1, b1, w2, b2), dy)
    assert (y8.float() - y16.float()).norm() / y16.float().norm() < 0.08
    for a_, r_, name in zip(g8, g16, ("dx", "dw1", "db1", "dw2", "db2")):
        err = (a_.float() - r_.float()).norm() / r_.float().norm()
        assert err < 0.12, (name, float(err))


# ---- fp32 operands as TF32 (tcgen05.mma.kind::tf32) -------------------------------------------------------------------------
@pytest.fixture
def allow_tf32():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    yield
    torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1536, 3072, 1024), (200, 328, 136), (4, 4, 4), (1000, 1000, 1000), (260, 516, 36)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_tf32_layouts(cuda_dev, allow_tf32, M, N, K, a_mn, b_mn):
    from apex_b200.ops import gemm as G
    torch.manual_seed(0)
    A = torch.randn(M, K, device=cuda_dev)
    B = torch.randn(N, K, device=cuda_dev)
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    before = G.stats.get("tf32", 0)
    d = G.gemm(a, b, a_mn=a_mn, b_mn=b_mn)
    assert d is not None and d.dtype == torch.float32, "tf32 GEMM refused an aligned fp32 problem"
    assert G.stats.get("tf32", 0) == before + 1
    ref = (A.double() @ B.double().t()).float()
    assert _rel(d, ref) < 1.5e-3, _rel(d, ref)      # 10-bit mantissa products (truncated operands), fp32 accumulation
    # the operands rounded to TF32 by hand must reproduce the kernel far more closely: the only error left is accumulation order
    trunc = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    rna = lambda t: ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    errs = [_rel(d, (f(A).double() @ f(B).double().t()).float()) for f in (trunc, rna)]
    assert min(errs) < 1e-5, errs


def test_gemm_tf32_epilogues(cuda_dev, allow_tf32):
    from apex_b200.ops import gemm as G
    torch.manual_seed(1)
    M, N, K = 520, 392, 264
    x, w, bias = torch.randn(M, K, device=cuda_dev), torch.randn(N, K, device=cuda_dev) * 0.05, torch.randn(N, device=cuda_dev)
    y = G.linear_fwd(x, w, bias)
    assert _rel(y, x @ w.t() + bias) < 2e-3
    aux = torch.empty(M, N, device=cuda_dev)
    h = G.linear_fwd(x, w, bias, epi=G.EPI_BIAS_GELU, aux=aux)
    pre = x.double() @ w.double().t() + bias.double()
    assert _rel(aux, pre.float()) < 2e-3 and _rel(h, F.gelu(pre).float()) < 2e-3
    dy = torch.randn(M, N, device=cuda_dev)
    dx, db = G.linear_dgrad(dy, w, want_colsum=True)
    assert _rel(dx, dy @ w) < 2e-3 and _rel(db, (dy @ w).sum(0)) < 2e-3
    acc = torch.randn(N, K, device=cuda_dev)
    ref = acc.double() + dy.double().t() @ x.double()
    G.linear_wgrad(dy, x, accum_into=acc)
    assert _rel(acc, ref.float()) < 2e-3


def test_fused_dense_fp32_runs_on_the_tf32_kernel_when_allowed(cuda_dev, allow_tf32):
    from apex_b200.fused_dense import FusedDenseGeluDense
    from apex_b200.ops import gemm as G
    torch.manual_seed(0)
    g = FusedDenseGeluDense(512, 1024, 256).to(cuda_dev)
    x = torch.randn(384, 512, device=cuda_dev, requires_grad=True)
    xr = x.detach().double().requires_grad_(True)
    ps = {n: p.detach().double().requires_grad_(True) for n, p in g.named_parameters()}
    fb, t0 = G.stats["fallback"], G.stats.get("tf32", 0)
    out = g(x)
    dy = torch.randn_like(out)
    out.backward(dy)
    assert G.stats["fallback"] == fb and G.stats.get("tf32", 0) >= t0 + 6      # 2 forward, 2 dgrad, 2 wgrad GEMMs, no library call
    out_ref = F.linear(F.gelu(F.linear(xr, ps["weight1"], ps["bias1"])), ps["weight2"], ps["bias2"])
    out_ref.backward(dy.double())
    assert _rel(out, out_ref) < 3e-3 and _rel(x.grad, xr.grad) < 3e-3
    for n, p in g.named_parameters():
        assert _rel(p.grad, ps[n].grad) < 3e-3, n


def test_fp32_without_tf32_is_an_announced_library_gemm(cuda_dev):
    from apex_b200.ops import gemm as G
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        fb = G.stats["fallback"]
        x, w = torch.randn(64, 64, device=cuda_dev), torch.randn(64, 64, device=cuda_dev)
        y = G.linear_fwd(x, w)
        assert G.stats["fallback"] == fb + 1
        torch.testing.assert_close(y, x @ w.t())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
