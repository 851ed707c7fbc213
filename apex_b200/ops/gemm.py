"""tcgen05 / TMEM / TMA GEMM front-end (csrc/gemm_sm100.cu): ``D = op(A) @ op(B)`` with fused epilogues.

``linear_*`` helpers express the three GEMMs of a linear layer (forward, dgrad, wgrad) without materialising transposes:
the kernel reads either operand K-major or MN-major straight from its row-major storage.
fp32 operands run on the same kernel as TF32 (tcgen05.mma.kind::tf32, fp32 accumulation) when the caller allows TF32 products
(``torch.backends.cuda.matmul.allow_tf32`` / ``torch.set_float32_matmul_precision("high")``, or APEX_B200_TF32=1); with IEEE fp32
products requested (torch's default) there is no tensor-core instruction to use and the GEMM is a library SGEMM, like the reference's
(csrc/fused_dense_cuda.cu:29-30, CUBLAS_COMPUTE_32F). fp64 and extents the kernel does not take (inner extents that are not multiples of
16 bytes) also go to cuBLAS via torch; every such call is counted in ``stats["fallback"]`` and warned about once per reason.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..parallel import param_sync as _param_sync

_lib.declare("ab_gemm_bf16", "p p p i i i l l l i i i i i p p l p l p p i i l l i p")
_lib.declare("ab_gemm_tf32", "p p p i i i l l l i i i p p l p l p i p")
_lib.declare("ab_colsum", "p p p i i l i p")
_lib.declare("ab_gemm_fp8", "p p p i i i l l l i i i i p p l f p p i p")
_lib.declare("ab_fp8_quantize_dual", "p p p i i p p i i p")
_lib.declare("ab_fp8_quantize", "p p l p p i i p")

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, EPI_ACCUM, EPI_BIAS_RELU, EPI_BIAS_SIGMOID, EPI_RELU, EPI_SIGMOID = range(9)

_ws: dict = {}
stats = {"native": 0, "fallback": 0, "guarded": 0}


_warned: set = set()


def _note_fallback(why: str) -> None:
    """The tcgen05 kernel takes bf16 / fp16 operands with 8-element-aligned extents; everything else is a plain library GEMM (cuBLAS via
    torch), as the module docstring says. Counted in ``stats`` and reported once per reason so that it is never silent."""
    stats["fallback"] += 1
    if why not in _warned:
        _warned.add(why)
        import warnings

        warnings.warn(f"apex_b200.ops.gemm: library (cuBLAS) GEMM used instead of the tcgen05 kernel: {why}", stacklevel=3)


def tf32_allowed() -> bool:
    """fp32 operands may be multiplied as TF32: torch's own switch, overridable with APEX_B200_TF32=0/1."""
    import os

    env = os.environ.get("APEX_B200_TF32")
    if env is not None:
        return env not in ("0", "", "false", "False")
    return bool(torch.backends.cuda.matmul.allow_tf32)


def _native_ok(a: torch.Tensor, b: torch.Tensor) -> bool:
    if not (a.is_cuda and a.dtype == b.dtype and _lib.available() and a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1):
        return False
    return a.dtype in (torch.bfloat16, torch.float16) or (a.dtype == torch.float32 and tf32_allowed())


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, out_dtype=None, epi: int = EPI_NONE,
         bias: torch.Tensor | None = None, aux: torch.Tensor | None = None, c: torch.Tensor | None = None,
         out: torch.Tensor | None = None, sms: int = 0, colsum_out: torch.Tensor | None = None) -> torch.Tensor | None:
    """a: [M,K] (a_mn=False) or [K,M] (a_mn=True), row-major with unit inner stride; b: [N,K] or [K,N] likewise.
    ``colsum_out``: fp32 [N], zero on entry — the column sums of D (a bias gradient) are accumulated into it by the epilogue.
    Returns D [M,N], or None if the native kernel cannot take this problem (caller falls back)."""
    if not _native_ok(a, b):
        if _param_sync._regions:   # the caller's library fallback reads the operands as a whole
            _param_sync.wait(a)
            _param_sync.wait(b)
        if a.is_cuda:
            if a.dtype == torch.float32 and b.dtype == torch.float32 and not tf32_allowed():
                _note_fallback("fp32 operands with IEEE fp32 products requested (allow_tf32 is off): library SGEMM")
            else:
                _note_fallback(f"{a.dtype} operands" if a.dtype not in (torch.bfloat16, torch.float16, torch.float32)
                               else "non-unit inner stride / mixed dtypes")
        return None
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    assert K == Kb, f"inner dimensions differ: {K} vs {Kb}"
    out_dtype = out_dtype or a.dtype
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    if c is not None:
        assert c.dtype == out.dtype and c.stride(1) == 1
    if bias is not None and bias.dtype != out.dtype:
        bias = bias.to(out.dtype)
    # B inside a parameter buffer whose all-gather is still in flight (overlap_param_sync): guard it tile by tile inside the kernel when
    # the layout allows (forward: K-major dense weight), else make the stream wait for the buckets under it
    guard = (None, 0, 0, 0, 0)
    if _param_sync._regions:
        hit = _param_sync.lookup(b)
        if hit is not None:
            region, g = hit
            if (not b_mn) and b.stride(0) == K and b.is_contiguous():
                guard = g
                stats["guarded"] += 1
            else:
                region.wait(b)
        if _param_sync.find(a) is not None:
            _param_sync.wait(a)
    if a.dtype == torch.float32:
        return _gemm_tf32(a, b, a_mn, b_mn, M, N, K, out, epi, bias, aux, c, colsum_out, sms)
    try:
        rc_ok = True
        _lib.fn("ab_gemm_bf16")(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), b.stride(0), out.stride(0), int(a_mn),
                                int(b_mn), _lib.dt(a), _lib.dt(out), int(epi), _lib.ptr(bias), _lib.ptr(aux),
                                aux.stride(0) if aux is not None else 0, _lib.ptr(c), c.stride(0) if c is not None else 0, _lib.ptr(colsum_out), guard[0], int(guard[1]), int(guard[2]), int(guard[3]), int(guard[4]), int(sms),
                                _lib.stream_ptr(a.device))
    except RuntimeError as e:
        if "bad argument (-10)" in str(e):
            rc_ok = False
        else:
            raise
    if not rc_ok:
        _note_fallback(f"extents / leading dimensions not multiples of 8 (M={M}, N={N}, K={K})")
        return None
    stats["native"] += 1
    return out


def _gemm_tf32(a, b, a_mn, b_mn, M, N, K, out, epi, bias, aux, c, colsum_out, sms):
    """fp32 operands on tcgen05.mma.kind::tf32; fp32 output / bias / aux / accumulate source."""
    if out.dtype != torch.float32 or (aux is not None and aux.dtype != torch.float32):
        _note_fallback("tf32 GEMM with a non-fp32 output")
        return None
    if _param_sync._regions:
        _param_sync.wait(a)
        _param_sync.wait(b)
    try:
        _lib.fn("ab_gemm_tf32")(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), b.stride(0), out.stride(0), int(a_mn), int(b_mn),
                                int(epi), _lib.ptr(bias), _lib.ptr(aux), aux.stride(0) if aux is not None else 0, _lib.ptr(c),
                                c.stride(0) if c is not None else 0, _lib.ptr(colsum_out), int(sms), _lib.stream_ptr(a.device))
    except RuntimeError as e:
        if "bad argument (-10)" not in str(e):
            raise
        _note_fallback(f"fp32 extents / leading dimensions not multiples of 4 (M={M}, N={N}, K={K})")
        return None
    stats["native"] += 1
    stats["tf32"] = stats.get("tf32", 0) + 1
    return out


def colsum(x: torch.Tensor) -> torch.Tensor:
    """Column sums of a 2-D row-major matrix (bias gradients), deterministic, fp32 accumulation."""
    if not (x.is_cuda and _lib.available() and x.dtype in (torch.float32, torch.float16, torch.bfloat16) and x.stride(1) == 1):
        return x.float().sum(0).to(x.dtype)
    M, N = x.shape
    key = x.device
    ws = _ws.get(key)
    if ws is None or ws.numel() < 64 * N:
        ws = _ws[key] = torch.empty(64 * N, dtype=torch.float32, device=x.device)
    out = torch.empty(N, dtype=x.dtype, device=x.device)
    _lib.fn("ab_colsum")(x.data_ptr(), out.data_ptr(), ws.data_ptr(), M, N, x.stride(0), _lib.dt(x), _lib.stream_ptr(x.device))
    return out


# ---- the GEMMs of y = x W^T + b  (x [M,in], W [out,in]) ----------------------------------------------------------------
def linear_fwd(x, w, bias=None, epi=None, aux=None, out_dtype=None):
    epi = (EPI_BIAS if bias is not None else EPI_NONE) if epi is None else epi
    y = gemm(x, w, epi=epi, bias=bias, aux=aux, out_dtype=out_dtype)
    if y is not None:
        return y
    y = torch.matmul(x, w.t())
    if bias is not None:
        y = y + bias
    if epi == EPI_BIAS_GELU:
        aux.copy_(y)
        y = torch.nn.functional.gelu(y)
    elif epi in (EPI_BIAS_RELU, EPI_RELU):
        y = torch.relu(y)
    elif epi in (EPI_BIAS_SIGMOID, EPI_SIGMOID):
        y = torch.sigmoid(y)
    return y


def linear_dgrad(dy, w, dgelu_aux=None, want_colsum: bool = False):
    """dx = dy @ W  (optionally fused with * gelu'(aux)). ``want_colsum``: also return the column sums of dx (the bias gradient of the
    layer below), accumulated by the same kernel's epilogue — no second pass over dx."""
    cs = torch.zeros(w.shape[1], dtype=torch.float32, device=dy.device) if (want_colsum and _native_ok(dy, w)) else None
    dx = gemm(dy, w, b_mn=True, epi=EPI_DGELU if dgelu_aux is not None else EPI_NONE, aux=dgelu_aux, colsum_out=cs)
    if dx is not None:
        return (dx, cs.to(dx.dtype)) if want_colsum else dx
    dx = torch.matmul(dy, w)
    if dgelu_aux is not None:
        a = dgelu_aux.float()
        cdf = 0.5 * (1 + torch.erf(a * 0.7071067811865476))
        pdf = torch.exp(-0.5 * a * a) * 0.3989422804014327
        dx = (dx.float() * (cdf + a * pdf)).to(dy.dtype)
    return (dx, colsum(dx)) if want_colsum else dx


def linear_wgrad(dy, x, accum_into: torch.Tensor | None = None, out_dtype=None):
    """dW = dy^T @ x; with ``accum_into`` the result is ADDED to that tensor (beta = 1 main-grad accumulation)."""
    if accum_into is not None:
        r = gemm(dy, x, a_mn=True, b_mn=True, epi=EPI_ACCUM, c=accum_into, out=accum_into, out_dtype=accum_into.dtype)
        if r is None:
            accum_into.add_(torch.matmul(dy.t().to(x.dtype), x).to(accum_into.dtype))
        return accum_into
    dw = gemm(dy, x, a_mn=True, b_mn=True, out_dtype=out_dtype)
    if dw is None:
        dw = torch.matmul(dy.t(), x)
        if out_dtype is not None:
            dw = dw.to(out_dtype)
    return dw


# ---- fp8 (E4M3 / E5M2) operands: tcgen05.mma.kind::f8f6f4, fp32 accumulation, per-tensor scales folded into the epilogue --------
_F8 = tuple(getattr(torch, n) for n in ("float8_e4m3fn", "float8_e5m2") if hasattr(torch, n))


_q_scratch: dict = {}


def quantize_fp8(x: torch.Tensor, dtype=None):
    """Per-tensor dynamic scaling: returns (x_fp8, inv_scale) with x ~= x_fp8.float() * inv_scale (amax mapped to the format's max).
    On CUDA this is two kernels (csrc/fp8_quant.cu: amax, scale + cast) and ``inv_scale`` is a 1-element device tensor: no host sync."""
    dtype = dtype or torch.float8_e4m3fn
    if x.is_cuda and _lib.available() and x.dtype in (torch.float32, torch.float16, torch.bfloat16):
        xc = x.detach().contiguous()
        q = torch.empty(xc.shape, dtype=dtype, device=x.device)
        inv = torch.empty(1, dtype=torch.float32, device=x.device)
        scr = _q_scratch.get(x.device)
        if scr is None:
            scr = _q_scratch[x.device] = torch.zeros(1, dtype=torch.int32, device=x.device)
        _lib.fn("ab_fp8_quantize")(xc.data_ptr(), q.data_ptr(), xc.numel(), scr.data_ptr(), inv.data_ptr(), _lib.dt(xc), _lib.dt(dtype),
                                   _lib.stream_ptr(x.device))
        return q, inv
    amax = x.detach().abs().amax().float().clamp_min(1e-12)
    scale = torch.finfo(dtype).max / amax
    return (x.float() * scale).to(dtype), (1.0 / scale).reshape(1)


def gemm_fp8(a8: torch.Tensor, b8: torch.Tensor, alpha: float = 1.0, *, scale_a: torch.Tensor | None = None,
             scale_b: torch.Tensor | None = None, out_dtype=torch.bfloat16, epi: int = EPI_NONE, bias: torch.Tensor | None = None,
             aux: torch.Tensor | None = None, out: torch.Tensor | None = None):
    """D [M, N] = alpha * scale_a * scale_b * a8 [M, K] @ b8 [N, K]^T for 8-bit float operands (both K-major; E4M3 / E5M2 may be mixed);
    ``scale_a`` / ``scale_b`` are optional 1-element fp32 DEVICE tensors (the dequantisation scales of :func:`quantize_fp8`).
    Returns None when the native kernel cannot take the problem (K or a leading dimension not a multiple of 16, CPU tensors)."""
    if not (a8.is_cuda and _lib.available() and a8.dtype in _F8 and b8.dtype in _F8 and a8.dim() == 2 and b8.dim() == 2
            and a8.stride(1) == 1 and b8.stride(1) == 1):
        return None
    M, K = a8.shape
    N = b8.shape[0]
    assert b8.shape[1] == K
    if K % 16 or a8.stride(0) % 16 or b8.stride(0) % 16:
        return None
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a8.device)
    if bias is not None and bias.dtype != out.dtype:
        bias = bias.to(out.dtype)
    _lib.fn("ab_gemm_fp8")(a8.data_ptr(), b8.data_ptr(), out.data_ptr(), M, N, K, a8.stride(0), b8.stride(0), out.stride(0), _lib.dt(a8),
                           _lib.dt(b8), _lib.dt(out), int(epi), _lib.ptr(bias), _lib.ptr(aux), aux.stride(0) if aux is not None else 0, float(alpha),
                           _lib.ptr(scale_a), _lib.ptr(scale_b), 0, _lib.stream_ptr(a8.device))
    stats["native"] += 1
    return out


def quantize_fp8_dual(x: torch.Tensor, dtype=None, want_q: bool = True, want_t: bool = True):
    """One read of a 2-D tensor -> (q [R, C] or None, qt [C, R] or None, inv_scale): the fp8 copy and / or its TRANSPOSE with one shared
    per-tensor scale (csrc/fp8_quant.cu quant_dual_kernel). The transposed copies are what the fp8 backward GEMMs consume: 8-bit tcgen05
    operands are K-major only, dgrad reduces over the output features and wgrad over the tokens."""
    dtype = dtype or torch.float8_e4m3fn
    R, C = x.shape
    if x.is_cuda and _lib.available() and x.dtype in (torch.float32, torch.float16, torch.bfloat16):
        xc = x.detach().contiguous()
        q = torch.empty(R, C, dtype=dtype, device=x.device) if want_q else None
        qt = torch.empty(C, R, dtype=dtype, device=x.device) if want_t else None
        inv = torch.empty(1, dtype=torch.float32, device=x.device)
        scr = _q_scratch.get(x.device)
        if scr is None:
            scr = _q_scratch[x.device] = torch.zeros(1, dtype=torch.int32, device=x.device)
        _lib.fn("ab_fp8_quantize_dual")(xc.data_ptr(), _lib.ptr(q), _lib.ptr(qt), R, C, scr.data_ptr(), inv.data_ptr(), _lib.dt(xc),
                                        _lib.dt(dtype), _lib.stream_ptr(x.device))
        return q, qt, inv
    amax = x.detach().abs().amax().float().clamp_min(1e-12)
    scale = torch.finfo(dtype).max / amax
    q = (x.float() * scale).to(dtype)
    return (q if want_q else None), (q.t().contiguous() if want_t else None), (1.0 / scale).reshape(1)


_wq_cache: dict = {}
_wqt_cache: dict = {}


def _quantize_weight_t_cached(w: torch.Tensor):
    """Transposed E4M3 copy of a weight [N, K] -> [K, N] for the fp8 dgrad, cached per (storage, version) like the forward copy."""
    key = (w.data_ptr(), tuple(w.shape))
    hit = _wqt_cache.get(key)
    if hit is not None and hit[0] == w._version:
        return hit[1], hit[2]
    _, wt8, sw = quantize_fp8_dual(w, want_q=False, want_t=True)
    if len(_wqt_cache) > 256:
        _wqt_cache.clear()
    _wqt_cache[key] = (w._version, wt8, sw)
    return wt8, sw


def linear_bwd_fp8(dy: torch.Tensor, w: torch.Tensor, xt8: torch.Tensor, sx: torch.Tensor, need_dx: bool = True, out_dtype=None):
    """fp8 backward of y = x W^T: the gradient is quantised ONCE to E5M2 (plain + transposed copy, one read), then
    dx = dy8 [M, N] @ Wt8 [K, N]^T  and  dW = dyt8 [N, M] @ xt8 [K, M]^T on the kind::f8f6f4 GEMM (E5M2 x E4M3 operands, fp32 accumulation in
    TMEM, per-tensor dequantisation scales applied in the epilogue). Returns (dx or None, dw) or None when the shapes do not fit
    (every reduction / leading dimension must be a multiple of 16)."""
    M, N = dy.shape
    K = w.shape[1]
    if not (dy.is_cuda and _lib.available() and _F8 and M % 16 == 0 and N % 16 == 0 and K % 16 == 0):
        return None
    out_dtype = out_dtype or dy.dtype
    dy8, dyt8, sdy = quantize_fp8_dual(dy, torch.float8_e5m2, want_q=need_dx, want_t=True)
    dx = None
    if need_dx:
        wt8, sw = _quantize_weight_t_cached(w)
        dx = gemm_fp8(dy8, wt8, 1.0, scale_a=sdy, scale_b=sw, out_dtype=out_dtype)
        if dx is None:
            return None
    dw = gemm_fp8(dyt8, xt8, 1.0, scale_a=sdy, scale_b=sx, out_dtype=w.dtype)
    if dw is None:
        return None
    return dx, dw


def _quantize_weight_cached(w: torch.Tensor):
    """Weights change once per optimizer step, not once per call: key the quantised copy on (storage, version)."""
    key = (w.data_ptr(), tuple(w.shape))
    hit = _wq_cache.get(key)
    if hit is not None and hit[0] == w._version:
        return hit[1], hit[2]
    w8, sw = quantize_fp8(w)
    if len(_wq_cache) > 256:
        _wq_cache.clear()
    _wq_cache[key] = (w._version, w8, sw)
    return w8, sw


def linear_fwd_fp8(x: torch.Tensor, w: torch.Tensor, bias=None, epi=None, aux=None, out_dtype=None):
    """y = x W^T (+bias, +GELU) with x and W quantised to E4M3 on the fly (per-tensor scales); falls back to the 16-bit GEMM."""
    epi = (EPI_BIAS if bias is not None else EPI_NONE) if epi is None else epi
    out_dtype = out_dtype or x.dtype
    if x.is_cuda and _lib.available() and _F8 and x.shape[1] % 16 == 0:
        x8, sx = quantize_fp8(x)
        w8, sw = _quantize_weight_cached(w)
        y = gemm_fp8(x8, w8, 1.0, scale_a=sx, scale_b=sw, out_dtype=out_dtype, epi=epi, bias=bias, aux=aux)
        if y is not None:
            return y
    return linear_fwd(x, w, bias, epi=epi, aux=aux, out_dtype=out_dtype)
