"""Head-to-head single-GPU op benchmarks: apex_b200 vs the UNMODIFIED reference (baseline/_ref: apex.optimizers, apex.normalization,
apex.contrib.xentropy, the Megatron softmax extensions) on the same box, same tensors. CUDA events, L2 flushed between iterations.
BASELINE.md §2 rows 2 and '+'. Prints one JSON line per case and writes gpurun_out/bench_vs_reference.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "baseline", "_ref"))
from apex_b200.utils.timing import measured_peaks, time_fn  # noqa: E402

rows = []


def rec(case, ours_ms, ref_ms, gbytes=None, **kw):
    r = {"case": case, "ours_ms": round(ours_ms, 4), "reference_ms": round(ref_ms, 4) if ref_ms else None,
         "speedup": round(ref_ms / ours_ms, 2) if ref_ms else None}
    if gbytes:
        r["ours_GBps"] = round(gbytes / ours_ms * 1e3, 0)
        r["ours_frac_hbm"] = round(r["ours_GBps"] / measured_peaks()["hbm_gbs"], 3)
    r.update(kw)
    rows.append(r)
    print(json.dumps(r), flush=True)


def opt_case(name, ours_cls, ref_cls, n_tensors, bytes_per_elem, **kw):
    g = torch.Generator().manual_seed(0)
    sizes = torch.randint(1000, 200000, (n_tensors,), generator=g).tolist()
    t = {}
    for tag, cls in (("ours", ours_cls), ("ref", ref_cls)):
        if cls is None:
            t[tag] = None
            continue
        ps = [torch.nn.Parameter(torch.randn(n, device="cuda")) for n in sizes]
        for p in ps:
            p.grad = torch.randn_like(p)
        opt = cls(ps, lr=1e-3, **kw)
        opt.step()
        t[tag] = time_fn(lambda: opt.step(), warmup=3, iters=10)[0]
        del ps, opt
    rec(f"{name} step, {n_tensors} fp32 tensors ({sum(sizes) / 1e6:.0f} M elements)", t["ours"], t["ref"], sum(sizes) * bytes_per_elem / 1e9)


def main():
    dev = torch.device("cuda:0")
    import apex_b200.optimizers as O
    try:
        import apex.optimizers as RO
        import apex.normalization as RN
    except Exception as e:
        print(json.dumps({"reference": "unavailable", "why": str(e)[:200]}))
        RO = RN = None
    opt_case("FusedAdam", O.FusedAdam, RO and RO.FusedAdam, 10000, 28, weight_decay=0.01)
    opt_case("FusedLAMB", O.FusedLAMB, RO and RO.FusedLAMB, 10000, 44)
    opt_case("FusedSGD(momentum)", O.FusedSGD, RO and RO.FusedSGD, 10000, 20, momentum=0.9)
    opt_case("FusedNovoGrad", O.FusedNovoGrad, RO and RO.FusedNovoGrad, 10000, 24)
    # normalisation
    import apex_b200.normalization as N
    for hidden in (1024, 4096, 8192, 16384):
        nrows = (1 << 29) // (hidden * 2)
        for name in ("FusedLayerNorm", "FusedRMSNorm"):
            res = {}
            for tag, mod in (("ours", N), ("ref", RN)):
                if mod is None:
                    res[tag] = (None, None)
                    continue
                m = getattr(mod, name)(hidden).to(dev, torch.bfloat16)
                x = torch.randn(nrows, hidden, device=dev, dtype=torch.bfloat16, requires_grad=True)
                dy = torch.randn_like(x)
                f = time_fn(lambda: m(x), warmup=3, iters=10)[0]
                y = m(x)
                b = time_fn(lambda: torch.autograd.grad(y, (x, m.weight), dy, retain_graph=True), warmup=3, iters=10)[0]
                res[tag] = (f, b)
                del m, x, dy, y
            nb = nrows * hidden * 2 / 1e9
            rec(f"{name} fwd bf16 h={hidden} rows={nrows}", res["ours"][0], res["ref"][0], 2 * nb)
            rec(f"{name} bwd bf16 h={hidden} rows={nrows}", res["ours"][1], res["ref"][1], 3 * nb)
    # fused softmax cross-entropy (BASELINE: N=128*74 rows, 32320 classes)
    from apex_b200.contrib.xentropy import SoftmaxCrossEntropyLoss as OX
    try:
        from apex.contrib.xentropy import SoftmaxCrossEntropyLoss as RX
    except Exception:
        RX = None
    R, C = 128 * 74, 32320
    lab = torch.randint(0, C, (R,), device=dev)
    res = {}
    for tag, fn in (("ours", OX), ("ref", RX)):
        if fn is None:
            res[tag] = (None, None)
            continue
        x = torch.randn(R, C, device=dev, dtype=torch.bfloat16 if tag == "ours" else torch.float16, requires_grad=True)
        f = time_fn(lambda: fn.apply(x, lab, 0.1, 0, True), warmup=3, iters=10)[0]
        loss = fn.apply(x, lab, 0.1, 0, True)
        gl = torch.ones_like(loss)
        b = time_fn(lambda: torch.autograd.grad(loss, x, gl, retain_graph=True), warmup=3, iters=10)[0]
        res[tag] = (f, b)
    nb = R * C * 2 / 1e9
    rec(f"xentropy fwd {R}x{C} 16-bit logits", res["ours"][0], res["ref"][0], nb)
    rec(f"xentropy bwd {R}x{C} 16-bit logits", res["ours"][1], res["ref"][1], 2 * nb)
    # causal softmax [attn_batches, s, s]
    from apex_b200.transformer.functional import scaled_upper_triang_masked_softmax as ours_sm
    try:
        import scaled_upper_triang_masked_softmax_cuda as RS
    except Exception:
        RS = None
    for s_len in (2048, 4096):
        x = torch.randn(32, s_len, s_len, device=dev, dtype=torch.bfloat16)
        fo = time_fn(lambda: ours_sm(x, 0.5), warmup=3, iters=10)[0]
        fr = time_fn(lambda: RS.forward(x, 0.5), warmup=3, iters=10)[0] if RS is not None else None
        rec(f"scaled_upper_triang_masked_softmax fwd 32x{s_len}x{s_len} bf16", fo, fr, x.numel() * 2 * 1.0 / 1e9)  # ~half read + half written
        del x
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"peaks": measured_peaks(), "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "bench_vs_reference.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
