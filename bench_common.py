"""Pieces of the headline benchmark shared by BOTH arms of bench.py. Imports torch only — never ``apex_b200`` — so the reference
arm's process does not load any of this repo's package or native code.

  * Llama-3 parameter sets (shapes; random init — there is no network for checkpoints);
  * a plain-PyTorch functional Llama forward + loss (torch.nn.functional only) used by the end-to-end measurement of both arms,
    so that the only difference between the arms is the optimizer under test;
  * nvidia-smi clock / throttle-reason sampler and the measured peaks (MEASURED_PEAKS.json, profiles/results/link_peaks.json).
"""
from __future__ import annotations

import json
import statistics
import subprocess
import threading
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent

CONFIGS = {
    "llama3-8b": dict(vocab=128256, hidden=4096, layers=32, kv=1024, ffn=14336, head_dim=128),
    "llama3-1b-ish": dict(vocab=32768, hidden=2048, layers=16, kv=512, ffn=8192, head_dim=128),
    "tiny": dict(vocab=1024, hidden=256, layers=2, kv=128, ffn=512, head_dim=64),
}


def param_shapes(name: str = "llama3-8b"):
    c = CONFIGS[name]
    h = c["hidden"]
    shapes = [("tok_embeddings.weight", (c["vocab"], h))]
    for i in range(c["layers"]):
        p = f"layers.{i}."
        shapes += [(p + "attention.wq.weight", (h, h)), (p + "attention.wk.weight", (c["kv"], h)), (p + "attention.wv.weight", (c["kv"], h)),
                   (p + "attention.wo.weight", (h, h)), (p + "feed_forward.w1.weight", (c["ffn"], h)),
                   (p + "feed_forward.w3.weight", (c["ffn"], h)), (p + "feed_forward.w2.weight", (h, c["ffn"])),
                   (p + "attention_norm.weight", (h,)), (p + "ffn_norm.weight", (h,))]
    shapes += [("norm.weight", (h,)), ("output.weight", (c["vocab"], h))]
    return shapes


def make_params(name: str = "llama3-8b", device="cuda", dtype=torch.bfloat16, std: float = 0.02):
    out = []
    for n, s in param_shapes(name):
        p = torch.empty(s, device=device, dtype=dtype)
        if len(s) == 1:
            p.fill_(1.0)
        else:
            p.normal_(0.0, std)
        out.append((n, torch.nn.Parameter(p)))
    return out


def num_params(name: str = "llama3-8b") -> int:
    n = 0
    for _, s in param_shapes(name):
        k = 1
        for d in s:
            k *= d
        n += k
    return n


# ---------------------------------------------------------------------------------------------------------------------------
def _rope_tables(seq: int, dim: int, device, base: float = 500000.0):
    inv = 1.0 / (base ** (torch.arange(0, dim, 2, device=device, dtype=torch.float32) / dim))
    ang = torch.outer(torch.arange(seq, device=device, dtype=torch.float32), inv)
    return torch.cos(ang), torch.sin(ang)


def _rope(x, cos, sin):
    """x [B, heads, S, d]: rotate (even, odd) pairs."""
    x1, x2 = x.float().unflatten(-1, (-1, 2)).unbind(-1)
    y = torch.stack((x1 * cos - x2 * sin, x1 * sin + x2 * cos), -1).flatten(-2)
    return y.to(x.dtype)


def llama_loss(P: dict, tokens: torch.Tensor, labels: torch.Tensor, name: str) -> torch.Tensor:
    """Causal-LM loss of the named Llama architecture with parameters ``P`` (name -> tensor). Plain PyTorch: F.rms_norm, F.linear,
    F.scaled_dot_product_attention (GQA), SwiGLU, fp32 cross-entropy."""
    c = CONFIGS[name]
    H, hd = c["hidden"], c["head_dim"]
    nh, nkv = H // hd, c["kv"] // hd
    B, S = tokens.shape
    h = F.embedding(tokens, P["tok_embeddings.weight"])
    cos, sin = _rope_tables(S, hd, tokens.device)
    for i in range(c["layers"]):
        p = f"layers.{i}."
        x = F.rms_norm(h, (H,), P[p + "attention_norm.weight"], 1e-5)
        q = F.linear(x, P[p + "attention.wq.weight"]).view(B, S, nh, hd).transpose(1, 2)
        k = F.linear(x, P[p + "attention.wk.weight"]).view(B, S, nkv, hd).transpose(1, 2)
        v = F.linear(x, P[p + "attention.wv.weight"]).view(B, S, nkv, hd).transpose(1, 2)
        q, k = _rope(q, cos, sin), _rope(k, cos, sin)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True)
        h = h + F.linear(a.transpose(1, 2).reshape(B, S, H), P[p + "attention.wo.weight"])
        x = F.rms_norm(h, (H,), P[p + "ffn_norm.weight"], 1e-5)
        h = h + F.linear(F.silu(F.linear(x, P[p + "feed_forward.w1.weight"])) * F.linear(x, P[p + "feed_forward.w3.weight"]),
                         P[p + "feed_forward.w2.weight"])
    h = F.rms_norm(h, (H,), P["norm.weight"], 1e-5)
    logits = F.linear(h, P["output.weight"])
    return F.cross_entropy(logits.float().view(-1, c["vocab"]), labels.view(-1))


# ---------------------------------------------------------------------------------------------------------------------------
def measured_peaks():
    out = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback",
           "link_gbs": 900.0, "link_source": "nominal NVLink-5 per direction"}
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        out.update({k: d[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in d})
        out["source"] = "measured"
    lp = ROOT / "profiles" / "results" / "link_peaks.json"
    if lp.exists():
        try:
            d = json.loads(lp.read_text())
            if d.get("link_gbs"):
                out["link_gbs"], out["link_source"] = float(d["link_gbs"]), d.get("how", "benchmarks/bench_symm.py")
        except Exception:  # noqa: BLE001
            pass
    return out


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons in a background thread during a timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0, period_s=0.2):
        self.gpu_index, self.period = gpu_index, period_s
        self.rows = []
        self._stop = threading.Event()
        self._th = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu_index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        self._th = threading.Thread(target=self._loop, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._th:
            self._th.join(timeout=6)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}
