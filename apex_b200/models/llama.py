"""Llama-3 parameter sets (shapes only — the optimizer benchmarks need the parameter tensors, not the forward pass).

Llama-3-8B: 291 tensors, 8,030,261,248 parameters: embed + lm_head 128256x4096; 32 x {wq, wo 4096x4096; wk, wv
1024x4096; gate, up 14336x4096; down 4096x14336; 2 RMSNorm 4096}; final norm."""
from __future__ import annotations

import torch

CONFIGS = {
    "llama3-8b": dict(vocab=128256, hidden=4096, layers=32, kv=1024, ffn=14336),
    "llama3-1b-ish": dict(vocab=32768, hidden=2048, layers=16, kv=512, ffn=8192),
    "tiny": dict(vocab=1024, hidden=256, layers=2, kv=64, ffn=512),
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
    """Random-init parameters of the named architecture (there is no network for checkpoints)."""
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
