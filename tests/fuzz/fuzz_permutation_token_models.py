import sys; sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", ".."))
import random, torch, torch.nn as nn, torch.nn.functional as F
from apex_b200.contrib.sparsity.permutation_lib import Permutation as P
P.search_options = {"strategy": "exhaustive", "stripe_group_size": 8, "escape_attempts": 1}

class RandSeqNet(nn.Module):
    """Random token model over [B, S, H] tensors."""
    def __init__(self, rng):
        super().__init__()
        self.ops, self.mods = [], nn.ModuleDict()
        self.emb = nn.Embedding(30, 16) if rng.random() < 0.5 else None
        self.inp = None if self.emb is not None else nn.Linear(8, 16)
        width = {"x": 16}
        live = ["x"]
        n = 0
        def add_mod(m):
            nonlocal n
            n += 1; name = f"m{n}"; self.mods[name] = m; return name
        for step in range(rng.randint(4, 10)):
            kind = rng.choice(["lin", "lin", "ln", "gelu", "add", "cat", "mha", "mul_attr", "reuse", "ffn"])
            src = rng.choice(live); h = width[src]; out = f"t{step}"
            if kind == "lin":
                ho = rng.choice([16, 24, 32, 64]); name = add_mod(nn.Linear(h, ho, bias=rng.random() < 0.7))
                self.ops.append(("mod", out, [src], name)); width[out] = ho
            elif kind == "ffn":
                a = add_mod(nn.Linear(h, 4 * h)); b = add_mod(nn.Linear(4 * h, h))
                self.ops.append(("ffn", out, [src], (a, b))); width[out] = h
            elif kind == "ln":
                name = add_mod(nn.LayerNorm(h)); self.ops.append(("mod", out, [src], name)); width[out] = h
            elif kind == "gelu":
                self.ops.append(("gelu", out, [src], None)); width[out] = h
            elif kind == "add":
                others = [t for t in live if width[t] == h and t != src]
                if not others: continue
                self.ops.append(("add", out, [src, rng.choice(others)], None)); width[out] = h
            elif kind == "cat":
                other = rng.choice(live); self.ops.append(("cat", out, [src, other], None)); width[out] = h + width[other]
            elif kind == "mha":
                if h % 4: continue
                name = add_mod(nn.MultiheadAttention(h, 4, batch_first=True)); self.ops.append(("mha", out, [src], name)); width[out] = h
            elif kind == "mul_attr":
                pname = f"p{step}"; setattr(self, pname, nn.Parameter(torch.randn(h))); self.ops.append(("mul_attr", out, [src], pname)); width[out] = h
            elif kind == "reuse":
                cands = [nm for (k2, o, ins, nm) in self.ops if k2 == "mod" and isinstance(self.mods[nm], nn.Linear) and self.mods[nm].in_features == h]
                if not cands: continue
                nm = rng.choice(cands); self.ops.append(("mod", out, [src], nm)); width[out] = self.mods[nm].out_features
            live.append(out)
        used = {i for (_, _, ins, _) in self.ops for i in ins}
        self.leaves = [t for t in live if t not in used and t != "x"] or [live[-1]]
        self.heads = nn.ModuleList(nn.Linear(width[t], 5) for t in self.leaves)
    def forward(self, x):
        env = {"x": self.emb(x) if self.emb is not None else self.inp(x)}
        for kind, out, ins, name in self.ops:
            a = env[ins[0]]
            if kind == "mod": env[out] = self.mods[name](a)
            elif kind == "ffn": env[out] = a + self.mods[name[1]](F.gelu(self.mods[name[0]](a)))
            elif kind == "gelu": env[out] = F.gelu(a)
            elif kind == "add": env[out] = a + env[ins[1]]
            elif kind == "cat": env[out] = torch.cat([a, env[ins[1]]], dim=-1)
            elif kind == "mha": env[out] = self.mods[name](a, a, a)[0]
            elif kind == "mul_attr": env[out] = a * getattr(self, name)
        y = 0
        for h, t in zip(self.heads, self.leaves):
            y = y + h(env[t])
        return y

if __name__ == "__main__":
    bad = 0; permuted = 0
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        rng = random.Random(seed); torch.manual_seed(seed)
        net = RandSeqNet(rng).eval()
        for m in net.modules():
            if isinstance(m, nn.LayerNorm): m.weight.data.normal_(); m.bias.data.normal_()
        x = torch.randint(0, 30, (2, 6)) if net.emb is not None else torch.randn(2, 6, 8)
        y0 = net(x).detach()
        try:
            rep = P.permute_model(net)
        except Exception as e:
            print("seed", seed, "EXC", type(e).__name__, e); bad += 1; continue
        y1 = net(x).detach()
        err = (y1 - y0).abs().max().item() / (y0.abs().max().item() + 1e-9)
        permuted += len(rep)
        if err > 1e-4:
            print("seed", seed, "MISMATCH", err, [o[0] for o in net.ops]); bad += 1
    print("done bad", bad, "spaces permuted", permuted)
