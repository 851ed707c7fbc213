import sys; sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", ".."))
import random, torch, torch.nn as nn, torch.nn.functional as F
from apex_b200.contrib.sparsity.permutation_lib import Permutation as P
P.search_options = {"strategy": "exhaustive", "stripe_group_size": 8, "escape_attempts": 1}

class RandNet(nn.Module):
    """Random conv net: a list of ops over a dict of live tensors."""
    def __init__(self, rng):
        super().__init__()
        self.ops = []           # (kind, out_name, in_names, module_name or None, extra)
        self.mods = nn.ModuleDict()
        chans = {"x": 8}
        live = ["x"]
        n = 0
        def add_mod(m):
            nonlocal n
            n += 1; name = f"m{n}"; self.mods[name] = m; return name
        widths = [8, 16, 24, 32]
        for step in range(rng.randint(4, 10)):
            kind = rng.choice(["conv", "conv", "conv1", "bn", "relu", "add", "cat", "dw", "gconv", "gn", "mul_attr", "pool", "reuse"])
            src = rng.choice(live)
            c = chans[src]
            out = f"t{step}"
            if kind in ("conv", "conv1"):
                co = rng.choice(widths); k = 3 if kind == "conv" else 1
                name = add_mod(nn.Conv2d(c, co, k, padding=k // 2, bias=rng.random() < 0.5))
                self.ops.append(("mod", out, [src], name)); chans[out] = co
            elif kind == "bn":
                name = add_mod(nn.BatchNorm2d(c)); self.ops.append(("mod", out, [src], name)); chans[out] = c
            elif kind == "gn":
                g = rng.choice([1, 2, c]) if c % 2 == 0 else 1
                name = add_mod(nn.GroupNorm(g, c)); self.ops.append(("mod", out, [src], name)); chans[out] = c
            elif kind == "relu":
                self.ops.append(("relu", out, [src], None)); chans[out] = c
            elif kind == "pool":
                name = add_mod(nn.AvgPool2d(3, stride=1, padding=1)); self.ops.append(("mod", out, [src], name)); chans[out] = c
            elif kind == "dw":
                name = add_mod(nn.Conv2d(c, c, 3, padding=1, groups=c)); self.ops.append(("mod", out, [src], name)); chans[out] = c
            elif kind == "gconv":
                g = rng.choice([2, 4]);
                if c % g or (c // g) % 4: continue
                co = rng.choice([c, 2 * c])
                if co % g: continue
                name = add_mod(nn.Conv2d(c, co, 3, padding=1, groups=g)); self.ops.append(("mod", out, [src], name)); chans[out] = co
            elif kind == "add":
                others = [t for t in live if chans[t] == c and t != src]
                if not others: continue
                self.ops.append(("add", out, [src, rng.choice(others)], None)); chans[out] = c
            elif kind == "cat":
                other = rng.choice(live)
                self.ops.append(("cat", out, [src, other], None)); chans[out] = c + chans[other]
            elif kind == "mul_attr":
                pname = f"p{step}"; setattr(self, pname, nn.Parameter(torch.randn(1, c, 1, 1)))
                self.ops.append(("mul_attr", out, [src], pname)); chans[out] = c
            elif kind == "reuse":
                cands = [(o, nm) for (k2, o, ins, nm) in self.ops if k2 == "mod" and isinstance(self.mods[nm], nn.Conv2d) and self.mods[nm].in_channels == c and self.mods[nm].groups == 1]
                if not cands: continue
                _, nm = rng.choice(cands)
                self.ops.append(("mod", out, [src], nm)); chans[out] = self.mods[nm].out_channels
            live.append(out)
        # head: every leaf goes through a 1x1 conv to 4 channels and is summed
        used = {i for (_, _, ins, _) in self.ops for i in ins}
        leaves = [t for t in live if t not in used and t != "x"] or [live[-1]]
        self.leaves = leaves
        self.head_kind = [rng.choice(["conv", "pool_linear", "flat_linear"]) for _ in leaves]
        self.heads = nn.ModuleList(nn.Conv2d(chans[t], 4, 1) if k == "conv" else nn.Linear(chans[t] * (1 if k == "pool_linear" else 25), 4)
                                   for t, k in zip(leaves, self.head_kind))
    def forward(self, x):
        env = {"x": x}
        for kind, out, ins, name in self.ops:
            if kind == "mod": env[out] = self.mods[name](env[ins[0]])
            elif kind == "relu": env[out] = F.relu(env[ins[0]])
            elif kind == "add": env[out] = env[ins[0]] + env[ins[1]]
            elif kind == "cat": env[out] = torch.cat([env[ins[0]], env[ins[1]]], dim=1)
            elif kind == "mul_attr": env[out] = env[ins[0]] * getattr(self, name)
        y = 0
        for h, t, k in zip(self.heads, self.leaves, self.head_kind):
            if k == "conv":
                y = y + h(env[t]).mean((2, 3))
            elif k == "pool_linear":
                y = y + h(torch.flatten(F.adaptive_avg_pool2d(env[t], 1), 1))
            else:
                y = y + h(torch.flatten(env[t], 1))
        return y

bad = 0; permuted = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = random.Random(seed); torch.manual_seed(seed)
    net = RandNet(rng).eval()
    for m in net.modules():
        if isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
            m.weight.data.normal_(); m.bias.data.normal_()
            if hasattr(m, "running_mean"): m.running_mean.normal_(); m.running_var.uniform_(0.5, 2)
    x = torch.randn(2, 8, 5, 5)
    y0 = net(x).detach()
    try:
        rep = P.permute_model(net)
    except Exception as e:
        print("seed", seed, "EXC", type(e).__name__, e); bad += 1; continue
    y1 = net(x).detach()
    err = (y1 - y0).abs().max().item() / (y0.abs().max().item() + 1e-9)
    permuted += len(rep)
    if err > 1e-4:
        print("seed", seed, "MISMATCH", err, [o[0] + ":" + str(o[3]) for o in net.ops]); bad += 1
print("done bad", bad, "spaces permuted", permuted)
