import sys; sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", ".."))
import random, operator, torch, torch.nn as nn, torch.fx as fx
from apex_b200.contrib.torchsched.inductor.scheduler import MultiCudaStreamScheduler
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = random.Random(seed); torch.manual_seed(seed)
    g = fx.Graph(); root = nn.Module(); root.lin = nn.Linear(16, 16); root.w = nn.Parameter(torch.randn(16, 16) / 4)
    vals = [g.placeholder("a"), g.placeholder("b")]
    w = g.get_attr("w")
    for i in range(rng.randint(5, 30)):
        a = rng.choice(vals); kind = rng.randint(0, 11)
        if kind == 0: vals.append(g.call_function(torch.matmul, (a, w)))
        elif kind == 1: vals.append(g.call_module("lin", (a,)))
        elif kind == 2: vals.append(g.call_function(torch.add, (a, rng.choice(vals)), {"alpha": rng.choice([1, 2, 0.5])}))
        elif kind == 3:
            sp = g.call_function(torch.chunk, (a, 2), {"dim": -1}); x0 = g.call_function(operator.getitem, (sp, 0)); x1 = g.call_function(operator.getitem, (sp, 1))
            vals.append(g.call_function(torch.cat, ([x1, x0],), {"dim": -1}))
        elif kind == 4: vals.append(g.call_function(operator.getitem, (a, (slice(None), slice(0, 16, None)))))
        elif kind == 5: vals.append(g.call_method("to", (a,), {"dtype": torch.float32}))
        elif kind == 6: vals.append(g.call_function(torch.clamp, (a,), {"min": -1.0, "max": float("inf")}))
        elif kind == 7: vals.append(g.call_function(torch.where, (g.call_function(torch.gt, (a, 0)), a, rng.choice(vals))))
        elif kind == 8: vals.append(g.call_method("mul", (a, rng.choice([2, 0.5, True]))))
        elif kind == 9: vals.append(g.call_function(torch.nn.functional.layer_norm, (a, (16,)), {"eps": 1e-5}))
        elif kind == 10: vals.append(g.call_function(torch.softmax, (a,), {"dim": -1, "dtype": None}))
        else: vals.append(g.call_function(torch.full_like, (a, rng.choice([0.0, 1.5]))))
    leaves = [v for v in vals[2:] if not v.users] or [vals[-1]]
    out_struct = rng.choice(["tuple", "dict", "single"])
    g.output(tuple(leaves) if out_struct == "tuple" else ({"y": leaves[0], "rest": list(leaves[1:])} if out_struct == "dict" else leaves[0]))
    gm = fx.GraphModule(root, g)
    args = (torch.randn(4, 16), torch.randn(4, 16))
    try:
        want = gm(*args)
        class Rec(fx.Interpreter):
            def run_node(self, n):
                o = super().run_node(n); n.meta["example_value"] = o; return o
        Rec(gm).run(*args)
        s = MultiCudaStreamScheduler(gm, num_streams=rng.choice([0, 2, 8]), multi_stream=False)
        s.codegen(); fn = s.compile()
        got = fn(*args)
        flat = lambda o: [o] if isinstance(o, torch.Tensor) else [t for v in (o.values() if isinstance(o, dict) else o) for t in flat(v)]
        for x, y in zip(flat(got), flat(want)):
            assert torch.equal(x, y)
        # the multi-stream text must at least be generated and compile as Python
        s2 = MultiCudaStreamScheduler(gm, num_streams=3, multi_stream=True); src = s2.codegen(); compile(src, "<x>", "exec")
    except Exception as e:
        import traceback; print("seed", seed, type(e).__name__, str(e)[:150], traceback.format_exc().splitlines()[-3][:150]); bad += 1
print("bad", bad)
