import sys; sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", ".."))
import itertools, torch
from apex_b200.contrib.sparsity import sparse_masklib as M
torch.manual_seed(0)
bad = 0
# all valid 4x4 2d patterns: each row and col has exactly 2 ones
valid = [torch.tensor(p, dtype=torch.float32).reshape(4,4) for p in itertools.product([0,1], repeat=16)
         if all(sum(p[r*4:(r+1)*4])==2 for r in range(4)) and all(sum(p[c::4])==2 for c in range(4))]
print(len(valid), "valid 2d patterns")
for trial in range(300):
    shape_kind = trial % 4
    if shape_kind == 0: shape = (torch.randint(1, 9, (1,)).item()*4, torch.randint(1, 9, (1,)).item()*4)
    elif shape_kind == 1: shape = (torch.randint(1, 20, (1,)).item(), torch.randint(1, 9, (1,)).item()*4)
    elif shape_kind == 2: shape = (torch.randint(1, 9,(1,)).item()*4, torch.randint(1, 9, (1,)).item()*4, 3, 3)
    else: shape = (torch.randint(1, 6,(1,)).item()*4, torch.randint(1, 6, (1,)).item()*4, 3)
    w = torch.randn(*shape)
    for pattern in ("m4n2_1d", "m4n2_2d_best", "m4n2_2d_greedy"):
        if pattern != "m4n2_1d" and (shape[0] % 4 or (len(shape) == 2 and shape[1] % 4)):
            continue
        try:
            mask = M.create_mask(w, pattern)
        except Exception as e:
            print("EXC", shape, pattern, type(e).__name__, e); bad += 1; continue
        if mask.shape != w.shape:
            print("SHAPE", shape, pattern, mask.shape); bad += 1; continue
        # 2d view as the reference does: conv [K,C,R,S] -> permute(2,3,0,1) -> [R*S*K, C]; 3d [K,C,R] -> permute(0,2,1)? check on the 1d property along C
        if len(shape) == 2:
            m2, w2 = mask.float(), w
        elif len(shape) == 4:
            m2, w2 = mask.float().permute(2,3,0,1).reshape(-1, shape[1]), w.permute(2,3,0,1).reshape(-1, shape[1])
        else:
            m2, w2 = mask.float().permute(0,2,1).reshape(-1, shape[1]), w.permute(0,2,1).reshape(-1, shape[1])
        g = m2.reshape(m2.shape[0], -1, 4).sum(-1)
        if pattern == "m4n2_1d":
            if not bool((g == 2).all()): print("1D not 2:4", shape); bad += 1
            kept = (w2.abs() * m2).sum(); best = w2.abs().reshape(w2.shape[0], -1, 4).topk(2, -1).values.sum()
            if abs(float(kept - best)) > 1e-3: print("1D not optimal", shape, float(kept), float(best)); bad += 1
        else:
            if not bool((g <= 2).all()): print(pattern, "row groups > 2", shape); bad += 1
            if len(shape) == 2:
                blocks = m2.reshape(shape[0]//4, 4, shape[1]//4, 4).permute(0,2,1,3)
                if not bool((blocks.sum(-2) <= 2).all()): print(pattern, "col groups > 2", shape); bad += 1
                if pattern == "m4n2_2d_best":
                    wb = w2.abs().reshape(shape[0]//4, 4, shape[1]//4, 4).permute(0,2,1,3).reshape(-1, 16)
                    pat = torch.stack(valid).reshape(-1, 16)
                    best = (wb @ pat.T).max(1).values.sum()
                    kept = (w2.abs() * m2).sum()
                    if abs(float(kept - best)) > 1e-3: print("2D best not optimal", shape, float(kept), float(best)); bad += 1
print("bad", bad)
