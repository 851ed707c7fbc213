"""2:4 (m:n) structured-sparsity mask calculators. Reference: apex/contrib/sparsity/sparse_masklib.py:11-230
(``m4n2_1d`` exhaustive 1-D, ``m4n2_2d_greedy``, ``m4n2_2d_best``, ``create_mask`` for 1-D/2-D/3-D/4-D weights).
The 1-D calculator keeps the n largest magnitudes of every group of m with a top-k (no pattern table, no matmul against all
permutations); the 2-D variants keep the reference's definitions (every row AND column of each m x m block is n:m)."""
from __future__ import annotations

import itertools

import torch

_patterns_2d: dict = {}


def _pad_cols(matrix, m):
    if matrix.shape[1] % m:
        pad = m - matrix.shape[1] % m
        return torch.nn.functional.pad(matrix, (0, pad)), matrix.shape
    return matrix, matrix.shape


def fill(x):
    """Fraction of non-zero entries (reference sparse_masklib.py:11-12)."""
    return float(x.nonzero().size(0)) / torch.numel(x)


def reshape_1d(matrix, m):
    """(h, w) -> (h*w/m, m), zero-padding w to a multiple of m; also returns the padded shape (reference :18-28)."""
    mat, _ = _pad_cols(matrix, m)
    return mat.reshape(-1, m), mat.shape


def compute_valid_1d_patterns(m, n):
    """All 0/1 vectors of length m with n ones (reference :35-46). The calculators here use top-k instead of this table."""
    return torch.tensor(sorted(set(itertools.permutations([1.0] * n + [0.0] * (m - n)))))


def mn_1d_best(matrix, m, n):
    """Keep the n largest |w| in every consecutive group of m along the last dim."""
    mat, shape = _pad_cols(matrix, m)
    g = mat.abs().reshape(-1, m)
    idx = g.topk(n, dim=1).indices
    mask = torch.zeros_like(g, dtype=torch.int32).scatter_(1, idx, 1)
    return mask.view(mat.shape)[:, :shape[1]].contiguous()


def m4n2_1d(mat, density):
    return mn_1d_best(mat, 4, 2)


def compute_valid_2d_patterns(m, n):
    key = (m, n)
    if key not in _patterns_2d:
        rows = [p for p in itertools.product((0, 1), repeat=m) if sum(p) == n]
        pats = []
        for combo in itertools.product(rows, repeat=m):
            t = torch.tensor(combo)
            if (t.sum(0) == n).all():
                pats.append(t)
        _patterns_2d[key] = torch.stack(pats).float()
    return _patterns_2d[key]


def _blocks(matrix, m):
    h, w = matrix.shape
    hp, wp = (h + m - 1) // m * m, (w + m - 1) // m * m
    mat = torch.nn.functional.pad(matrix, (0, wp - w, 0, hp - h))
    return mat.view(hp // m, m, wp // m, m).permute(0, 2, 1, 3).contiguous(), (h, w, hp, wp)


def _unblocks(blocks, dims, m):
    h, w, hp, wp = dims
    return blocks.permute(0, 2, 1, 3).contiguous().view(hp, wp)[:h, :w].contiguous()


def mn_2d_best(matrix, m, n):
    pats = compute_valid_2d_patterns(m, n).to(matrix.device)
    blocks, dims = _blocks(matrix.abs().float(), m)
    score = torch.matmul(blocks.view(*blocks.shape[:2], m * m), pats.view(pats.shape[0], m * m).t())
    best = pats[score.argmax(dim=2)]
    return _unblocks(best, dims, m).to(torch.int32)


def m4n2_2d_best(mat, density):
    return mn_2d_best(mat, 4, 2)


def mn_2d_greedy(matrix, m, n):
    blocks, dims = _blocks(matrix.abs().float().cpu(), m)
    out = torch.zeros_like(blocks)
    for bi in range(blocks.shape[0]):
        for bj in range(blocks.shape[1]):
            blk = blocks[bi, bj]
            order = torch.argsort(blk.flatten(), descending=True)
            rows, cols = [0] * m, [0] * m
            for k in order.tolist():
                r, c = divmod(k, m)
                if rows[r] < n and cols[c] < n:
                    out[bi, bj, r, c] = 1
                    rows[r] += 1
                    cols[c] += 1
    return _unblocks(out, dims, m).to(torch.int32).to(matrix.device)


def m4n2_2d_greedy(mat, density):
    return mn_2d_greedy(mat, 4, 2)


_FUNCS = {"m4n2_1d": m4n2_1d, "m4n2_2d_best": m4n2_2d_best, "m4n2_2d_greedy": m4n2_2d_greedy}


def create_mask(tensor, pattern="m4n2_1d", density=0.5):
    func = _FUNCS[pattern] if isinstance(pattern, str) else pattern
    shape, dtype = tensor.shape, tensor.dtype
    t = tensor.detach().float().contiguous()
    if t.dim() == 1:
        return func(t.view(1, -1), density).view(shape).to(dtype)
    if t.dim() == 2:
        return func(t, density).view(shape).to(dtype)
    if t.dim() == 3:   # 1-D convs (K, C, R): prune along C
        m = func(t.permute(0, 2, 1).contiguous().view(shape[0] * shape[2], shape[1]), density)
        return m.view(shape[0], shape[2], shape[1]).permute(0, 2, 1).contiguous().to(dtype)
    if t.dim() == 4:   # 2-D convs (K, C, R, S): prune along C
        m = func(t.permute(2, 3, 0, 1).contiguous().view(shape[2] * shape[3] * shape[0], shape[1]), density)
        return m.view(shape[2], shape[3], shape[0], shape[1]).permute(2, 3, 0, 1).contiguous().to(dtype)
    raise ValueError("create_mask supports 1-D to 4-D tensors")
