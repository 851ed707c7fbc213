"""Channel-permutation search for 2:4 structured sparsity: find a column order of a weight matrix that maximises the magnitude kept
by 2:4 pruning. Reference: apex/contrib/sparsity/permutation_search_kernels/{exhaustive_search,channel_swap,permutation_utilities,
call_permutation_search_kernels}.py (1,443 LoC) over permutation_search_cuda (4 kernels).

Strategies (same names / options as the reference):
  * ``exhaustive`` / ``optimize_stripe_groups,<cols>,<escapes>`` — for every group of ``cols/4`` stripes (a stripe = 4 adjacent
    columns) score ALL distinct ways of re-dealing the group's columns into stripes, greedily apply the best non-overlapping
    ones, repeat until nothing improves; bounded random "escape" perturbations afterwards.
  * ``channel_swap,<escapes>`` — the same loop restricted to single column swaps between two stripes.
  * ``random,<n>`` — best of n random permutations.
All candidate scoring runs in csrc/perm_search.cu (one thread per candidate, stripe-group columns staged in shared memory); the
greedy selection is a few host-side steps per round. CPU tensors take a vectorised PyTorch path with the same results.
"""
from __future__ import annotations

import functools
import itertools
import time

import numpy as np
import torch

from ... import _lib

_lib.declare("ab_perm_eval", "p i i p i p p")
_lib.declare("ab_stripe_search", "p i i p i i p i p p p")


# ------------------------------------------------------------------------------------------------------------- candidates
@functools.lru_cache(maxsize=None)
def generate_all_unique_combinations(C: int, M: int = 4) -> np.ndarray:
    """All distinct ways to deal C columns into C/M unordered groups of M (order inside a group is irrelevant to 2:4 pruning).
    Canonical form: each group ascending, groups ordered by their first element. Row 0 is the identity. [P, C] uint8."""
    assert C % M == 0

    def rec(rest):
        if not rest:
            yield ()
            return
        first, others = rest[0], rest[1:]
        for comb in itertools.combinations(others, M - 1):
            grp = (first,) + comb
            left = tuple(x for x in others if x not in comb)
            for tail in rec(left):
                yield grp + tail

    out = np.array(list(rec(tuple(range(C)))), dtype=np.uint8)
    assert (out[0] == np.arange(C)).all()
    return out


@functools.lru_cache(maxsize=None)
def _swap_candidates() -> np.ndarray:
    """identity + the 16 single swaps between two stripes, as arrangements of 8 columns."""
    rows = [list(range(8))]
    for i in range(4):
        for j in range(4, 8):
            p = list(range(8))
            p[i], p[j] = p[j], p[i]
            rows.append(p)
    return np.array(rows, dtype=np.uint8)


# ------------------------------------------------------------------------------------------------------------------ scoring
def sum_after_2_to_4(matrix: torch.Tensor, perms: torch.Tensor | None = None) -> torch.Tensor:
    """Magnitude kept by 2:4 pruning along the columns; with ``perms`` [P, C] (int32) one value per column permutation."""
    m = matrix.detach().float().contiguous()
    R, C = m.shape
    if m.is_cuda and _lib.available():
        P = 1 if perms is None else perms.shape[0]
        out = torch.empty(P, dtype=torch.float32, device=m.device)
        pp = None if perms is None else perms.to(device=m.device, dtype=torch.int32).contiguous()
        _lib.fn("ab_perm_eval")(m.data_ptr(), R, C, _lib.ptr(pp), P, out.data_ptr(), _lib.stream_ptr(m.device))
        return out if perms is not None else out[0]
    a = m.abs()
    if perms is None:
        return a.view(R, C // 4, 4).topk(2, dim=-1).values.sum()
    return torch.stack([a[:, p.long()].view(R, C // 4, 4).topk(2, dim=-1).values.sum() for p in perms])


def _stripe_sums(m: torch.Tensor) -> torch.Tensor:
    R, C = m.shape
    return m.abs().view(R, C // 4, 4).topk(2, dim=-1).values.sum((0, 2))  # [C/4]


def _score_groups(m: torch.Tensor, groups: torch.Tensor, cands_np: np.ndarray):
    """-> (improvement [G], best candidate id [G]) for every stripe group."""
    R, C = m.shape
    G, S = groups.shape
    P, W = cands_np.shape
    base = _stripe_sums(m)[groups.long()].sum(1)
    noise = 2e-6 * base.abs()  # the two sides are summed in different orders: differences below fp32 rounding are not improvements
    if m.is_cuda and _lib.available():
        cands = torch.from_numpy(cands_np).to(m.device)
        chunks = (P + 255) // 256
        pv = torch.empty(G, chunks, dtype=torch.float32, device=m.device)
        pi = torch.empty(G, chunks, dtype=torch.int32, device=m.device)
        _lib.fn("ab_stripe_search")(m.data_ptr(), R, C, groups.data_ptr(), G, S, cands.data_ptr(), P, pv.data_ptr(), pi.data_ptr(),
                                    _lib.stream_ptr(m.device))
        best, ch = pv.max(1)
        # ties across chunks -> the lowest candidate id
        is_best = pv == best[:, None]
        idx = torch.where(is_best, pi, torch.full_like(pi, 2 ** 31 - 1)).min(1).values
        imp = best - base
        return torch.where((imp > noise) & (idx != 0), imp, torch.zeros_like(imp)), idx.long()
    a = m.abs()
    cands = torch.from_numpy(cands_np.astype(np.int64))
    cols = (groups.long()[:, :, None] * 4 + torch.arange(4)).reshape(G, W)          # [G, W] matrix columns of each group
    best = torch.empty(G)
    idx = torch.empty(G, dtype=torch.long)
    step = max(1, (1 << 22) // max(1, R * W))
    for g in range(G):
        sub = a[:, cols[g]]                                                          # [R, W]
        bv, bi = -1.0, 0
        for c0 in range(0, P, step):
            sc = sub[:, cands[c0:c0 + step]].reshape(R, -1, S, 4).topk(2, dim=-1).values.sum((0, 2, 3))
            v, i = sc.max(0)
            if float(v) > bv:
                bv, bi = float(v), c0 + int(i)
        best[g], idx[g] = bv, bi
    imp = best - base
    return torch.where((imp > noise) & (idx != 0), imp, torch.zeros_like(imp)), idx


# ---------------------------------------------------------------------------------------------------------------- strategies
def _all_groups(num_stripes: int, S: int, device) -> torch.Tensor:
    return torch.tensor(list(itertools.combinations(range(num_stripes), S)), dtype=torch.int32, device=device)


def _greedy(matrix: torch.Tensor, S: int, cands_np: np.ndarray, escape_attempts: int, permutation=None, seed: int = 0):
    start = time.perf_counter()
    m = matrix.detach().float().contiguous().clone()
    R, C = m.shape
    assert C % 4 == 0, "2:4 permutation search needs a multiple of 4 columns"
    S = min(S, C // 4)
    if S < 2:
        return matrix, time.perf_counter() - start, list(range(C)) if permutation is None else list(permutation)
    if cands_np.shape[1] != 4 * S:
        cands_np = generate_all_unique_combinations(4 * S, 4)
    perm = torch.arange(C, device=m.device) if permutation is None else torch.as_tensor(permutation, device=m.device).long()
    if permutation is not None:
        m = m[:, perm].contiguous()
    groups = _all_groups(C // 4, S, m.device)
    rng = np.random.default_rng(seed)
    eps = float(np.finfo(np.float16).tiny) * 5.0
    best_total, best_perm, escapes = float(sum_after_2_to_4(m)), perm.clone(), 0
    while True:
        imp, idx = _score_groups(m, groups, cands_np)
        order = torch.argsort(imp, descending=True).tolist()
        imp_h, idx_h, groups_h = imp.tolist(), idx.tolist(), groups.tolist()
        used: set = set()
        col_map = torch.arange(C, device=m.device)
        applied = 0
        for g in order:
            if imp_h[g] <= eps:
                break
            sg = groups_h[g]
            if any(s in used for s in sg):
                continue
            used.update(sg)
            cols = torch.tensor([s * 4 + k for s in sg for k in range(4)], device=m.device)
            col_map[cols] = cols[torch.from_numpy(cands_np[idx_h[g]].astype(np.int64)).to(m.device)]
            applied += 1
        if applied:
            m = m[:, col_map].contiguous()
            perm = perm[col_map]
            total = float(sum_after_2_to_4(m))
            if total > best_total:
                best_total, best_perm = total, perm.clone()
            continue
        if escapes >= escape_attempts:
            break
        # bounded regression: swap two random columns of different stripes and keep searching from there
        escapes += 1
        a, b = rng.choice(C // 4, 2, replace=False)
        ca, cb = int(a) * 4 + int(rng.integers(4)), int(b) * 4 + int(rng.integers(4))
        col_map = torch.arange(C, device=m.device)
        col_map[ca], col_map[cb] = cb, ca
        m = m[:, col_map].contiguous()
        perm = perm[col_map]
    out = matrix[:, best_perm.to(matrix.device)]
    return out, time.perf_counter() - start, best_perm.tolist()


def Exhaustive_Search(matrix, stripe_group_size=-1, escape_attempts=0, permutation=None):
    """Reference exhaustive_search.py:374-463. ``stripe_group_size`` in columns (8, 12 or 16); -1 = the whole matrix at once."""
    C = matrix.shape[1]
    if stripe_group_size == -1 or stripe_group_size >= C:
        if C > 16:
            raise ValueError("a full exhaustive search is only tractable up to 16 columns; pass stripe_group_size")
        stripe_group_size = C
    S = stripe_group_size // 4
    return _greedy(matrix, S, generate_all_unique_combinations(4 * S, 4), escape_attempts, permutation)


def Channel_Swap(matrix, escape_attempts=0, verbosity=0, permutation=None):
    """Reference channel_swap.py:209-265: greedy single-column swaps between stripes."""
    return _greedy(matrix, 2, _swap_candidates(), escape_attempts, permutation)


def Random_Search(matrix, num_seeds=10, seed=0):
    start = time.perf_counter()
    C = matrix.shape[1]
    g = torch.Generator().manual_seed(seed)
    perms = torch.stack([torch.arange(C)] + [torch.randperm(C, generator=g) for _ in range(num_seeds)]).int()
    best = None
    for c0 in range(0, perms.shape[0], 16384):
        sc = sum_after_2_to_4(matrix, perms[c0:c0 + 16384])
        v, i = sc.max(0)
        if best is None or float(v) > best[0]:
            best = (float(v), perms[c0 + int(i)].long())
    p = best[1]
    return matrix[:, p.to(matrix.device)], time.perf_counter() - start, p.tolist()


def accelerated_search_for_good_permutation(matrix_group, options=None, verbosity=0):
    """Reference call_permutation_search_kernels.py:6-105: ``options['strategy']`` in {exhaustive, progressive channel swap, random}
    -> the permutation (list of column indices) for ``matrix_group`` [rows, channels]."""
    options = dict(options or {})
    strategy = options.get("strategy", "exhaustive")
    m = torch.as_tensor(matrix_group)
    if strategy == "exhaustive":
        _, _, perm = Exhaustive_Search(m, stripe_group_size=options.get("stripe_group_size", 8), escape_attempts=options.get("escape_attempts", 100))
    elif strategy in ("progressive channel swap", "channel_swap"):
        _, _, perm = Channel_Swap(m, escape_attempts=options.get("escape_attempts", options.get("improvement_threshold", 0) and 0))
    elif strategy == "random":
        _, _, perm = Random_Search(m, num_seeds=options.get("num_seeds", 10))
    else:
        raise ValueError(f"unknown permutation search strategy {strategy!r}")
    return perm


def efficacy(optimal_lost_magnitude, base_lost_magnitude, cur_lost_magnitude):
    """Reference permutation_utilities.py:116-125: 0 = no better than the default order, 100 = as good as unstructured-rows."""
    if base_lost_magnitude == optimal_lost_magnitude:
        return 100.0
    return 100.0 * (base_lost_magnitude - cur_lost_magnitude) / (base_lost_magnitude - optimal_lost_magnitude)
