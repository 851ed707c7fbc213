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
    elif strategy == "user defined":
        # the reference leaves the permutation untouched here and expects the user to replace this function; a callable under
        # options["function"] (matrix -> permutation) is the same hook without monkey-patching
        fn = options.get("function")
        perm = [int(c) for c in fn(m)] if callable(fn) else list(range(m.shape[1]))
    else:   # as the reference: report and keep the channel order
        if verbosity >= 0:
            print("[accelerated_search_for_good_permutation] Cannot find the implementation of the required strategy!")
        perm = list(range(m.shape[1]))
    return perm


def efficacy(optimal_lost_magnitude, base_lost_magnitude, cur_lost_magnitude):
    """Reference permutation_utilities.py:116-125: 0 = no better than the default order, 100 = as good as unstructured-rows."""
    if base_lost_magnitude == optimal_lost_magnitude:
        return 100.0
    return 100.0 * (base_lost_magnitude - cur_lost_magnitude) / (base_lost_magnitude - optimal_lost_magnitude)


# ----------------------------------------------------------------------------------------------------------- analysis utilities
# The helper surface of the reference's permutation_utilities.py / exhaustive_search.py / channel_swap.py that its analysis scripts and
# tests use (apply_2_to_4, try_swap, search_matrix, permutation_distance, ...). Matrices may be numpy arrays or tensors; results come back
# in the kind that went in. Everything is vectorised over rows (the reference loops over rows and stripes in Python on its CPU path).
def _t(matrix) -> torch.Tensor:
    return matrix if isinstance(matrix, torch.Tensor) else torch.as_tensor(np.asarray(matrix))


def _back(result: torch.Tensor, like):
    return result if isinstance(like, torch.Tensor) else result.cpu().numpy()


def use_gpu(initial_override: bool = True) -> bool:
    """Whether candidate scoring runs in csrc/perm_search.cu (reference permutation_utilities.py:23-43 probes nvidia-smi and the extension)."""
    return bool(initial_override) and torch.cuda.is_available() and _lib.available()


def apply_2_to_4(matrix):
    """Zero the two smallest magnitudes of every group of 4 adjacent columns, in place (reference :46-52)."""
    m = _t(matrix)
    R, C = m.shape
    drop = m.abs().reshape(R, C // 4, 4).argsort(dim=-1)[..., :2]                      # the two smallest of each group
    mask = torch.ones(R, C // 4, 4, dtype=torch.bool, device=m.device).scatter_(-1, drop, False).reshape(R, C)
    if isinstance(matrix, torch.Tensor):
        matrix.mul_(mask)
    else:
        matrix[~mask.numpy()] = 0
    return matrix


def unstructured_prune(matrix, sparsity: float):
    """Zero the ``sparsity`` fraction of entries with the smallest magnitude (whole-matrix ranking). The reference (:87-94) ranks by signed
    value; its callers pass magnitudes, for which the two agree."""
    m = _t(matrix).clone()
    k = int(m.numel() * sparsity)
    if k:
        flat = m.reshape(-1)
        flat[flat.abs().argsort()[:k]] = 0
    return _back(m, matrix)


def magnitude_after_pruning_rows(matrix, rate: float = 0.5):
    """Magnitude kept when every row independently drops its ``rate`` smallest entries: the bound no column permutation can beat (:127-135)."""
    a = _t(matrix).abs().float()
    return float(a.sort(dim=1).values[:, int(a.shape[1] * rate):].sum())


def _stripe_of(col: int) -> slice:
    return slice(col // 4 * 4, col // 4 * 4 + 4)


def try_swap(matrix, dst: int, src: int):
    """(magnitude the two touched stripes keep after swapping columns ``src`` and ``dst``, improvement over not swapping) — the matrix is
    left as it was (:98-112)."""
    m = _t(matrix).float()
    before = float(sum_after_2_to_4(m[:, _stripe_of(src)].contiguous())) + float(sum_after_2_to_4(m[:, _stripe_of(dst)].contiguous()))
    sw = m.clone()
    sw[:, [src, dst]] = m[:, [dst, src]]
    after = float(sum_after_2_to_4(sw[:, _stripe_of(src)].contiguous())) + float(sum_after_2_to_4(sw[:, _stripe_of(dst)].contiguous()))
    if src // 4 == dst // 4:            # same stripe: one stripe counted twice on both sides, nothing can change
        return after, 0.0
    return after, after - before


def try_permutations_on_matrix(matrix, permutations):
    """(best improvement over the identity order, the permutation that achieves it) among ``permutations`` [P, C] (:144-171)."""
    m = _t(matrix).float()
    perms = torch.as_tensor(np.asarray(permutations).astype(np.int64))
    base = float(sum_after_2_to_4(m))
    best_v, best_i = -float("inf"), 0
    for c0 in range(0, perms.shape[0], 16384):
        sc = sum_after_2_to_4(m, perms[c0:c0 + 16384].int())
        v, i = sc.max(0)
        if float(v) > best_v:
            best_v, best_i = float(v), c0 + int(i)
    return best_v - base, np.asarray(permutations)[best_i]


def find_permutation(A, B) -> list:
    """For every column of A, the index of an identical column of B (:174-182); columns without a twin are skipped, like the reference."""
    a, b = _t(A), _t(B)
    eq = (a.T[:, None, :] == b.T[None, :, :]).all(-1)                                  # [colsA, colsB]
    return [int(row.float().argmax()) for row in eq if bool(row.any())]


def predict_unique_combinations(C: int, M: int) -> int:
    """Number of ways to deal C columns into C/M unordered groups of M: C! / (M!^(C/M) (C/M)!) (exhaustive_search.py:102-105), in
    exact integer arithmetic (the reference goes through floats and loses the low digits beyond 2^53)."""
    import math

    assert C % M == 0
    G = C // M
    return math.factorial(C) // (math.factorial(M) ** G * math.factorial(G))


def is_canonical(perm, col: int) -> bool:
    """May ``col`` extend the partial arrangement ``perm`` in canonical form (groups ascending inside, ordered by first element)?
    (exhaustive_search.py:21-32)"""
    if len(perm) % 4 == 0:
        return all(v in perm for v in range(col)) and (len(perm) == 0 or col > perm[-4])
    return col > perm[-1]


def search_matrix(matrix, group_width: int = 4):
    """Try every distinct arrangement of the whole matrix (<= 16 columns in practice): (permuted matrix, seconds, permutation,
    improvement); refuses with (matrix, prediction, identity) beyond 1e10 candidates (exhaustive_search.py:114-148)."""
    start = time.perf_counter()
    C = matrix.shape[1]
    prediction = predict_unique_combinations(C, group_width)
    identity = list(range(C))
    if prediction > 1e10:
        print(f"There are {prediction} unique combinations with {C} columns and a group width of {group_width}, not searching.")
        return matrix, prediction, identity
    if group_width != 4:
        raise NotImplementedError("scoring is 2:4 specific: group_width must be 4")
    improvement, perm = try_permutations_on_matrix(matrix, generate_all_unique_combinations(C, group_width))
    if improvement <= 0:
        improvement, perm = 0.0, np.asarray(identity)
    perm = [int(c) for c in perm]
    return matrix[:, perm], time.perf_counter() - start, perm, improvement


def generate_stripe_groups(num_stripes: int, window_size: int) -> set:
    """All ascending ``window_size``-tuples of stripe indices (exhaustive_search.py:184-199)."""
    return set(itertools.combinations(range(num_stripes), window_size))


def collect_stripes(matrix, stripes, group_width: int = 4):
    """The columns of the listed stripes, side by side (exhaustive_search.py:156-162)."""
    cols = [s * group_width + k for s in stripes for k in range(group_width)]
    return matrix[:, cols]


def apply_stripe_group_permutation(sgp, stripes, group_width: int, permutation):
    """Fold a permutation ``sgp`` of the columns gathered by :func:`collect_stripes` back into the full-width ``permutation``
    (exhaustive_search.py:166-180)."""
    cols = np.asarray([s * group_width + k for s in stripes for k in range(group_width)])
    out = np.array(permutation, copy=True)
    out[cols] = np.asarray(permutation)[cols[np.asarray(sgp, dtype=np.int64)]]
    return out if isinstance(permutation, np.ndarray) else out.tolist()


def stripes_and_swap_idx_to_columns(stripe0: int, stripe1: int, idx: int):
    """Swap number ``idx`` (0..15) between two stripes -> the two matrix columns it exchanges (channel_swap.py:30-37)."""
    if not 0 <= idx < 16:
        return None
    return stripe0 * 4 + idx // 4, stripe1 * 4 + idx % 4


def columns_to_stripes_and_swap_idx(col0: int, col1: int):
    """Inverse of :func:`stripes_and_swap_idx_to_columns` (channel_swap.py:41-53)."""
    return col0 // 4, col1 // 4, (col0 % 4) * 4 + col1 % 4


def build_stripe_pairs(matrix, used_stripes):
    """Stripe pairs whose swap scores are stale because one of their stripes changed (channel_swap.py:57-67)."""
    total = matrix.shape[1] // 4
    used = set(int(s) for s in used_stripes)
    return np.asarray([[a, b] for a in range(total - 1) for b in range(a, total) if a in used or b in used])


# ---- distance between two permutations: how many column swaps turn B's grouping into A's ---------------------------------------------
def make_grouped(A, width: int = 4) -> list:
    """The permutation as a list of groups of ``width``, each sorted (order inside a stripe does not matter to 2:4 pruning)."""
    A = [int(v) for v in A]
    return [sorted(A[i:i + width]) for i in range(0, len(A), width)]


def _group_sets(A):
    return {tuple(g) for g in make_grouped(A)}


def common_groups(A, B) -> list:
    return [list(g) for g in sorted(_group_sets(A) & _group_sets(B))]


def remove_common_groups(A, B):
    """(A, B) without the groups they share, flattened again in canonical order."""
    sa, sb = _group_sets(A), _group_sets(B)
    return [v for g in sorted(sa - sb) for v in g], [v for g in sorted(sb - sa) for v in g]


def group_differences(A, B) -> list:
    """(value, its group in B, the group of A that holds it) for every value of B that sits in a different group position than in A."""
    where_a = {v: i for i, g in enumerate(make_grouped(A)) for v in g}
    return [(v, i, where_a[v]) for i, g in enumerate(make_grouped(B)) for v in g if where_a[v] != i]


def dictify(wrong_entries) -> dict:
    out: dict = {}
    for val, cur, want in wrong_entries:
        out.setdefault((cur, want), []).append(val)
    return out


def move_groups_to_match(B, A, debug: bool = False) -> list:
    """B with its groups reordered so that group i of B is the one sharing the most values with group i of A (an assignment problem;
    the reference resolves it greedily, permutation_utilities.py:286-401)."""
    from scipy.optimize import linear_sum_assignment

    ga, gb = make_grouped(A), make_grouped(B)
    overlap = np.array([[len(set(x) & set(y)) for y in gb] for x in ga])
    rows, cols = linear_sum_assignment(-overlap)
    order = [int(cols[list(rows).index(i)]) for i in range(len(ga))]
    return [v for i in order for v in gb[i]]


def swap_and_correct(permutation, src: int, tgt: int) -> list:
    """Exchange positions ``src`` and ``tgt`` and return the permutation in canonical (sorted inside groups) form."""
    p = [int(v) for v in permutation]
    p[src], p[tgt] = p[tgt], p[src]
    return [v for g in make_grouped(p) for v in g]


def move_permutation_towards(B, A, debug: bool = False) -> list:
    """One swap of B that moves its grouping towards A's: prefer a swap that puts BOTH exchanged values into their groups."""
    B = move_groups_to_match(B, A, debug)
    wrong = dictify(group_differences(A, B))
    if not wrong:
        return B
    pos = {v: i for i, v in enumerate(B)}
    for (cur, want), vals in wrong.items():
        if (want, cur) in wrong:                                  # two values that want each other's group
            return swap_and_correct(B, pos[vals[0]], pos[wrong[(want, cur)][0]])
    (cur, want), vals = next(iter(wrong.items()))
    partner = next(v for (c, _), vs in wrong.items() if c == want for v in vs)     # somebody in the wanted group is misplaced too
    return swap_and_correct(B, pos[vals[0]], pos[partner])


def permutation_distance(A, B, matrix=None, magnitude_targets=None, debug: bool = False, verbosity: int = 0):
    """(number of swaps that turn B's grouping into A's, per magnitude target the (magnitude, permutation) met on the way that came
    closest to it — or None). Reference permutation_utilities.py:558-618."""
    A, B = [int(v) for v in A], [int(v) for v in B]
    swaps, common = 0, []
    limit = 2 ** max(len(A) // 4 - 1, 0) + 3
    results = None
    if magnitude_targets is not None:
        assert matrix is not None, "magnitude targets need the matrix"
        start = float(sum_after_2_to_4(_t(matrix).float()[:, A].contiguous()))
        results = [(start, list(A)) for _ in magnitude_targets]
    while _group_sets(A) != _group_sets(B):
        common += [v for g in common_groups(A, B) for v in g]
        A, B = remove_common_groups(A, B)
        if not A:
            break
        B = move_permutation_towards(B, A, debug)
        swaps += 1
        if matrix is not None and (results is not None or verbosity > 0):
            full = B + common
            mag = float(sum_after_2_to_4(_t(matrix).float()[:, full].contiguous()))
            for i, target in enumerate(magnitude_targets or ()):
                if abs(target - mag) < abs(target - results[i][0]):
                    results[i] = (mag, full)
            if verbosity > 0:
                print(f"swap {swaps:>4} {mag:>15.3f}")
        if swaps > limit:
            break
    return swaps, results


# ---- incremental map stages: the pieces of the reference's greedy loops, for callers that drive the loop themselves ------------------
# `_greedy` above rescans every stripe group per round on the device; these keep a score table between rounds and rescore only the
# entries whose stripes changed (reference exhaustive_search.py:36-72, 209-372 and channel_swap.py:72-206). Scoring goes through the
# same batched `_score_groups` (ab_stripe_search on a GPU) instead of one search_matrix / try_swap call per entry.
def generate_unique_combinations(built_permutation, remaining_columns, full_permutation_list, group_width: int = 4):
    """Append to ``full_permutation_list`` every canonical arrangement that extends ``built_permutation`` with ``remaining_columns``
    (ascending): columns ascend inside a group and a group opens with the smallest column not placed yet. The two argument lists are
    left as they were (exhaustive_search.py:36-72; iterative here, no recursion depth limit)."""
    stack = [(list(built_permutation), list(remaining_columns))]
    while stack:
        built, rest = stack.pop()
        if not rest:
            full_permutation_list.append(np.asarray(built))
            continue
        if len(built) % group_width == 0:
            # only the smallest unused column may open a group, and it must follow the opener of the group before
            nxt = [0] if all(v in built for v in range(rest[0])) and (not built or rest[0] > built[-group_width]) else []
        else:
            nxt = [i for i, c in enumerate(rest) if c > built[-1]]
        for i in reversed(nxt):                                   # reversed: the stack pops them in ascending order
            stack.append((built + [rest[i]], rest[:i] + rest[i + 1:]))
    return full_permutation_list


_stripe_group_cache: dict = {}


def _ordered_stripe_groups(num_stripes: int, window: int):
    key = (num_stripes, window)
    if key not in _stripe_group_cache:
        _stripe_group_cache[key] = list(itertools.combinations(range(num_stripes), window))
    return _stripe_group_cache[key]


def build_stripe_map(matrix, group_width, window_size, stripe_map, stripe_ids, perm_map, used_stripes):
    """Score table of the exhaustive strategy: one entry per group of ``window_size / group_width`` stripes = (best improvement any
    arrangement of its columns gives, that arrangement). Entries are created on the first call; afterwards only groups containing a
    stripe of ``used_stripes`` are rescored. -> (stripe_map, stripe_ids, perm_map), updated in place (exhaustive_search.py:209-288)."""
    if group_width != 4:
        raise NotImplementedError("scoring is 2:4 specific: group_width must be 4")
    C = matrix.shape[1]
    S = int(window_size) // group_width
    assert C % group_width == 0 and 2 <= S <= C // group_width
    groups = _ordered_stripe_groups(C // group_width, S)
    used = set(int(s) for s in used_stripes)
    stale = []
    for i, sg in enumerate(groups):
        if i >= len(stripe_map):
            stripe_ids.append(list(sg))
            stripe_map.append(0.0)
            perm_map.append(list(range(group_width * S)))
            stale.append(i)
        elif used.intersection(sg):
            stale.append(i)
    if stale:
        m = _t(matrix).detach().float().contiguous()
        cands = generate_all_unique_combinations(group_width * S, group_width)
        imp, idx = _score_groups(m, torch.tensor([groups[i] for i in stale], dtype=torch.int32, device=m.device), cands)
        for i, v, k in zip(stale, imp.tolist(), idx.tolist()):
            stripe_map[i] = float(v)
            perm_map[i] = [int(c) for c in cands[k if v > 0 else 0]]
    return stripe_map, stripe_ids, perm_map


sm_perturbations = 0          # escape moves spent / allowed by use_stripe_map (module state, as in the reference)
sm_perturbation_limit = 0


def use_stripe_map(matrix, group_width, stripe_map, stripe_ids, perm_map, permutation):
    """Apply the table of :func:`build_stripe_map` greedily, best improvement first, skipping groups that share a stripe with one already
    applied this round. ``matrix`` (numpy or tensor) is permuted in place. With nothing left to gain, and while
    ``sm_perturbations < sm_perturbation_limit``, one random group is applied with two of its columns exchanged across its halves.
    -> (matrix, groups applied, stripe_map, stripe_ids, stripes whose content changed, summed improvement, permutation)
    (exhaustive_search.py:296-372)."""
    global sm_perturbations
    eps = float(np.finfo(np.float16).tiny) * 5.0
    order = np.argsort(-np.asarray(stripe_map, dtype=np.float64), kind="stable")
    touched: set = set()
    changed_stripes: list = []
    applied, gain = 0, 0.0
    for gid in order:
        gid = int(gid)
        arrangement = list(perm_map[gid])
        if stripe_map[gid] <= eps:
            if touched or sm_perturbations >= sm_perturbation_limit:
                break
            sm_perturbations += 1
            gid = int(order[np.random.randint(len(order))])
            arrangement = list(perm_map[gid])
            half = len(arrangement) // 2
            a, b = np.random.randint(half), half + np.random.randint(half)
            arrangement[a], arrangement[b] = arrangement[b], arrangement[a]
        sg = stripe_ids[gid]
        if touched.intersection(sg):
            continue
        touched.update(sg)
        cols = [s * group_width + k for s in sg for k in range(group_width)]
        src = [cols[c] for c in arrangement]
        matrix[..., cols] = matrix[..., src]
        permutation = apply_stripe_group_permutation(arrangement, sg, group_width, permutation)
        for k, s in enumerate(sg):
            blk = arrangement[k * group_width:(k + 1) * group_width]
            # a stripe keeps its content when it received one whole source stripe in order
            if blk[0] % group_width or any(blk[j] != blk[0] + j for j in range(1, group_width)):
                changed_stripes.append(s)
        gain += stripe_map[gid]
        applied += 1
    return matrix, applied, stripe_map, stripe_ids, changed_stripes, gain, permutation


def compute_swap_map(matrix, used_stripes):
    """{(col0, col1): improvement of exchanging the two columns} for the 16 swaps of every stripe pair that contains a stripe of
    ``used_stripes``, all pairs scored in one batch (channel_swap.py:72-90). Runs on the CPU too (the reference asserts a GPU)."""
    pairs = build_stripe_pairs(matrix, used_stripes)
    out: dict = {}
    if len(pairs) == 0:
        return out
    m = _t(matrix).detach().float().contiguous()
    a = m.abs()
    R = a.shape[0]
    kept = _stripe_sums(m)                                                              # [C/4]
    p = torch.as_tensor(pairs.astype(np.int64), device=m.device)
    stripes = a.view(R, -1, 4)
    imp = torch.empty(len(pairs), 16, device=m.device)
    step = max(1, (1 << 22) // max(1, R * 4))                                           # bound the [R, pairs, 4] temporaries
    for c0 in range(0, len(pairs), step):
        q = p[c0:c0 + step]
        left, right = stripes[:, q[:, 0]], stripes[:, q[:, 1]]                           # [R, P, 4]
        base = kept[q[:, 0]] + kept[q[:, 1]]
        for i in range(4):
            for j in range(4):
                l2, r2 = left.clone(), right.clone()
                l2[..., i], r2[..., j] = right[..., j], left[..., i]
                imp[c0:c0 + step, i * 4 + j] = l2.topk(2, dim=-1).values.sum((0, 2)) + r2.topk(2, dim=-1).values.sum((0, 2)) - base
    imp[p[:, 0] == p[:, 1]] = 0.0                                                       # a stripe paired with itself: nothing moves
    imp_h = imp.tolist()
    for (s0, s1), row in zip(pairs.tolist(), imp_h):
        for k in range(16):
            out[stripes_and_swap_idx_to_columns(s0, s1, k)] = row[k]
    return out


def build_swap_map(matrix, swap_map, swap_ids, used_stripes, verbosity=0):
    """Score table of the channel-swap strategy: one entry per column pair (src < dst) from two different stripes, in row-major order.
    First call (empty ``swap_map``) scores everything, later calls only pairs touching ``used_stripes``. -> (swap_map, swap_ids),
    updated in place (channel_swap.py:94-138)."""
    C = matrix.shape[1]
    fresh = len(swap_map) == 0
    used = list(range(C // 4)) if fresh else sorted(set(int(s) for s in used_stripes))
    scores = compute_swap_map(matrix, used)
    used_set = set(used)
    idx = updates = 0
    for src in range(C - 1):
        for dst in range(src + 1, C):
            if src // 4 == dst // 4:
                continue
            if fresh or src // 4 in used_set or dst // 4 in used_set:
                v = float(scores[(src, dst)])
                if idx >= len(swap_map):
                    swap_map.append(v)
                    swap_ids.append((src, dst))
                else:
                    swap_map[idx], swap_ids[idx] = v, (src, dst)
                updates += 1
            idx += 1
    if verbosity > 15:
        print(f"\tupdated {updates} map entries")
    return swap_map, swap_ids


def use_swap_map(matrix, swap_map, swap_ids, threshold, used_escape_attempts, escape_attempts, permutation, verbosity=0):
    """Apply the table of :func:`build_swap_map`: swaps in order of benefit while the benefit stays above ``threshold`` x the best one
    (clamped to [1e-4, 1]), at most one swap per stripe per round; once converged, up to ``escape_attempts`` random swaps.
    ``matrix`` and ``permutation`` are modified in place. -> (matrix, swaps, swap_map, swap_ids, used_stripes, improvement,
    used_escape_attempts, permutation) (channel_swap.py:141-206)."""
    scores = np.asarray(swap_map, dtype=np.float64)
    order = np.argsort(-scores, kind="stable")
    floor = min(max(float(scores[order[0]]) * threshold, 1e-4), 1.0)
    used_stripes: list = []
    swaps, gain = 0, 0.0
    for sid in order:
        sid = int(sid)
        if scores[sid] < floor:
            if used_stripes or used_escape_attempts >= escape_attempts:
                break
            sid = int(order[np.random.randint(len(order))])
            used_escape_attempts += 1
            if verbosity > 15:
                print(f"converged, escape attempt #{used_escape_attempts}: swapping columns {swap_ids[sid]}")
        src, dst = swap_ids[sid]
        if src // 4 in used_stripes or dst // 4 in used_stripes:
            continue
        used_stripes += [src // 4, dst // 4]
        matrix[..., [src, dst]] = matrix[..., [dst, src]]
        permutation[src], permutation[dst] = permutation[dst], permutation[src]
        gain += float(scores[sid])
        swaps += 1
    return matrix, swaps, swap_map, swap_ids, used_stripes, gain, used_escape_attempts, permutation
