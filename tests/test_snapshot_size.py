"""The gpurun / driver snapshot of the repo must stay far below the 512 MiB cap (round 1 was refused for exceeding it).

Walks the tree the way the snapshot does: everything except `.git/`, `gpurun_out/` and the patterns in `.gpurunignore`.
"""
import fnmatch
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIMIT_MIB = 400


def _patterns():
    pats = [".git/", "gpurun_out/"]
    with open(os.path.join(ROOT, ".gpurunignore")) as f:
        for line in f:
            line = line.strip()
            if line and not line.startswith("#"):
                pats.append(line)
    return pats


def _ignored(rel, is_dir, pats):
    name = os.path.basename(rel)
    for p in pats:
        if p.endswith("/"):
            d = p.rstrip("/")
            if is_dir and (rel == d or name == d and "/" not in d):
                return True
        elif fnmatch.fnmatch(name, p) or fnmatch.fnmatch(rel, p):
            return True
    return False


def snapshot_bytes():
    pats = _patterns()
    total, biggest = 0, []
    for dirpath, dirnames, filenames in os.walk(ROOT):
        rel_dir = os.path.relpath(dirpath, ROOT)
        rel_dir = "" if rel_dir == "." else rel_dir
        dirnames[:] = [d for d in dirnames if not _ignored(os.path.join(rel_dir, d) if rel_dir else d, True, pats)]
        for f in filenames:
            rel = os.path.join(rel_dir, f) if rel_dir else f
            if _ignored(rel, False, pats):
                continue
            try:
                n = os.path.getsize(os.path.join(dirpath, f))
            except OSError:
                continue
            total += n
            biggest.append((n, rel))
    biggest.sort(reverse=True)
    return total, biggest[:8]


def test_snapshot_stays_under_the_cap():
    total, biggest = snapshot_bytes()
    mib = total / 2**20
    assert mib < LIMIT_MIB, f"snapshot is {mib:.0f} MiB (limit {LIMIT_MIB}); largest: {biggest}"


def test_no_checkpoints_or_profiler_captures_tracked():
    import subprocess
    out = subprocess.run(["git", "ls-files"], cwd=ROOT, capture_output=True, text=True).stdout.split("\n")
    bad = [f for f in out if f.endswith((".pth.tar", ".pt", ".pth", ".ckpt", ".ncu-rep", ".so", ".o"))]
    assert not bad, bad
    big = [f for f in out if f and os.path.exists(os.path.join(ROOT, f)) and os.path.getsize(os.path.join(ROOT, f)) > 5 * 2**20]
    assert not big, big


if __name__ == "__main__":
    t, b = snapshot_bytes()
    print(f"{t / 2**20:.1f} MiB", b)
