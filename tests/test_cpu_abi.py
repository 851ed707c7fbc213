"""The ctypes layer has no compiler to catch a drifted signature: every ``_lib.declare(name, spec)`` must match the ``AB_API``
prototype of ``name`` in csrc/ argument for argument (pointer / int32 / int64 / float / double)."""
import glob
import os
import re
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_prototypes():
    protos = {}
    for f in glob.glob(os.path.join(ROOT, "apex_b200/csrc/**/*.cu"), recursive=True) + glob.glob(os.path.join(ROOT, "apex_b200/csrc/*.cpp")):
        for m in re.finditer(r"AB_API\s+([\w\s\*]+?)\s+(\w+)\s*\(([^)]*)\)\s*\{", Path(f).read_text(), re.S):
            codes = []
            for a in (x.strip() for x in m.group(3).replace("\n", " ").split(",") if x.strip()):
                a = re.sub(r"\s+", " ", a)
                if "*" in a or "cudaStream_t" in a:
                    codes.append("p")
                elif re.match(r"(const )?(unsigned )?long long", a) or "int64_t" in a or a.startswith("size_t"):
                    codes.append("l")
                elif re.match(r"(const )?(unsigned )?(int|unsigned|uint32_t)\b", a):
                    codes.append("i")
                elif re.match(r"(const )?float", a):
                    codes.append("f")
                elif a.startswith("double"):
                    codes.append("d")
                else:
                    codes.append("?" + a)
            protos[m.group(2)] = codes
    return protos


def _declarations():
    decl = {}
    for f in glob.glob(os.path.join(ROOT, "apex_b200/**/*.py"), recursive=True):
        src = Path(f).read_text().replace('_T + "', '"p i i i i ').replace('_T+"', '"p i i i i ')
        for m in re.finditer(r'declare\(\s*"(\w+)"\s*,\s*"([^"]*)"', src):
            decl[m.group(1)] = (m.group(2).split(), os.path.relpath(f, ROOT))
    return decl


def test_every_declared_signature_matches_its_c_prototype():
    protos, decl = _c_prototypes(), _declarations()
    assert len(decl) >= 40
    problems = []
    for name, (spec, where) in decl.items():
        if name not in protos:
            problems.append(f"{name} ({where}): no AB_API prototype")
        elif protos[name] != spec:
            problems.append(f"{name} ({where}): python {' '.join(spec)}  !=  C {' '.join(protos[name])}")
    assert not problems, "\n".join(problems)


def test_built_library_exports_every_declared_symbol():
    """A stale in-tree ``_kernels.so`` (sources changed, library not rebuilt) would only fail on the GPU box: catch it here. Symbols that
    live in csrc/experimental are only required when the library was built with APEX_B200_EXPERIMENTAL=1."""
    import ctypes

    import pytest

    so = os.path.join(ROOT, "apex_b200", "_kernels.so")
    if not os.path.exists(so):
        pytest.skip("library not built")
    lib = ctypes.CDLL(so)
    experimental = set()
    for f in glob.glob(os.path.join(ROOT, "apex_b200/csrc/experimental/*.cu")):
        experimental.update(re.findall(r"AB_API\s+[\w\s\*]+?\s+(\w+)\s*\(", Path(f).read_text()))
    names = set(_declarations())
    for f in glob.glob(os.path.join(ROOT, "apex_b200/**/*.py"), recursive=True):
        names.update(re.findall(r'raw_fn\(\s*"(\w+)"', Path(f).read_text()))
    has_experimental = any(hasattr(lib, n) for n in experimental)
    missing = [n for n in sorted(names) if not hasattr(lib, n) and (has_experimental or n not in experimental)]
    assert not missing, f"rebuild the library (python -m apex_b200._build): {missing}"


def test_kernel_call_sites_pass_as_many_arguments_as_their_signature_declares():
    """ctypes only complains about a wrong argument count when the call executes, and most call sites are CUDA-only branches: count the
    arguments of every ``_lib.fn("name")(...)`` statically (``*table.head()`` expands to 5; calls with other star-arguments are skipped)."""
    import ast

    specs = {name: len(spec) for name, (spec, _) in _declarations().items()}
    for f in glob.glob(os.path.join(ROOT, "benchmarks/*.py")):
        for m in re.finditer(r'declare\(\s*"(\w+)"\s*,\s*"([^"]*)"', Path(f).read_text()):
            specs[m.group(1)] = len(m.group(2).split())
    problems, checked = [], 0
    files = glob.glob(os.path.join(ROOT, "apex_b200/**/*.py"), recursive=True) + glob.glob(os.path.join(ROOT, "benchmarks/*.py")) + \
        glob.glob(os.path.join(ROOT, "tests/*.py"))
    for f in files:
        for node in ast.walk(ast.parse(Path(f).read_text())):
            if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Call)):
                continue
            inner = node.func
            if not (isinstance(inner.func, ast.Attribute) and inner.func.attr == "fn" and inner.args and isinstance(inner.args[0], ast.Constant)
                    and isinstance(inner.args[0].value, str)):
                continue
            name, n, unknown = inner.args[0].value, 0, False
            for a in node.args:
                if not isinstance(a, ast.Starred):
                    n += 1
                elif isinstance(a.value, ast.Call) and isinstance(a.value.func, ast.Attribute) and a.value.func.attr == "head":
                    n += 5
                elif isinstance(a.value, (ast.Tuple, ast.List)):
                    n += len(a.value.elts)
                else:
                    unknown = True
            if unknown or name not in specs:
                continue
            checked += 1
            if n != specs[name]:
                problems.append(f"{os.path.relpath(f, ROOT)}:{node.lineno}: {name} called with {n} arguments, declared with {specs[name]}")
    assert checked >= 50 and not problems, "\n".join(problems)


def test_cuda_branches_execute_against_type_checking_stubs():
    """tests/_fake_cuda_dryrun.py: the python side of the CUDA-only branches (marshalling, autograd plumbing, state_dict round trips) runs on a
    machine without a GPU with every kernel replaced by a stub that validates argument count and types. Own process: it patches torch globally."""
    import subprocess
    import sys

    r = subprocess.run([sys.executable, str(Path(__file__).parent / "_fake_cuda_dryrun.py")], capture_output=True, text=True, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith(("ok", "FAIL"))]
    assert r.returncode == 0 and lines and not [ln for ln in lines if ln.startswith("FAIL")], r.stdout[-4000:] + r.stderr[-2000:]
