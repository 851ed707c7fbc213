"""The ctypes layer has no compiler to catch a drifted signature: every ``_lib.declare(name, spec)`` must match the ``AB_API``
prototype of ``name`` in csrc/ argument for argument (pointer / int32 / int64 / float / double)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_prototypes():
    protos = {}
    for f in glob.glob(os.path.join(ROOT, "apex_b200/csrc/**/*.cu"), recursive=True) + glob.glob(os.path.join(ROOT, "apex_b200/csrc/*.cpp")):
        for m in re.finditer(r"AB_API\s+([\w\s\*]+?)\s+(\w+)\s*\(([^)]*)\)\s*\{", open(f).read(), re.S):
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
        src = open(f).read().replace('_T + "', '"p i i i i ').replace('_T+"', '"p i i i i ')
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
