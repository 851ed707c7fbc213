"""Packaging for apex_b200. The native libraries are built by apex_b200/_build.py (nvcc -gencode arch=compute_100a,code=sm_100a for
every csrc/*.cu -> apex_b200/_kernels.so, g++ for the pybind host runtime -> apex_b200/_C.so); this file only hooks that driver into
setuptools so `pip install --no-build-isolation -e .` / `python setup.py build_ext --inplace` work.

The reference gates each of its 31 extensions behind an APEX_* flag (setup.py:24-52); here there is one library and one switch:
APEX_B200_EXPERIMENTAL=1 adds csrc/experimental/*.cu."""
import importlib.util
import os

from setuptools import Command, find_packages, setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))


def _build_native():
    spec = importlib.util.spec_from_file_location("apex_b200_build", os.path.join(ROOT, "apex_b200", "_build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_all(verbose=True)


class BuildNative(Command):
    description = "compile the sm_100a kernels and the pybind host runtime in-tree"
    user_options = [("inplace", "i", "accepted for compatibility (the build is always in-tree)")]

    def initialize_options(self):
        self.inplace = True

    def finalize_options(self):
        pass

    def run(self):
        _build_native()


class BuildPyWithNative(build_py):
    def run(self):
        _build_native()
        super().run()


setup(
    name="apex_b200",
    version="0.1.0",
    description="Blackwell-native (B200, sm_100a) counterpart of NVIDIA/apex: fused optimizers, norms, tcgen05 GEMMs, in-kernel collectives",
    packages=find_packages(include=["apex_b200", "apex_b200.*"]),
    package_data={"apex_b200": ["_kernels.so", "_C.so", "csrc/*", "csrc/experimental/*"]},
    python_requires=">=3.10",
    install_requires=["torch"],
    cmdclass={"build_ext": BuildNative, "build_py": BuildPyWithNative},
)
