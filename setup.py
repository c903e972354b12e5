"""Packaging for sparkdl-b200: installs the ``sparkdl`` API facade and the ``spark_deep_learning_b200`` runtime and
compiles the native libraries in-tree (g++ for the engine, nvcc -gencode arch=compute_100a,code=sm_100a for the
kernels; nvcc cross-compiles without a GPU)."""
import os
import sys

from setuptools import find_packages, setup
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

# `python setup.py ... --with-tests` also packages the test-suite (same switch as the reference's setup.py:13-16)
WITH_TESTS = "--with-tests" in sys.argv
if WITH_TESTS:
    sys.argv.remove("--with-tests")
PACKAGES = ["sparkdl", "sparkdl.*", "spark_deep_learning_b200", "spark_deep_learning_b200.*"] + \
    (["tests", "tests.*"] if WITH_TESTS else [])


class BuildNative(build_py):
    def run(self):
        from spark_deep_learning_b200 import build as native
        native.build_core(verbose=True)
        if native.nvcc_path() is not None:
            native.build_cuda(verbose=True)
        else:
            print("nvcc not found: skipping libsparkdl_cuda.so (GPU collectives unavailable)")
        build_py.run(self)


setup(
    name="sparkdl-b200",
    version="0.1.0",
    description="Blackwell-native HorovodRunner, collectives and xgboost estimators with the sparkdl API",
    packages=find_packages(include=PACKAGES),
    package_data={"spark_deep_learning_b200": ["*.so", "csrc/core/*", "csrc/cuda/*"]},
    python_requires=">=3.9",
    install_requires=["cloudpickle", "numpy", "torch"],
    extras_require={"xgboost": ["pandas"], "test": ["pytest", "pandas"]},
    cmdclass={"build_py": BuildNative},
    platforms=["Linux"],
    license="Apache-2.0",
)
