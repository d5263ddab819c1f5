"""pip install . : builds the native library with `make` (nvcc for sm_100a when available, NO_CUDA=1 otherwise) and
installs the mlsl_b200 package with the library inside it.  The source tree itself needs no installation: the tests,
the benchmarks and __graft_entry__ use it in place."""
import os
import shutil
import subprocess

from setuptools import setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))


class BuildWithMake(build_py):
    def run(self):
        env = dict(os.environ)
        args = ["make", "-C", ROOT, "-j%d" % (os.cpu_count() or 4), "mlsl_b200/lib/libmlsl_b200.so"]
        if not shutil.which(env.get("NVCC", "/usr/local/cuda/bin/nvcc")):
            args.append("NO_CUDA=1")
        subprocess.check_call(args, env=env)
        super().run()


setup(
    name="mlsl_b200",
    version="2026.1",
    description="Blackwell-native deep-learning collective library with the capabilities of Intel MLSL",
    packages=["mlsl_b200", "mlsl_b200.models", "mlsl_b200.ops", "mlsl_b200.parallel", "mlsl_b200.utils"],
    package_data={"mlsl_b200": ["lib/libmlsl_b200.so"]},
    python_requires=">=3.9",
    cmdclass={"build_py": BuildWithMake},
)
