"""Build hook: compile the sm_100a kernels and the host runtime in-tree before the package is collected, so that a wheel
carries ``petals_b200/_native/*.so`` (``nvcc`` cross-compiles without a GPU; without ``nvcc`` the sources still ship and
``petals_b200.ops.native`` builds them on first use on a machine that has it)."""
import importlib.util
import os
import sys

from setuptools import setup
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_builder():
    spec = importlib.util.spec_from_file_location("_pb_build", os.path.join(HERE, "petals_b200", "_build.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_pb_build"] = mod
    spec.loader.exec_module(mod)
    return mod


class build_py_with_native(build_py):
    def run(self):
        try:
            _load_builder().build(verbose=True)
        except Exception as e:  # noqa: BLE001 - a source-only install is still usable where nvcc exists
            print(f"warning: native libraries were not built ({e}); they will be built on first use", file=sys.stderr)
        super().run()


setup(cmdclass={"build_py": build_py_with_native})
