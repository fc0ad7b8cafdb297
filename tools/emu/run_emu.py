#!/usr/bin/env python
"""Run the `-m gpu` parity tests against the CPU fiber EMULATION build of the kernels (tools/emu/_build/
libj2kgpu_emu.so: the same .cu sources compiled by g++ against a cuda_runtime.h shim).

Development aid for a container without a GPU: it finds indexing / halo / barrier bugs before GPU minutes are
spent.  It proves nothing about the product -- parity claims come only from `pytest -m gpu` on a B200.  The
package never loads the emulation library by itself; this script patches LIB_PATH of the already imported
module and strips the "no CUDA device" skip that tests/conftest.py adds.

usage:  python tools/emu/run_emu.py [pytest args...]      e.g.  -k "whole_path and 96" -x
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EMU_LIB = os.environ.get("J2K_EMU_LIB") or os.path.join(HERE, "_build", "libj2kgpu_emu.so")


class EmuPlugin:
    def pytest_configure(self, config):
        sys.path.insert(0, ROOT)
        from __graft_entry__ import load_package
        mod = load_package()
        mod.LIB_PATH = EMU_LIB
        mod._lib = None

    def pytest_collection_modifyitems(self, config, items):
        for item in items:
            item.own_markers = [m for m in item.own_markers if m.name != "skip"]


EmuPlugin.pytest_collection_modifyitems = __import__("pytest").hookimpl(trylast=True)(EmuPlugin.pytest_collection_modifyitems)

if __name__ == "__main__":
    import pytest
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE])
    os.chdir(ROOT)
    args = sys.argv[1:] or ["-x", "-q"]
    paths = [] if any(a.startswith("tests") for a in args) else ["tests"]       # explicit test paths replace the whole suite
    sys.exit(pytest.main(paths + ["-m", "gpu", "-p", "no:cacheprovider"] + args, plugins=[EmuPlugin()]))
