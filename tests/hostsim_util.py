"""builds (g++) and loads the sequential HOST SIMULATION of the CUDA kernel source
(tests/hostsim/njode_hostsim.cpp) -- test infrastructure for the CPU-only suite."""
import os
import subprocess

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "hostsim", "njode_hostsim.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "_hostsim")
OUT = os.path.join(OUT_DIR, "libnjode_hostsim.so")
DEPS = [SRC] + [os.path.join(ROOT, "njode_b200", "csrc", f) for f in ("njode_core.cuh", "njode_hash.cuh", "njode_seg.cuh", "njode_path.cuh", "njode_tpn.cuh", "njode_plan.h")] + \
       [os.path.join(ROOT, "include", "njode_b200.h")]

_runner = None


def build():
    os.makedirs(OUT_DIR, exist_ok=True)
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                           "-o", OUT, SRC])
    return OUT


def runner():
    global _runner
    if _runner is None:
        from njode_b200 import _ext
        _runner = _ext.Runner(_ext.Lib(build()), torch.device("cpu"))
    return _runner


_saved = []


def install():
    """route njode_b200 to the host simulation: replaces ``_ext.cuda_runner`` (the one place the product obtains its
    runner) for the duration of a test.  The product has no such switch of its own."""
    from njode_b200 import _ext
    _saved.append(_ext.cuda_runner)
    r = runner()
    _ext.cuda_runner = lambda device: r


def uninstall():
    from njode_b200 import _ext
    if _saved:
        _ext.cuda_runner = _saved[0]
        del _saved[:]


def plan_kind(model, pb, which="fwd"):
    """kernel family the planner picks: set of {"seg", "segtpn", "path", "pathstat", "pipe", "tpn", "tpn_fwd"} (see njode_hostsim_plan_kind)"""
    import ctypes as C
    dll = runner().lib.dll
    mt = model._model_struct(0)
    bits = dll.njode_hostsim_plan_kind(C.byref(mt), C.byref(getattr(pb, which)))
    assert bits >= 0, bits
    return {n for i, n in enumerate(("seg", "segtpn", "path", "pathstat", "pipe", "tpn", "tpn_fwd")) if bits >> i & 1}
