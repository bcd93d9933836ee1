"""Builds libhalo_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "lib", "libhalo_b200.so")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = []
    for d in (os.path.join(HERE, "csrc"), os.path.join(HERE, "..", "include")):
        srcs += [os.path.join(d, f) for f in os.listdir(d)]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, jobs=None):
    if force or _stale():
        jobs = jobs or os.cpu_count() or 4
        subprocess.check_call(["make", "-C", HERE, f"-j{jobs}", "all"] + (["-B"] if force else []))
    return LIB
