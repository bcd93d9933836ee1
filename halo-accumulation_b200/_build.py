"""Builds libhalo_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# HALO_B200_CURVE=vesta binds the Vesta build of the same sources (include/halo_b200.h: halo_curve_name)
CURVE = os.environ.get("HALO_B200_CURVE", "pallas")
if CURVE not in ("pallas", "vesta"):
    raise RuntimeError(f"HALO_B200_CURVE={CURVE!r}: expected 'pallas' or 'vesta'")
_SUFFIX = "" if CURVE == "pallas" else "_vesta"
LIB = os.path.join(HERE, "lib", f"libhalo_b200{_SUFFIX}.so")
HOST_LIB = os.path.join(HERE, "lib", f"libhalo_host{_SUFFIX}.so")


def _stale():
    if not os.path.exists(LIB):
        return True
    libs = [os.path.join(HERE, "lib", f) for f in ("libhalo_b200.so", "libhalo_b200_vesta.so", "libhalo_host.so", "libhalo_host_vesta.so")]
    if not all(os.path.exists(l) for l in libs):
        return True
    t = min(os.path.getmtime(l) for l in libs)
    srcs = []
    for d in (os.path.join(HERE, "csrc"), os.path.join(HERE, "host"), os.path.join(HERE, "..", "include")):
        srcs += [os.path.join(d, f) for f in os.listdir(d)]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, jobs=None):
    if force or _stale():
        jobs = jobs or os.cpu_count() or 4
        subprocess.check_call(["make", "-C", HERE, f"-j{jobs}", "all"] + (["-B"] if force else []))
    return LIB
