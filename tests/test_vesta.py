"""The Vesta instantiation: libhalo_b200_vesta.so / libhalo_host_vesta.so are the same sources as the
Pallas libraries compiled with -DHALO_CURVE_VESTA (coordinate and scalar field swapped; own GLV constants from
tools/gen_glv_consts.py).  The reference has no Vesta code path, constants or tests, so parity here is
GPU == oracle (oracle/liboracle_vesta.so, the same restatement with the moduli swapped) == pyref (independent big-int
arithmetic): the whole parity suite is re-run with HALO_B200_CURVE=vesta, minus the assertions pinned to consts.rs.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PALLAS_P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
PALLAS_R = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001


def _rerun(files, marker, timeout):
    env = dict(os.environ, HALO_B200_CURVE="vesta")
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", marker, "-p", "no:cacheprovider", *files]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    return tail


def test_vesta_cpu_suite():
    """Oracle (Vesta build) against pyref, the host-compiled field core / group law / GLV split of the Vesta libraries,
    exported symbols, and the Vesta known-answer file."""
    import halo_accumulation_b200 as H

    H.build()
    tail = _rerun(["tests/test_oracle_golden.py", "tests/test_host_build.py"], "not gpu", 900)
    assert " passed" in tail and "failed" not in tail


def test_glv_constants_are_derived():
    """tools/gen_glv_consts.py reproduces the Pallas constants of csrc/glv.cuh and its Vesta output is what glv.cuh carries."""
    tool = os.path.join(ROOT, "halo-accumulation_b200", "tools", "gen_glv_consts.py")
    src = open(os.path.join(ROOT, "halo-accumulation_b200", "csrc", "glv.cuh")).read().replace(" ", "")
    out = subprocess.run([sys.executable, tool, "pallas"], capture_output=True, text=True, check=True).stdout
    assert "# matches csrc/glv.cuh (Pallas)" in out
    out = subprocess.run([sys.executable, tool, "vesta"], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines() if "=" in l and not l.startswith("#")]
    assert len(lines) == 6
    for l in lines:
        words = [w.rstrip("ull").rstrip("u") for w in l.split("=", 1)[1].strip(" {}").replace(" ", "").split(",")]
        words = [w for w in words if int(w, 16) > 1]  # the sources abbreviate 0 / 1 limbs
        assert all(w in src for w in words), l


def _open(path):
    lib = C.CDLL(path)  # RTLD_LOCAL: the two libraries export the same names
    lib.halo_curve_name.restype = C.c_char_p
    lib.halo_ctx_create.argtypes = [C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.halo_ctx_destroy.argtypes = [C.c_void_p]
    lib.halo_ctx_destroy.restype = None
    return lib


def test_both_libraries_load_side_by_side():
    import halo_accumulation_b200 as H

    H.build()
    d = os.path.join(ROOT, "halo-accumulation_b200", "lib")
    a, b = _open(os.path.join(d, "libhalo_b200.so")), _open(os.path.join(d, "libhalo_b200_vesta.so"))
    assert a.halo_curve_name() == b"pallas" and b.halo_curve_name() == b"vesta"


def _aff_ints(row, p):
    rinv = pow(1 << 256, -1, p)
    v = [sum(int(x) << (64 * i) for i, x in enumerate(row[k:k + 4])) * rinv % p for k in (0, 4)]
    return v[0], v[1]


def _jac_to_affine(j, p):
    rinv = pow(1 << 256, -1, p)
    x, y, z = (sum(int(v) << (64 * i) for i, v in enumerate(j[k:k + 4])) * rinv % p for k in (0, 4, 8))
    zi = pow(z, -1, p)
    return x * zi * zi % p, y * zi * zi * zi % p


def _add(a, b, p):
    if a is None:
        return b
    (x1, y1), (x2, y2) = a, b
    lam = (3 * x1 * x1 * pow(2 * y1, -1, p) if a == b else (y2 - y1) * pow(x2 - x1, -1, p)) % p
    x3 = (lam * lam - x1 - x2) % p
    return x3, (lam * (x1 - x3) - y1) % p


@pytest.mark.gpu
def test_both_curves_in_one_process():
    """A process that needs both halves of the cycle dlopens both libraries: each derives its own generators (on its own
    curve) and computes an MSM that a big-int restatement confirms; neither library's internals bind to the other's."""
    d = os.path.join(ROOT, "halo-accumulation_b200", "lib")
    n = 64
    for name, p, r in (("libhalo_b200.so", PALLAS_P, PALLAS_R), ("libhalo_b200_vesta.so", PALLAS_R, PALLAS_P),
                       ("libhalo_b200.so", PALLAS_P, PALLAS_R)):
        lib = _open(os.path.join(d, name))
        h = C.c_void_p()
        assert lib.halo_ctx_create(0, C.c_uint64(1024), C.byref(h)) == 0
        try:
            assert lib.halo_derive_generators(h, C.c_uint64(n)) == 0
            gs = np.zeros((n, 8), dtype=np.uint64)
            assert lib.halo_get_generators(h, C.c_uint64(0), C.c_uint64(n), gs.ctypes.data_as(C.POINTER(C.c_uint64))) == 0
            pts = [_aff_ints(g, p) for g in gs]
            assert all((y * y - x * x * x - 5) % p == 0 for x, y in pts)
            ks = [(7 * i + 1) % 5 + 1 for i in range(n)]  # small scalars: the expected sum by repeated affine addition
            sc = np.array([[(k * (1 << 256) % r >> (64 * i)) & (2**64 - 1) for i in range(4)] for k in ks], dtype=np.uint64)
            out = np.zeros(12, dtype=np.uint64)
            assert lib.halo_msm_gens(h, sc.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_uint64(0), C.c_uint64(n),
                                     out.ctypes.data_as(C.POINTER(C.c_uint64))) == 0
            exp = None
            for pt, k in zip(pts, ks):
                for _ in range(k):
                    exp = _add(exp, pt, p)
            assert _jac_to_affine(out, p) == exp, name
        finally:
            lib.halo_ctx_destroy(h)


@pytest.mark.gpu
def test_vesta_gpu_suite():
    """Field core, group law, generator derivation, every MSM mode, PCDL commit / open / check, ASDL prover / verifier /
    decider and the Vesta known-answer file on the GPU, bit-exact with the Vesta oracle."""
    tail = _rerun(["tests/test_gpu_core.py", "tests/test_gpu_pcdl.py"], "gpu", 1500)
    assert " passed" in tail and "failed" not in tail
