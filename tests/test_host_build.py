"""CPU tests of the product without a GPU: the shared libraries load, export every symbol the headers declare,
fail loudly without a device, and the shared field / curve / hash code (compiled as plain C++) agrees with the oracle."""
import ctypes as C
import os
import random
import re
import subprocess

import numpy as np
import pytest

from oracle import pyref as PR
from oracle.oracle import CURVE  # "pallas" unless HALO_B200_CURVE=vesta (tests/test_vesta.py re-runs this file that way)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libs():
    import halo_accumulation_b200 as H

    H.build()
    from halo_accumulation_b200 import _capi, _host

    return _capi.load(), _host.lib()


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(halo_[a-z0-9_]+)\s*\(", text)))


def test_cuda_library_exports_declared_symbols(libs):
    cuda, host = libs
    for name in _declared("halo_b200.h") + _declared("halo_b200_test.h"):
        assert hasattr(cuda, name), f"libhalo_b200.so does not export {name}"
    for name in _declared("halo_pcdl.h"):
        assert hasattr(host, name), f"libhalo_host.so does not export {name}"


def test_library_contains_sm100a_code():
    from halo_accumulation_b200 import _build

    out = subprocess.run(["cuobjdump", "-lelf", _build.LIB],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_device(libs):
    """The product path must fail loudly when no CUDA device is usable."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import halo_accumulation_b200 as H

    with pytest.raises(H.HaloError):
        H.Context(0, 1024)


def test_product_sources_do_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "halo-accumulation_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f in (), f"{f} mentions the oracle"


@pytest.fixture(scope="module")
def hostcheck():
    """The product's __host__ __device__ headers compiled by g++ (portable 32-bit limb path = the algorithm the
    device runs; and the 64-bit host path used by the transcript glue)."""
    d = os.path.join(ROOT, "tests", "hostcheck")
    libs = {}
    for tag, flags in (("portable", ["-DHALO_FP_FORCE_PORTABLE"]), ("host64", []), ("host64_c", ["-DHALO_FP_NO_X64_ASM"])):
        so = os.path.join(d, f"libhostcheck_{tag}{'' if CURVE == 'pallas' else '_vesta'}.so")
        flags = flags + (["-DHALO_CURVE_VESTA"] if CURVE == "vesta" else [])
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", *flags, "-o", so,
                               os.path.join(d, "hostcheck.cpp")])
        libs[tag] = C.CDLL(so)
    return libs


def _p32(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


@pytest.mark.parametrize("tag", ["portable", "host64", "host64_c"])
def test_shared_field_core_vs_python(hostcheck, oracle, tag):
    hc = hostcheck[tag]
    rnd = random.Random(3)
    for which, mod in ((0, PR.P), (1, PR.R)):
        edge = [0, 1, 2, mod - 1, mod - 2, (1 << 256) % mod, (1 << 255) % mod, 0xFFFFFFFF, 1 << 32, (1 << 254) % mod, mod >> 1]
        vals = edge + [rnd.randrange(mod) for _ in range(150)]
        rinv = pow(1 << 256, -1, mod)
        for i, a in enumerate(vals):
            for b in edge + [vals[(7 * i + 3) % len(vals)]]:
                A = np.array(oracle.int_to_limbs(a), dtype=np.uint64)
                B = np.array(oracle.int_to_limbs(b), dtype=np.uint64)
                r = np.zeros(4, dtype=np.uint64)
                hc.hc_fp_mul(which, _p32(A), _p32(B), _p32(r))
                assert oracle.limbs_to_int(r) == a * b * rinv % mod
                hc.hc_fp_addsub(which, 0, _p32(A), _p32(B), _p32(r))
                assert oracle.limbs_to_int(r) == (a + b) % mod
                hc.hc_fp_addsub(which, 1, _p32(A), _p32(B), _p32(r))
                assert oracle.limbs_to_int(r) == (a - b) % mod
        for a in vals[1:30]:
            A = np.array(oracle.int_to_limbs(a), dtype=np.uint64)
            r = np.zeros(4, dtype=np.uint64)
            hc.hc_fp_inv(which, _p32(A), _p32(r))
            assert oracle.limbs_to_int(r) == pow(a, -1, mod) * pow(1 << 256, 2, mod) % mod
            hc.hc_fp_canon(which, 1, _p32(A), _p32(r))
            assert oracle.limbs_to_int(r) == a * rinv % mod


@pytest.mark.parametrize("tag", ["portable", "host64"])
def test_division_step_inversion_vs_fermat(hostcheck, oracle, tag):
    """fp_inv (Bernstein-Yang division steps on signed 30-bit limbs, csrc/fp.cuh) against the Fermat power it replaced and
    against Python's pow(a, -1, p): edge values (0, 1, 2, p - 1, p - 2, powers of two around the limb boundaries, values with
    long runs of zero / one bits) and 20 000 random residues per field."""
    hc = hostcheck[tag]
    hc.hc_fp_inv_cross.restype = C.c_uint64
    rnd = random.Random(99)
    for which, mod in ((0, PR.P), (1, PR.R)):
        edge = [0, 1, 2, 3, mod - 1, mod - 2, (mod + 1) // 2, (1 << 256) % mod, (1 << 255) % mod, (1 << 254) - 1, 1 << 253]
        edge += [1 << k for k in (29, 30, 31, 32, 59, 60, 61, 89, 90, 119, 120, 149, 150, 239, 240, 241, 252)]
        edge += [(1 << k) - 1 for k in (30, 60, 90, 120, 150, 180, 210, 240, 254)] + [mod - (1 << k) for k in (30, 60, 120, 240)]
        vals = edge + [rnd.randrange(mod) for _ in range(20000)]
        A = np.array([oracle.int_to_limbs(v) for v in vals], dtype=np.uint64)
        assert hc.hc_fp_inv_cross(which, _p32(A), C.c_uint64(len(vals))) == 0
        r2 = pow(1 << 256, 2, mod)
        for a in edge[1:] + vals[len(edge):len(edge) + 200]:
            Aa = np.array(oracle.int_to_limbs(a), dtype=np.uint64)
            r = np.zeros(4, dtype=np.uint64)
            hc.hc_fp_inv(which, _p32(Aa), _p32(r))
            assert oracle.limbs_to_int(r) == pow(a, -1, mod) * r2 % mod


def test_host_field_implementations_agree(hostcheck, oracle):
    """The host field core exists three times: MULX / ADC chains in x86-64 inline assembly (what the library's host glue runs on
    a CPU with BMI2: Horner finish of every variable-base MSM, the verifier's scalar multiplications), the 64-bit C versions
    (other CPUs) and the portable 32-bit-limb multiplication (the device algorithm).  All must agree bit for bit -- 40 000
    random operand pairs per field, the edge values crossed with each other, results aliasing an operand, squares."""
    rnd = random.Random(2024)
    for tag in ("host64", "host64_c"):
        hc = hostcheck[tag]
        hc.hc_fp_impl_cross.restype = C.c_uint64
        for which, mod in ((0, PR.P), (1, PR.R)):
            edge = [0, 1, 2, mod - 1, mod - 2, (mod + 1) // 2, mod >> 1, (1 << 256) % mod, (1 << 255) % mod, (1 << 254) - 1, 1 << 253,
                    (1 << 64) - 1, 1 << 64, (1 << 128) - 1, 1 << 128, (1 << 192) - 1, 1 << 192, mod - (1 << 64), mod - (1 << 128),
                    mod - (1 << 192), (1 << 254) % mod, 0xFFFFFFFF, 1 << 32]
            pairs = [(a, b) for a in edge for b in edge] + [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(40000)]
            A = np.array([oracle.int_to_limbs(a) for a, _ in pairs], dtype=np.uint64)
            B = np.array([oracle.int_to_limbs(b) for _, b in pairs], dtype=np.uint64)
            assert hc.hc_fp_impl_cross(which, _p32(A), _p32(B), C.c_uint64(len(pairs))) == 0, (tag, which)
    # this suite must have exercised the assembly path where the CPU offers it (x86-64 with BMI2), and never in the C build
    assert hostcheck["host64_c"].hc_fp_uses_x64_asm() == 0
    import platform
    if platform.machine() == "x86_64" and "bmi2" in open("/proc/cpuinfo").read():
        assert hostcheck["host64"].hc_fp_uses_x64_asm() == 2


@pytest.mark.parametrize("tag", ["portable", "host64"])
def test_shared_group_law_vs_oracle(hostcheck, oracle, tag):
    hc, O = hostcheck[tag], oracle
    GS = O.derive_points(2, 32)
    aff = np.concatenate([GS[:10], GS[3:4], GS[3:4], GS[5:6], np.zeros((1, 8), dtype=np.uint64), GS[:10]])
    neg = np.zeros(len(aff), dtype=np.uint8)
    neg[12] = 1
    neg[-10:] = 1
    out = np.zeros(12, dtype=np.uint64)
    hc.hc_madd_chain(_p32(aff), neg.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint64(len(aff)), _p32(out))
    exp = None
    for a, ng in zip(aff, neg):
        if a.any():
            x, y = O.from_mont(a.reshape(2, 4), 0)
            exp = PR.pt_add(exp, (x, (-y) % PR.P if ng else y))
    assert O.pt_to_affine_ints(out) == exp
    quad = np.concatenate([GS[:1]] * 4)
    n4 = np.array([0, 0, 1, 1], dtype=np.uint8)
    hc.hc_madd_chain(_p32(quad), n4.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint64(4), _p32(out))
    assert O.pt_to_affine_ints(out) is None
    jacs = np.array([O.pt_mul(O.affine_to_jac(GS[i])[0], O.random_scalars(1, i)[0]) for i in range(8)])
    jacs2 = np.concatenate([jacs, jacs[:1], jacs[2:3]])
    hc.hc_add_chain(_p32(jacs2), C.c_uint64(len(jacs2)), 3, _p32(out))
    exp = None
    for j in jacs2:
        exp = PR.pt_add(exp, O.pt_to_affine_ints(j))
    assert O.pt_to_affine_ints(out) == PR.pt_mul(exp, 8)
    k = O.random_scalars(1, 77)[0]
    kc = np.zeros(4, dtype=np.uint64)
    O.lib().orc_fp_to_canon(1, O._p(k), O._p(kc))
    hc.hc_mul(_p32(jacs[2]), _p32(kc), _p32(out))
    assert O.pt_eq(out, O.pt_mul(jacs[2], k))
    # `Projective * Fr` of the host layer (GLV + joint sparse form, csrc/glv.cuh) against the oracle's double-and-add:
    # edge scalars (0, +-1, +-lambda-sized, powers of two around the 128-bit split), random scalars, infinity, a
    # non-normalised point
    rnd = random.Random(41)
    ks = [0, 1, 2, 3, PR.R - 1, PR.R - 2, (1 << 127), (1 << 128) - 1, 1 << 128, (1 << 128) + 1, (1 << 254), PR.R // 2, PR.R // 3]
    ks += [rnd.randrange(PR.R) for _ in range(60)]
    pts = [jacs[3], O.affine_to_jac(GS[0])[0], O.pt_from_affine_ints(None)]
    for i, kv in enumerate(ks):
        km = O.to_mont([kv])[0]
        pj = pts[i % 3] if i >= 13 else pts[0]
        hc.hc_mul_glv(_p32(np.ascontiguousarray(pj)), _p32(km), _p32(out))
        assert O.pt_eq(out, O.pt_mul(pj, km)), hex(kv)
    hc.hc_mul_glv(_p32(np.ascontiguousarray(pts[2])), _p32(O.to_mont([5])[0]), _p32(out))
    assert O.pt_to_affine_ints(out) is None


def test_host_transcript_serialisation_vs_oracle(libs, oracle):
    """Host layer's compressed point encoding == oracle's (both restate arkworks' 33-byte form)."""
    _, host = libs
    O = oracle
    GS = O.derive_points(2, 12)
    pts = [O.affine_to_jac(g)[0] for g in GS] + [O.pt_from_affine_ints(None)]
    pts.append(O.pt_mul(pts[0], O.random_scalars(1, 1)[0]))
    for p in pts:
        out = (C.c_uint8 * 33)()
        host.halo_point_serialize_compressed(p.ctypes.data_as(C.POINTER(C.c_uint64)), out)
        assert bytes(out) == O.pt_serialize_compressed(p)
    # HPoly::eval on the host layer
    xs, z = O.random_scalars(8, 5), O.random_scalars(1, 6)[0]
    out = np.zeros(4, dtype=np.uint64)
    assert host.halo_h_eval(O._p(xs), 7, O._p(z), O._p(out)) == 0
    assert out.tolist() == O.h_eval(xs, z).tolist()


def test_glv_joint_sparse_form(libs, oracle):
    """K4 host side (k_fold_multi's digits): xi = k1 + k2 * lambda (mod r) with |k1|, |k2| < 2^130, written in joint sparse form -- digits in {0, +-1}, the value is preserved,
    and on average at most half of the positions carry an addition."""
    cuda, _ = libs
    LAMBDA = 0x397E65A7D7C1AD71AEE24B27E308F0A61259527EC1D4752E619D1840AF55F1B1
    BETA = 0x2D33357CB532458ED3552A23A8554E5005270D29D19FC7D27B7FD22F0201B547
    if CURVE == "vesta":  # the fields swap roles, and so do the two cube roots (tools/gen_glv_consts.py)
        LAMBDA, BETA = BETA, LAMBDA
    assert pow(LAMBDA, 3, PR.R) == 1 and pow(BETA, 3, PR.P) == 1
    assert PR.pt_mul(PR.GEN, LAMBDA) == (BETA * PR.GEN[0] % PR.P, PR.GEN[1])  # phi(P) = (beta x, y) = lambda * P
    rnd = random.Random(10)
    vals = [0, 1, 2, 3, 5, PR.R - 1, PR.R - 2, LAMBDA, PR.R // 2, 1 << 254, (1 << 128) - 1, 1 << 128] + [rnd.randrange(PR.R) for _ in range(400)]
    busy = total = 0
    for k in vals:
        xi = oracle.to_mont([k])[0]
        d1 = (C.c_int8 * 136)()
        d2 = (C.c_int8 * 136)()
        top = C.c_int()
        assert cuda.halo_test_glv_decompose_jsf(oracle._p(xi), d1, d2, C.byref(top)) == 0
        k1 = sum(int(d) << i for i, d in enumerate(d1))
        k2 = sum(int(d) << i for i, d in enumerate(d2))
        assert (k1 + k2 * LAMBDA - k) % PR.R == 0
        assert abs(k1) < 1 << 131 and abs(k2) < 1 << 131
        assert all(-1 <= x <= 1 for x in d1) and all(-1 <= x <= 1 for x in d2) and top.value <= 132
        assert all(d1[i] == 0 and d2[i] == 0 for i in range(top.value + 1, 136))
        # JSF property: of any three consecutive positions at least one is zero in both rows
        assert all(any(d1[i + j] == 0 and d2[i + j] == 0 for j in range(3)) for i in range(133))
        if k > (1 << 200):
            busy += sum(1 for i in range(top.value + 1) if d1[i] or d2[i])
            total += top.value + 1
    assert busy / total < 0.55
