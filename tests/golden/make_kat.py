"""Generates tests/golden/kat_pcdl_2_10.json: a known-answer file for BASELINE config 1 (PCDL commit + open + check and
one ASDL accumulation step + decider at n = 2^10) that a third party with cargo can check against the real reference.

SURVEY.md section 8(c): the reference stores no expected bytes for any commitment, L / R, U, challenge or accumulator, so
"bit-exact" is GPU == oracle with the oracle anchored to the reference's golden generators.  This file closes the loop
from the other side: every input is derived by a rule that is three lines of Rust (below), every output is written in
the reference's own canonical encodings (points: arkworks `serialize_compressed`, 33 bytes; scalars: 32-byte
little-endian canonical integers), so `cargo test` in the reference crate can recompute and compare (INTEGRATION.md
holds the test).  The non-hiding opening draws nothing from `rng` and is checkable against the UNMODIFIED reference;
the hiding opening and the accumulation step need the reference's `rng` draws replaced by the listed values
(draw order: pcdl.rs:141 p_bar coefficients, pcdl.rs:146 omega_bar; acc.rs:192 h_0, acc.rs:198 omega).

Input rule:   x(tag, i) = Fr::from_le_bytes_mod_order(Sha3_256(tag || (i as u64).to_le_bytes()))

Run:  python tests/golden/make_kat.py      (CPU only; uses the oracle, which tests/test_oracle_golden.py pins to consts.rs)
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

R_MOD = O.R_MOD  # scalar field of the curve selected by HALO_B200_CURVE
# HALO_B200_CURVE=vesta writes the same file for the Vesta instantiation (ark_vesta types in place of ark_pallas,
# same derivation rule for S, H, G_i; nothing in the reference pins it: SURVEY 8(f).4)
KAT_FILE = "kat_pcdl_2_10.json" if O.CURVE == "pallas" else "kat_pcdl_2_10_vesta.json"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), KAT_FILE)


def x(tag, i):
    return int.from_bytes(hashlib.sha3_256(tag.encode() + int(i).to_bytes(8, "little")).digest(), "little") % R_MOD


def xs(tag, k):
    return O.to_mont([x(tag, i) for i in range(k)])


def fr_hex(limbs):
    return O.from_mont(np.asarray(limbs, dtype=np.uint64).reshape(1, 4))[0].to_bytes(32, "little").hex()


def pt_hex(p):
    return O.pt_serialize_compressed(np.asarray(p, dtype=np.uint64)).hex()


def proof_dict(pi):
    d = {"Ls": [pt_hex(np.array(pi.Ls[i])) for i in range(pi.lg_n)], "Rs": [pt_hex(np.array(pi.Rs[i])) for i in range(pi.lg_n)],
         "U": pt_hex(np.array(pi.U)), "c": fr_hex(np.array(pi.c))}
    if pi.hiding:
        d["C_bar"] = pt_hex(np.array(pi.C_bar))
        d["w_prime"] = fr_hex(np.array(pi.w_prime))
    return d


def build():
    n, d, deg = 1 << 10, (1 << 10) - 1, (1 << 10) - 100
    O.derive_params(n)  # S, H, G_0..G_{n-1} by the rule of main.rs:18-45 (== consts.rs, pinned by test_oracle_golden.py)
    S, H, gs = O.params()
    p = xs("halo-b200-kat/p", deg)
    z = xs("halo-b200-kat/z", 1)[0]
    w = xs("halo-b200-kat/w", 1)[0]
    pbar = xs("halo-b200-kat/pbar", deg - 1)
    wbar = xs("halo-b200-kat/wbar", 1)[0]
    v = O.scalar_dot(p, O.construct_powers(z, deg))
    kat = {
        "about": "known answers for rasmus-kirk/halo-accumulation, n = 2^10 (config 1); see tests/golden/make_kat.py"
                 + ("" if O.CURVE == "pallas" else " -- VESTA instantiation (ark_vesta in place of ark_pallas)"),
        "n": n, "d": d, "poly_len": deg,
        "input_rule": "x(tag, i) = Fr::from_le_bytes_mod_order(Sha3_256(tag || u64_le(i)))",
        "inputs": {"p": "x('halo-b200-kat/p', i), i < poly_len", "z": "x('halo-b200-kat/z', 0)", "w": "x('halo-b200-kat/w', 0)",
                   "p_bar (rng draw pcdl.rs:141, degree poly_len - 2)": "x('halo-b200-kat/pbar', i), i < poly_len - 1",
                   "omega_bar (rng draw pcdl.rs:146)": "x('halo-b200-kat/wbar', 0)",
                   "h_0 (rng draw acc.rs:192, 2 coefficients)": "x('halo-b200-kat/h0', i), i < 2",
                   "omega (rng draw acc.rs:198)": "x('halo-b200-kat/acc-w', 0)",
                   "acc p_bar / omega_bar (draws inside the prover's pcdl::open)": "x('halo-b200-kat/acc-pbar', i), i < n - 1; x('halo-b200-kat/acc-wbar', 0)"},
        "samples": {"p[0]": fr_hex(p[0]), "p[1]": fr_hex(p[1]), "z": fr_hex(z), "w": fr_hex(w)},
        "encodings": {"point": "ark_serialize::CanonicalSerialize::serialize_compressed (33 bytes, hex)", "scalar": "32-byte little-endian canonical integer (hex)"},
        "params": {"S": pt_hex(S), "H": pt_hex(H), "G_0": pt_hex(O.affine_to_jac(gs[:1])[0]), "G_1023": pt_hex(O.affine_to_jac(gs[1023:1024])[0])},
        "v = p(z)": fr_hex(v),
    }
    # --- plain MSM (config 2 shape at n = 2^10): sum_i p_i G_i over the first poly_len generators
    kat["msm <p, GS[0..poly_len)>"] = pt_hex(O.msm_affine(gs[:deg], p))
    # --- non-hiding: checkable against the unmodified reference
    C0 = O.pcdl_commit(p, d, None)
    pi0 = O.pcdl_open(p, C0, d, z)
    assert O.pcdl_check(C0, d, z, v, pi0) == 0
    kat["non_hiding"] = {"C = commit(p, d, None)": pt_hex(C0), "proof = open(rng, p, C, d, z, None)": proof_dict(pi0), "check": "Ok"}
    # --- hiding
    C1 = O.pcdl_commit(p, d, w)
    pi1 = O.pcdl_open(p, C1, d, z, w, pbar, wbar)
    assert O.pcdl_check(C1, d, z, v, pi1) == 0
    rc, xis, U = O.pcdl_succinct_check(C1, d, z, v, pi1)
    assert rc == 0
    kat["hiding"] = {"C = commit(p, d, Some(w))": pt_hex(C1), "proof": proof_dict(pi1), "check": "Ok",
                     "succinct_check.h.xis": [fr_hex(xi) for xi in xis], "succinct_check.U": pt_hex(U)}
    # --- one accumulation step over the hiding instance, then verifier and decider
    inst = O.make_instance(C1, d, z, v, pi1)
    h0 = xs("halo-b200-kat/h0", 2)
    aw = xs("halo-b200-kat/acc-w", 1)[0]
    apbar = xs("halo-b200-kat/acc-pbar", n - 1)
    awbar = xs("halo-b200-kat/acc-wbar", 1)[0]
    acc = O.acc_prover(d, [inst], h0, aw, apbar, awbar)
    assert O.acc_verifier(d, [inst], acc) == 0 and O.acc_decider(acc) == 0
    kat["accumulator = acc::prover(rng, d, [instance])"] = {
        "C_bar": pt_hex(np.array(acc.C_bar)), "d": int(acc.d), "z": fr_hex(np.array(acc.z)), "v": fr_hex(np.array(acc.v)),
        "pi": proof_dict(acc.pi), "pi_V.h0": [fr_hex(np.array(acc.h0[i])) for i in range(2)], "pi_V.U0": pt_hex(np.array(acc.U0)),
        "pi_V.w": fr_hex(np.array(acc.w)), "verifier": "Ok", "decider": "Ok"}
    return kat


if __name__ == "__main__":
    kat = build()
    with open(OUT, "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
