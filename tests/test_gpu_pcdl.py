"""GPU parity tests for the PCDL / ASDL path through the host layer (mirror of pcdl.rs / acc.rs) and the C ABI:
identical commitments, L/R values, accumulators and accept/reject decisions as the CPU oracle on the same inputs.
Shapes follow the reference's own tests (pcdl.rs:440-483 test_check*, acc.rs:298-315 test_acc_scheme,
pcdl.rs:381-438 test_u_check, pcdl.rs:485-509 test_construct_h_with_degree_7, pedersen.rs:54-63)."""
import ctypes as C

import numpy as np
import pytest

from oracle.oracle import CURVE, P_MOD, R_MOD  # moduli of the curve selected by HALO_B200_CURVE (test infrastructure)

pytestmark = pytest.mark.gpu
pallas_only = pytest.mark.skipif(CURVE != "pallas", reason="pinned to the reference's Pallas constants")


@pytest.fixture(scope="module")
def env(ctx, oracle):
    """Context with 2^16 generators and the oracle holding the very same parameters."""
    from halo_accumulation_b200 import acc, group, pcdl, pedersen

    ctx.derive_generators(1 << 16)
    S, H = ctx.get_SH()
    oracle.set_params(S, H, ctx.get_generators(0, 1 << 16))
    return dict(ctx=ctx, O=oracle, pcdl=pcdl, acc=acc, pedersen=pedersen, group=group)


def _same_proof(O, a, b):
    """a: halo EvalProof, b: oracle EvalProof -- equal group elements and identical scalars."""
    assert a.lg_n == b.lg_n and a.hiding == b.hiding
    for i in range(a.lg_n):
        assert O.pt_eq(np.array(a.Ls[i]), np.array(b.Ls[i])), f"L[{i}]"
        assert O.pt_eq(np.array(a.Rs[i]), np.array(b.Rs[i])), f"R[{i}]"
    assert O.pt_eq(np.array(a.U), np.array(b.U))
    assert list(a.c) == list(b.c)
    if a.hiding:
        assert O.pt_eq(np.array(a.C_bar), np.array(b.C_bar))
        assert list(a.w_prime) == list(b.w_prime)


def _as_oracle_proof(O, pi):
    return O.EvalProof.from_buffer_copy(bytes(pi))


def test_struct_layouts_match(env):
    from halo_accumulation_b200 import _host

    O = env["O"]
    assert C.sizeof(_host.EvalProof) == C.sizeof(O.EvalProof)
    assert C.sizeof(_host.Instance) == C.sizeof(O.Instance)
    assert C.sizeof(_host.Accumulator) == C.sizeof(O.Accumulator)


def test_vector_helpers(env):
    ctx, O, group = env["ctx"], env["O"], env["group"]
    for n in (1, 2, 31, 1000, 70000):
        a, b = O.random_scalars(n, n), O.random_scalars(n, n + 1)
        assert group.scalar_dot(ctx, a, b).tolist() == O.scalar_dot(a, b).tolist()
        z = O.random_scalars(1, 3 * n)[0]
        assert np.array_equal(group.construct_powers(ctx, z, n), O.construct_powers(z, n))


@pytest.mark.parametrize("lg_n", [1, 2, 3, 4, 7, 10, 16])
def test_h_poly(env, lg_n):
    """HPoly::get_poly / eval (pcdl.rs:56-91), incl. the index convention pinned by pcdl.rs:485-509."""
    ctx, O, pcdl = env["ctx"], env["O"], env["pcdl"]
    xis = O.random_scalars(lg_n + 1, 40 + lg_n)
    h = pcdl.HPoly(xis)
    got = h.get_poly(ctx)
    assert np.array_equal(got, O.h_get_poly(xis))
    z = O.random_scalars(1, 9)[0]
    assert h.eval(z).tolist() == O.h_eval(xis, z).tolist()
    if lg_n == 3:  # test_construct_h_with_degree_7
        x = O.from_mont(xis)
        exp = [1, x[3], x[2], x[2] * x[3], x[1], x[1] * x[3], x[1] * x[2], x[1] * x[2] * x[3]]
        assert O.from_mont(got) == [e % R_MOD for e in exp]


def test_u_check_kat(env):
    """pcdl.rs:381-438 with the reference's fixed inputs GS[0..8], xi = (0,1,2,3); expected value from SURVEY A.6."""
    ctx, O, pcdl, pedersen = env["ctx"], env["O"], env["pcdl"], env["pedersen"]
    xis = O.to_mont([0, 1, 2, 3])
    h = pcdl.HPoly(xis).get_poly(ctx)
    assert O.from_mont(h) == [1, 3, 2, 6, 1, 3, 2, 6]
    U = pedersen.commit(ctx, None, ctx.get_generators(0, 8), h)
    if CURVE == "pallas":
        assert O.pt_to_affine_ints(U) == (0x18CEF7A91C998EAB6266EAA5C7523A520B6F9B56AEFE02B7CB48B226B9C0530C,
                                          0x2CC9CEE89D461087F1312759EFB678EC548F5CDA04F99A56429AE889CC2D7DA3)
    assert O.pt_eq(ctx._h and U, O.pedersen_commit(None, ctx.get_generators(0, 8), h))


def test_pedersen_homomorphism(env):
    """pedersen.rs:54-63"""
    ctx, O, pedersen = env["ctx"], env["O"], env["pedersen"]
    l = 64
    Gs = ctx.get_generators(0, l)
    for rep in range(3):
        m1, m2 = O.random_scalars(l, 10 * rep), O.random_scalars(l, 10 * rep + 1)
        w1, w2 = O.random_scalars(2, 10 * rep + 2)
        msum = ctx.test_fp_op(1, 1, m1, m2)
        wsum = ctx.test_fp_op(1, 1, w1.reshape(1, 4), w2.reshape(1, 4))[0]
        inner = pedersen.commit(ctx, wsum, Gs, msum)
        outer = O.pt_add(pedersen.commit(ctx, w1, Gs, m1), pedersen.commit(ctx, w2, Gs, m2))
        assert O.pt_eq(inner, outer)
        assert O.pt_eq(inner, O.pedersen_commit(wsum, Gs, msum))
    with pytest.raises(Exception):
        pedersen.commit(ctx, None, Gs, m1[:10])  # length mismatch asserts (pedersen.rs:7-12)


def test_serialize_compressed_matches_oracle(env):
    from halo_accumulation_b200 import _host

    ctx, O = env["ctx"], env["O"]
    pts = [O.affine_to_jac(g)[0] for g in ctx.get_generators(0, 16)]
    pts += [O.pt_mul(pts[0], O.to_mont([R_MOD - 1])[0]), O.pt_from_affine_ints(None), O.pt_add(pts[1], pts[2])]
    for p in pts:
        out = (C.c_uint8 * 33)()
        _host.lib().halo_point_serialize_compressed(p.ctypes.data_as(C.POINTER(C.c_uint64)), out)
        assert bytes(out) == O.pt_serialize_compressed(p)


@pytest.mark.parametrize("n,hiding", [(2, 0), (2, 1), (4, 1), (8, 0), (64, 1), (512, 0), (512, 1), (4096, 1), (1 << 14, 0)])
def test_pcdl_commit_open_check(env, n, hiding):
    """test_check / test_check_no_hiding (pcdl.rs:440-483) with ragged degree d' < d, against the oracle."""
    ctx, O, pcdl = env["ctx"], env["O"], env["pcdl"]
    d = n - 1
    dp = max(1, (2 * n) // 3 - 1) if n > 2 else 1
    p = O.random_scalars(dp + 1, 1000 + n)
    w = O.random_scalars(1, 7)[0] if hiding else None
    q = O.random_scalars(dp, 8) if hiding else None
    wb = O.random_scalars(1, 9)[0] if hiding else None
    Cm = pcdl.commit(ctx, p, d, w)
    assert O.pt_eq(Cm, O.pcdl_commit(p, d, w, threads=8))
    z = O.random_scalars(1, 10)[0]
    v = O.scalar_dot(p, O.construct_powers(z, dp + 1))
    pi = pcdl.open(ctx, p, Cm, d, z, w, q, wb)
    _same_proof(O, pi, O.pcdl_open(p, Cm, d, z, w, q, wb, threads=8))
    pcdl.check(ctx, Cm, d, z, v, pi)                                   # accepts
    assert O.pcdl_check(Cm, d, z, v, _as_oracle_proof(O, pi), threads=8) == 0  # the oracle accepts the GPU proof
    h, U = pcdl.succinct_check(ctx, Cm, d, z, v, pi)
    rc, xis, Uo = O.pcdl_succinct_check(Cm, d, z, v, _as_oracle_proof(O, pi))
    assert rc == 0 and np.array_equal(h.xis, xis) and O.pt_eq(U, Uo)
    # reject paths: same decision and same failing check as the oracle
    bad = type(pi).from_buffer_copy(bytes(pi))
    bad.c[0] ^= 1
    with pytest.raises(pcdl.Rejected) as e:
        pcdl.check(ctx, Cm, d, z, v, bad)
    assert e.value.code == O.pcdl_check(Cm, d, z, v, _as_oracle_proof(O, bad)) == -10
    if n > 2:
        bad = type(pi).from_buffer_copy(bytes(pi))
        C.memmove(bad.Ls[1], bytes(pi.Rs[0]), 96)
        with pytest.raises(pcdl.Rejected) as e:
            pcdl.check(ctx, Cm, d, z, v, bad)
        assert e.value.code == O.pcdl_check(Cm, d, z, v, _as_oracle_proof(O, bad)) == -10
    v_bad = O.random_scalars(1, 77)[0]
    with pytest.raises(pcdl.Rejected):
        pcdl.check(ctx, Cm, d, v_bad, v, pi)


def test_pcdl_full_check_rejects_wrong_U(env):
    """pcdl.rs:339: a proof that passes the succinct check but whose U is not <G, h> must fail step 5.  Built by
    running the honest prover against a context whose generator G_5 was replaced (so its U is consistent with its
    own L/R but not with the verifier's key)."""
    ctx, O, pcdl = env["ctx"], env["O"], env["pcdl"]
    import halo_accumulation_b200 as H

    n, d = 64, 63
    S, Hh = ctx.get_SH()
    gs = ctx.get_generators(0, n).copy()
    gs[5] = ctx.get_generators(1000, 1)[0]
    other = H.Context(0, 1 << 10)
    try:
        other.load_generators(S, Hh, gs)
        p, z = O.random_scalars(n, 1), O.random_scalars(1, 2)[0]
        p[5] = 0  # coefficient 5 is zero, so C is the same under both keys ...
        Cm = pcdl.commit(other, p, d)
        assert O.pt_eq(Cm, pcdl.commit(ctx, p, d))
        v = O.scalar_dot(p, O.construct_powers(z, n))
        pi = pcdl.open(other, p, Cm, d, z)
        pcdl.check(other, Cm, d, z, v, pi)
    finally:
        other.close()
    # ... but L/R/U were folded with the foreign G_5: whichever check trips, GPU and oracle must agree
    rc = O.pcdl_check(Cm, d, z, v, _as_oracle_proof(O, pi))
    assert rc in (-10, -11)
    with pytest.raises(pcdl.Rejected) as e:
        pcdl.check(ctx, Cm, d, z, v, pi)
    assert e.value.code == rc


def test_malformed_inputs_are_refused_at_the_boundary(env):
    """Raw limbs crossing the C ABI into the verifier-side calls are validated once (canonical residues, points on the curve):
    arkworks types cannot hold anything else, so such input is HALO_EINVAL, not an accept / reject decision."""
    from halo_accumulation_b200 import HaloError

    ctx, O, pcdl, acc = env["ctx"], env["O"], env["pcdl"], env["acc"]
    n, d = 64, 63
    q, _ = _random_instance(env, d, 8800)
    h0, (w, wb), qq = O.random_scalars(2, 8801), O.random_scalars(2, 8802), O.random_scalars(n - 1, 8803)
    a = acc.prover(ctx, d, [q], h0, w, qq, wb)
    acc.verifier(ctx, d, [q], a)
    pi = q.pi
    Cm, z, v = np.array(q.C), np.array(q.z), np.array(q.v)
    pcdl.check(ctx, Cm, d, z, v, pi)

    def refused(f):
        with pytest.raises(HaloError) as e:
            f()
        assert e.value.code == -1, e.value.code

    bad = type(pi).from_buffer_copy(bytes(pi))
    bad.Ls[2][5] ^= 4                                    # y of L_2: off the curve
    refused(lambda: pcdl.check(ctx, Cm, d, z, v, bad))
    refused(lambda: pcdl.succinct_check(ctx, Cm, d, z, v, bad))
    bad = type(pi).from_buffer_copy(bytes(pi))
    bad.U[3] = 0xFFFFFFFFFFFFFFFF                        # x >= p: not a canonical residue
    refused(lambda: pcdl.check(ctx, Cm, d, z, v, bad))
    bad = type(pi).from_buffer_copy(bytes(pi))
    bad.c[3] = 0x7FFFFFFFFFFFFFFF                        # scalar >= r
    refused(lambda: pcdl.check(ctx, Cm, d, z, v, bad))
    Cbad = Cm.copy()
    Cbad[0] ^= np.uint64(1)
    refused(lambda: pcdl.check(ctx, Cbad, d, z, v, pi))
    zbad = z.copy()
    zbad[3] = np.uint64(0x4000000000000000) + np.uint64(1 << 40)  # >= r
    refused(lambda: pcdl.check(ctx, Cm, d, zbad, v, pi))
    abad = type(a).from_buffer_copy(bytes(a))
    abad.U0[1] ^= 2
    refused(lambda: acc.verifier(ctx, d, [q], abad))
    refused(lambda: acc.decider(ctx, abad))
    abad = type(a).from_buffer_copy(bytes(a))
    abad.C_bar[8] ^= 1                                   # z coordinate of a Jacobian point: no longer on the curve
    refused(lambda: acc.decider(ctx, abad))
    qbad = type(q).from_buffer_copy(bytes(q))
    qbad.pi.Rs[0][0] ^= 1
    refused(lambda: acc.verifier(ctx, d, [qbad], a))
    refused(lambda: acc.prover(ctx, d, [qbad], h0, w, qq, wb))
    # infinity is a valid point (arkworks: z = 0): a decision, not an error
    inf = type(pi).from_buffer_copy(bytes(pi))
    C.memset(inf.Ls[1], 0, 96)
    with pytest.raises(pcdl.Rejected):
        pcdl.check(ctx, Cm, d, z, v, inf)
    acc.verifier(ctx, d, [q], a)                         # the context is still usable


def test_pcdl_argument_errors(env):
    import halo_accumulation_b200 as H

    ctx, O, pcdl = env["ctx"], env["O"], env["pcdl"]
    p = O.random_scalars(8, 1)
    with pytest.raises(H.HaloError):
        pcdl.commit(ctx, p, 6)            # d + 1 not a power of two (pcdl.rs:102)
    with pytest.raises(H.HaloError):
        pcdl.commit(ctx, p, 3)            # degree exceeds d (pcdl.rs:103)
    with pytest.raises(H.HaloError):
        pcdl.commit(ctx, p, (1 << 17) - 1)  # d > D (pcdl.rs:104)
    # halo_h_msm_with (pcdl::check in one call): equals halo_h_msm + halo_msm on the same inputs; argument errors
    from halo_accumulation_b200._capi import p64

    lib = ctx._lib
    lg = 10
    xis = O.random_scalars(lg + 1, 2)
    k = 22
    bases = ctx.get_generators(7, k).copy()
    inf = np.zeros(k, dtype=np.uint8)
    inf[3] = 1
    sc = O.random_scalars(k, 3)
    out_h, out_s, ref_h = (np.zeros(12, dtype=np.uint64) for _ in range(3))
    assert lib.halo_h_msm_with(ctx._h, p64(xis), lg, p64(bases), inf.ctypes.data_as(C.POINTER(C.c_uint8)), p64(sc), C.c_uint64(k),
                               p64(out_h), p64(out_s)) == 0
    assert lib.halo_h_msm(ctx._h, p64(xis), lg, p64(ref_h)) == 0
    assert O.pt_eq(out_h, ref_h) and O.pt_eq(out_s, ctx.msm(bases, sc, inf))
    assert O.pt_eq(out_h, O.msm_affine(ctx.get_generators(0, 1 << lg), O.h_get_poly(xis), threads=8))
    assert lib.halo_h_msm_with(ctx._h, p64(xis), lg, p64(bases), None, p64(sc), C.c_uint64(0), p64(out_h), p64(out_s)) == 0
    assert O.pt_eq(out_h, ref_h) and O.pt_to_affine(out_s)[1]          # empty companion: infinity
    assert lib.halo_h_msm_with(ctx._h, p64(xis), lg, None, None, p64(sc), C.c_uint64(k), p64(out_h), p64(out_s)) < 0
    assert lib.halo_h_msm_with(ctx._h, p64(xis), 40, p64(bases), None, p64(sc), C.c_uint64(k), p64(out_h), p64(out_s)) < 0
    assert lib.halo_h_msm_with(ctx._h, p64(xis), lg, p64(bases), None, p64(sc), C.c_uint64(5000), p64(out_h), p64(out_s)) < 0


def _random_instance(env, d, seed):
    """benches/acc.rs:15-29 / acc.rs:264-278 with explicit draws."""
    ctx, O, pcdl, acc = env["ctx"], env["O"], env["pcdl"], env["acc"]
    n = d + 1
    dp = max(1, n // 2)
    p = O.random_scalars(dp + 1, seed)
    w, z, wb = O.random_scalars(3, seed + 1)
    q = O.random_scalars(dp, seed + 2)
    Cm = pcdl.commit(ctx, p, d, w)
    v = O.scalar_dot(p, O.construct_powers(z, dp + 1))
    pi = pcdl.open(ctx, p, Cm, d, z, w, q, wb)
    pio = O.pcdl_open(p, O.pcdl_commit(p, d, w), d, z, w, q, wb, threads=8)
    _same_proof(O, pi, pio)
    return acc.new_instance(Cm, d, z, v, pi), O.make_instance(Cm, d, z, v, pio)


@pytest.mark.parametrize("n,steps", [(4, 3), (16, 4), (1024, 3)])
def test_acc_chain_matches_oracle(env, n, steps):
    """test_acc_scheme (acc.rs:298-315): prover + verifier per step, decider at the end; every accumulator field
    equals the oracle's."""
    ctx, O, acc = env["ctx"], env["O"], env["acc"]
    d = n - 1
    a, ao = None, None
    for s in range(steps):
        q, qo = _random_instance(env, d, 500 + 10 * s + n)
        qs = [acc.to_instance(a), q] if a is not None else [q]
        qso = [O.acc_to_instance(ao), qo] if ao is not None else [qo]
        h0, (w, wb), qq = O.random_scalars(2, 900 + s), O.random_scalars(2, 910 + s), O.random_scalars(n - 1, 920 + s)
        a = acc.prover(ctx, d, qs, h0, w, qq, wb)
        ao = O.acc_prover(d, qso, h0, w, qq, wb, threads=8)
        assert O.pt_eq(np.array(a.C_bar), np.array(ao.C_bar))
        assert (a.d, list(a.z), list(a.v), list(a.w)) == (ao.d, list(ao.z), list(ao.v), list(ao.w))
        assert O.pt_eq(np.array(a.U0), np.array(ao.U0)) and bytes(a.h0) == bytes(ao.h0)
        _same_proof(O, a.pi, ao.pi)
        acc.verifier(ctx, d, qs, a)
        assert O.acc_verifier(d, qso, ao) == 0
        # the oracle accepts the GPU accumulator and vice versa
        assert O.acc_verifier(d, qso, O.Accumulator.from_buffer_copy(bytes(a))) == 0
        # reject paths agree
        bad = type(a).from_buffer_copy(bytes(a))
        bad.v[0] ^= 1
        with pytest.raises(acc.Rejected) as e:
            acc.verifier(ctx, d, qs, bad)
        assert e.value.code == O.acc_verifier(d, qso, O.Accumulator.from_buffer_copy(bytes(bad))) == -17
        bad = type(a).from_buffer_copy(bytes(a))
        bad.h0[1][0] ^= 1
        with pytest.raises(acc.Rejected) as e:
            acc.verifier(ctx, d, qs, bad)
        assert e.value.code == O.acc_verifier(d, qso, O.Accumulator.from_buffer_copy(bytes(bad))) == -12
    acc.decider(ctx, a)
    assert O.acc_decider(ao, threads=8) == 0
    bad = type(a).from_buffer_copy(bytes(a))
    bad.z[0] ^= 1
    with pytest.raises(acc.Rejected):
        acc.decider(ctx, bad)


def test_open_check_2_16_full_parity(env):
    ctx, O, pcdl = env["ctx"], env["O"], env["pcdl"]
    n = 1 << 16
    d = n - 1
    p = O.random_scalars(n, 2)
    z = O.random_scalars(1, 3)[0]
    Cm = pcdl.commit(ctx, p, d)
    pi = pcdl.open(ctx, p, Cm, d, z)
    _same_proof(O, pi, O.pcdl_open(p, Cm, d, z, threads=8))
    v = O.scalar_dot(p, O.construct_powers(z, n))
    pcdl.check(ctx, Cm, d, z, v, pi)


def test_open_and_full_check_2_20(ctx, oracle):
    """BASELINE config 3: PCDL open + full check at n = 2^20 on one GPU.  The oracle cannot re-open at this size in
    seconds, so parity is by properties: the GPU proof is accepted by the GPU check AND by the oracle's full check
    (succinct check + h-expansion + its own MSM over the same generators), a flipped proof is rejected by both with the
    same failing condition, and v = p(z) agrees with the oracle's dot product."""
    from halo_accumulation_b200 import group, pcdl

    O = oracle
    n = 1 << 20
    d = n - 1
    ctx.derive_generators(n)
    try:
        ctx.precompute_generators(0)
        S, Hh = ctx.get_SH()
        O.set_params(S, Hh, ctx.get_generators(0, n))
        p = O.random_scalars(n - 12345, 2)  # ragged degree d' < d
        z = O.random_scalars(1, 3)[0]
        Cm = pcdl.commit(ctx, p, d)
        pi = pcdl.open(ctx, p, Cm, d, z)
        v = group.scalar_dot(ctx, p, group.construct_powers(ctx, z, p.shape[0]))
        assert v.tolist() == O.scalar_dot(p, O.construct_powers(z, p.shape[0])).tolist()
        pcdl.check(ctx, Cm, d, z, v, pi)
        assert O.pcdl_check(Cm, d, z, v, O.EvalProof.from_buffer_copy(bytes(pi)), threads=16) == 0
        bad = type(pi).from_buffer_copy(bytes(pi))
        C.memmove(bad.Rs[7], bytes(pi.Ls[7]), 96)  # a valid point in the wrong place: a verifier decision, same on both sides
        with pytest.raises(pcdl.Rejected) as e:
            pcdl.check(ctx, Cm, d, z, v, bad)
        assert e.value.code == O.pcdl_check(Cm, d, z, v, O.EvalProof.from_buffer_copy(bytes(bad)), threads=16) == -10
        bad.Rs[7][0] ^= 1  # off the curve: malformed input (HALO_EINVAL), not a decision
        from halo_accumulation_b200 import HaloError
        with pytest.raises(HaloError) as e:
            pcdl.check(ctx, Cm, d, z, v, bad)
        assert e.value.code == -1
        # hiding variant (deferred head rounds over the FIXED-base tables, blinding polynomial, w'): accepted by both
        w, wb = O.random_scalars(2, 5)
        q = O.random_scalars(p.shape[0] - 1, 6)
        Cw = pcdl.commit(ctx, p, d, w)
        piw = pcdl.open(ctx, p, Cw, d, z, w, q, wb)
        pcdl.check(ctx, Cw, d, z, v, piw)
        assert O.pcdl_check(Cw, d, z, v, O.EvalProof.from_buffer_copy(bytes(piw)), threads=16) == 0
        # without the FIXED-base tables: no deferral, variable-base round MSMs (pair-tree passes with the H' tail at this
        # size); same proof, since every L, R, U is the same group element
        ctx.set_fixed_base(False)
        try:
            piv = pcdl.open(ctx, p, Cm, d, z)
            for i in range(pi.lg_n):
                assert O.pt_eq(np.array(piv.Ls[i]), np.array(pi.Ls[i])) and O.pt_eq(np.array(piv.Rs[i]), np.array(pi.Rs[i])), i
            assert O.pt_eq(np.array(piv.U), np.array(pi.U)) and list(piv.c) == list(pi.c)
            pcdl.check(ctx, Cm, d, z, v, piv)
            # and the HIDING opening: the deferred / FIXED-base proof above must equal, field by field, the proof of the
            # round-by-round variable-base path on the same draws (two independent kernel paths; the oracle accepted it)
            piwv = pcdl.open(ctx, p, Cw, d, z, w, q, wb)
            for i in range(piw.lg_n):
                assert O.pt_eq(np.array(piwv.Ls[i]), np.array(piw.Ls[i])) and O.pt_eq(np.array(piwv.Rs[i]), np.array(piw.Rs[i])), i
            assert O.pt_eq(np.array(piwv.U), np.array(piw.U)) and list(piwv.c) == list(piw.c)
            assert O.pt_eq(np.array(piwv.C_bar), np.array(piw.C_bar)) and list(piwv.w_prime) == list(piw.w_prime)
        finally:
            ctx.set_fixed_base(True)
    finally:
        ctx.derive_generators(1 << 16)


@pytest.mark.parametrize("n,defer", [(2, 1), (4, 2), (16, 4), (64, 3), (1024, 1), (1024, 2), (1024, 3), (1 << 14, 3), (1 << 14, 4)])
def test_open_deferred_head_rounds(env, n, defer):
    """pcdl.rs:216-218 unrolled: the first `defer` rounds leave the generators untouched (L / R as MSMs over GS with
    per-index coefficients) and k_fold_multi materialises G^(defer) in one joint pass.  Same L, R, U, c as the oracle's
    round-by-round fold, hiding and non-hiding."""
    ctx, O, pcdl = env["ctx"], env["O"], env["pcdl"]
    d = n - 1
    p = O.random_scalars(max(1, n - n // 5), 4000 + n + defer)
    z = O.random_scalars(1, 11)[0]
    w, wb = O.random_scalars(2, 12)
    q = O.random_scalars(max(1, p.shape[0] - 1), 13)
    ctx.set_tuning("ipa_defer_rounds", defer)
    try:
        Cm = pcdl.commit(ctx, p, d)
        _same_proof(O, pcdl.open(ctx, p, Cm, d, z), O.pcdl_open(p, Cm, d, z, threads=8))
        if n > 2:
            Cw = pcdl.commit(ctx, p, d, w)
            _same_proof(O, pcdl.open(ctx, p, Cw, d, z, w, q, wb), O.pcdl_open(p, Cw, d, z, w, q, wb, threads=8))
    finally:
        ctx.set_tuning("ipa_defer_rounds", -1)


@pytest.mark.parametrize("lg,defer2,freeze", [(16, 4, 0), (15, 2, 0), (14, 3, 2048), (12, 4, 256)])
def test_open_second_deferred_stage(env, lg, defer2, freeze):
    """Later deferred stages over the materialised vector ("ipa_defer2_rounds"; off by default because it measured slower at
    2^20, profiles/r02_ipa_stage2_probe.txt): the next rounds run as per-index-coefficient MSMs over the current generator
    vector and ONE k_fold_multi takes it down 2^D-fold.  Same L, R, U, c as the oracle's round-by-round fold."""
    ctx, O, pcdl = env["ctx"], env["O"], env["pcdl"]
    n = 1 << lg
    d = n - 1
    p = O.random_scalars(n - 11, 6000 + lg)
    z = O.random_scalars(1, 61)[0]
    w, wb = O.random_scalars(2, 62)
    q = O.random_scalars(p.shape[0] - 1, 63)
    ctx.set_tuning("ipa_defer2_rounds", defer2)
    if freeze:
        ctx.set_tuning("ipa_freeze_len", freeze)
    try:
        Cm = pcdl.commit(ctx, p, d)
        _same_proof(O, pcdl.open(ctx, p, Cm, d, z), O.pcdl_open(p, Cm, d, z, threads=8))
        Cw = pcdl.commit(ctx, p, d, w)
        _same_proof(O, pcdl.open(ctx, p, Cw, d, z, w, q, wb), O.pcdl_open(p, Cw, d, z, w, q, wb, threads=8))
    finally:
        ctx.set_tuning("ipa_defer2_rounds", 0)
        ctx.set_tuning("ipa_freeze_len", 0)


@pytest.mark.parametrize("lg,defer", [(13, 0), (14, 2), (15, 3)])
def test_open_fold_kernel_with_out_of_line_multiplication(env, lg, defer):
    """k_fold_multi exists twice: inlined multiplication (few outputs) and, in a second translation unit, the multiplication as
    an out-of-line call (from 2^17 outputs: the joint fold of an opening at 2^20).  Forced here for every fold of small openings
    ("ipa_fold_call_min_lg" = 0), single-round folds and deferred head rounds: same L, R, U, c as the oracle, and as the inlined
    kernel ("ipa_fold_call_min_lg" = -1)."""
    ctx, O, pcdl = env["ctx"], env["O"], env["pcdl"]
    n = 1 << lg
    d = n - 1
    p = O.random_scalars(n - 3, 7000 + lg)
    z = O.random_scalars(1, 71)[0]
    exp = None
    try:
        ctx.set_tuning("ipa_defer_rounds", defer)
        Cm = pcdl.commit(ctx, p, d)
        exp = O.pcdl_open(p, Cm, d, z, threads=8)
        for min_lg in (0, -1):
            ctx.set_tuning("ipa_fold_call_min_lg", min_lg)
            _same_proof(O, pcdl.open(ctx, p, Cm, d, z), exp)
    finally:
        ctx.set_tuning("ipa_fold_call_min_lg", 17)
        ctx.set_tuning("ipa_defer_rounds", -1)


def test_open_deferred_head_fixed_base_2_17(ctx, oracle):
    """The automatic policy: with FIXED-base tables covering the opening, three rounds are deferred and their L / R take
    the shared-bucket-set path (dot * H' added on the host).  n = 2^17 is the smallest size it triggers at; the oracle
    re-opens by the reference's round-by-round algorithm."""
    from halo_accumulation_b200 import pcdl

    O = oracle
    n = 1 << 17
    d = n - 1
    ctx.derive_generators(n)
    try:
        ctx.precompute_generators(0)
        S, Hh = ctx.get_SH()
        O.set_params(S, Hh, ctx.get_generators(0, n))
        p = O.random_scalars(n - 777, 21)
        z = O.random_scalars(1, 22)[0]
        Cm = pcdl.commit(ctx, p, d)
        _same_proof(O, pcdl.open(ctx, p, Cm, d, z), O.pcdl_open(p, Cm, d, z, threads=16))
    finally:
        ctx.derive_generators(1 << 16)
        S, Hh = ctx.get_SH()
        O.set_params(S, Hh, ctx.get_generators(0, 1 << 16))


def test_kat_file_reproduced_on_the_gpu(env):
    """The committed known-answer file (tests/golden/kat_pcdl_2_10.json, canonical arkworks encodings): commitments,
    every L / R, U, c, the hiding proof and a full accumulation step computed on the GPU serialise to the same bytes."""
    import importlib.util
    import json
    import os

    ctx, O, pcdl, acc = env["ctx"], env["O"], env["pcdl"], env["acc"]
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_kat", os.path.join(here, "golden", "make_kat.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    kat = json.load(open(os.path.join(here, "golden", mk.KAT_FILE)))
    n, d, deg = kat["n"], kat["d"], kat["poly_len"]
    p, z, w = mk.xs("halo-b200-kat/p", deg), mk.xs("halo-b200-kat/z", 1)[0], mk.xs("halo-b200-kat/w", 1)[0]
    pbar, wbar = mk.xs("halo-b200-kat/pbar", deg - 1), mk.xs("halo-b200-kat/wbar", 1)[0]
    assert mk.fr_hex(p[0]) == kat["samples"]["p[0]"] and mk.fr_hex(z) == kat["samples"]["z"]
    S, H = ctx.get_SH()
    assert mk.pt_hex(S) == kat["params"]["S"] and mk.pt_hex(H) == kat["params"]["H"]
    assert mk.pt_hex(O.affine_to_jac(ctx.get_generators(1023, 1))[0]) == kat["params"]["G_1023"]
    from halo_accumulation_b200 import group

    v = group.scalar_dot(ctx, p, group.construct_powers(ctx, z, deg))
    assert mk.fr_hex(v) == kat["v = p(z)"]
    assert mk.pt_hex(ctx.msm_gens(p)) == kat["msm <p, GS[0..poly_len)>"]
    C0 = pcdl.commit(ctx, p, d)
    assert mk.pt_hex(C0) == kat["non_hiding"]["C = commit(p, d, None)"]
    assert mk.proof_dict(pcdl.open(ctx, p, C0, d, z)) == kat["non_hiding"]["proof = open(rng, p, C, d, z, None)"]
    C1 = pcdl.commit(ctx, p, d, w)
    assert mk.pt_hex(C1) == kat["hiding"]["C = commit(p, d, Some(w))"]
    pi1 = pcdl.open(ctx, p, C1, d, z, w, pbar, wbar)
    assert mk.proof_dict(pi1) == kat["hiding"]["proof"]
    h, U = pcdl.succinct_check(ctx, C1, d, z, v, pi1)
    assert [mk.fr_hex(x) for x in h.xis] == kat["hiding"]["succinct_check.h.xis"] and mk.pt_hex(U) == kat["hiding"]["succinct_check.U"]
    inst = acc.new_instance(C1, d, z, v, pi1)
    a = acc.prover(ctx, d, [inst], mk.xs("halo-b200-kat/h0", 2), mk.xs("halo-b200-kat/acc-w", 1)[0],
                   mk.xs("halo-b200-kat/acc-pbar", n - 1), mk.xs("halo-b200-kat/acc-wbar", 1)[0])
    acc.verifier(ctx, d, [inst], a)
    acc.decider(ctx, a)
    ka = kat["accumulator = acc::prover(rng, d, [instance])"]
    assert mk.pt_hex(np.array(a.C_bar)) == ka["C_bar"] and mk.fr_hex(np.array(a.z)) == ka["z"] and mk.fr_hex(np.array(a.v)) == ka["v"]
    assert mk.proof_dict(a.pi) == ka["pi"]
    assert mk.pt_hex(np.array(a.U0)) == ka["pi_V.U0"] and mk.fr_hex(np.array(a.w)) == ka["pi_V.w"]


@pytest.mark.parametrize("passes", [2, 4])
def test_open_with_pair_tree_in_the_round_msms(env, passes):
    """The L / R MSMs of large openings take pair-tree passes (automatic from 2^23 entries) with the H' term riding as
    a one-element tail: forced here at n = 2^12 and compared with the oracle's opening, with and without deferral."""
    ctx, O, pcdl = env["ctx"], env["O"], env["pcdl"]
    n = 1 << 12
    d = n - 1
    p = O.random_scalars(n - 5, 77)
    z = O.random_scalars(1, 78)[0]
    w, wb = O.random_scalars(2, 79)
    q = O.random_scalars(n - 6, 80)
    ref = O.pcdl_open(p, O.pcdl_commit(p, d, w, threads=8), d, z, w, q, wb, threads=8)
    ctx.set_tuning("pair_passes", passes)
    try:
        for defer in (0, 2):
            ctx.set_tuning("ipa_defer_rounds", defer)
            Cw = pcdl.commit(ctx, p, d, w)
            _same_proof(O, pcdl.open(ctx, p, Cw, d, z, w, q, wb), ref)
    finally:
        ctx.set_tuning("pair_passes", -1)
        ctx.set_tuning("ipa_defer_rounds", -1)


def test_verifier_batched_checks_keep_the_reference_order(env):
    """common_subroutine runs commit(h_0) and every instance's succinct check in one device round trip;
    the `ensure!`s must still fire in the reference's order (acc.rs:152-170): U_0 first, then per instance the succinct
    check and d_i == d.  Every decision is compared with the oracle's on the same (corrupted) inputs."""
    ctx, O, acc = env["ctx"], env["O"], env["acc"]
    n, d = 64, 63
    q1, q1o = _random_instance(env, d, 4100)
    q2, q2o = _random_instance(env, d, 4200)
    h0, (w, wb), qq = O.random_scalars(2, 4300), O.random_scalars(2, 4310), O.random_scalars(n - 1, 4320)
    a = acc.prover(ctx, d, [q1, q2], h0, w, qq, wb)
    ao = O.acc_prover(d, [q1o, q2o], h0, w, qq, wb, threads=8)
    _same_proof(O, a.pi, ao.pi)
    acc.verifier(ctx, d, [q1, q2], a)

    def both(qs, qso, accum):
        code_o = O.acc_verifier(d, qso, O.Accumulator.from_buffer_copy(bytes(accum)))
        with pytest.raises(acc.Rejected) as e:
            acc.verifier(ctx, d, qs, accum)
        assert e.value.code == code_o, (e.value.code, code_o)
        return code_o

    def corrupt(q, qo):  # second instance with a wrong evaluation proof scalar
        b, bo = type(q).from_buffer_copy(bytes(q)), type(qo).from_buffer_copy(bytes(qo))
        b.pi.c[0] ^= 1
        bo.pi.c[0] ^= 1
        return b, bo

    q2b, q2bo = corrupt(q2, q2o)
    assert both([q1, q2b], [q1o, q2bo], a) == -10          # succinct check of instance 2
    q1b, q1bo = corrupt(q1, q1o)
    assert both([q1b, q2b], [q1bo, q2bo], a) == -10        # ... of instance 1 (both wrong)
    badU = type(a).from_buffer_copy(bytes(a))
    badU.h0[0][0] ^= 1
    assert both([q1, q2b], [q1o, q2bo], badU) == -12       # U_0 is checked before any instance
    # d_i != d: an instance of another size is succinct-checked at ITS size first (accepted), then refused
    q3, q3o = _random_instance(env, 31, 4400)
    assert both([q1, q3], [q1o, q3o], a) == -13
    q3b, q3bo = corrupt(q3, q3o)
    assert both([q1, q3b], [q1o, q3bo], a) == -10          # its succinct check comes before the d comparison
    # malformed input (the reference panics): falls back to the one-by-one path and reports an error, not a decision
    q4 = type(q1).from_buffer_copy(bytes(q1))
    q4.pi.lg_n = 5
    from halo_accumulation_b200 import HaloError

    with pytest.raises(HaloError) as e:
        acc.verifier(ctx, d, [q4, q2], a)
    assert e.value.code == -1
    acc.verifier(ctx, d, [q1, q2], a)                      # the context is still usable
