"""GPU parity tests at the sizes BASELINE.json's configs name ("untested configs"):

  config 5 / headline  one MSM of 2^24 points, FIXED- and variable-base, against the oracle's arkworks-shaped Pippenger at the
                       SAME size, against the sum of 16 oracle-checked 2^20 sub-MSMs, and against the discrete-log property
                       of the derived generators (sum a_i G_i = (sum a_i s_i) * (-1, 2), main.rs:18-32);
  config 4             the IVC chain of benches/acc.rs:64-98 (k x {random_instance, prover}, then k verifiers + decider):
                       k = 64 at n = 2^12 with EVERY accumulator field equal to the oracle's chain, and at n = 2^20 with the
                       oracle's verifier on spot steps and the oracle's decider on the final accumulator;
  threads              two contexts opening concurrently from two host threads (per-context fold op list).

Everything goes through the C ABI (ctypes); the oracle is the checker only."""
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _threads(O):
    return max(1, O.lib().orc_num_threads())


def test_msm_2_24_fixed_and_variable_match_oracle(halo, oracle):
    O = oracle
    n = 1 << 24
    T = _threads(O)
    c = halo.Context(0, n)
    try:
        c.derive_generators(n)
        gs = c.get_generators(0, n)
        sc = O.random_scalars(n, 4)  # seed 4
        exp = O.msm_affine(gs, sc, threads=T)                       # the reference's algorithm at the same size
        assert O.pt_eq(O.msm_derived_by_dlog(0, sc, threads=T), exp)  # the property agrees with the Pippenger
        var = c.msm_gens(sc)                                        # no tables yet: variable base (W bucket sets)
        assert O.pt_eq(var, exp), "variable-base 2^24 MSM != oracle"
        c.precompute_generators(0)
        fix = c.msm_gens(sc)                                        # FIXED-base tables, pair tree P = 4, blocking split call
        assert O.pt_eq(fix, exp), "FIXED-base 2^24 MSM != oracle"
        t = c.msm_gens_submit(sc)
        assert O.pt_eq(c.msm_gens_collect(t), exp)                  # the pipelined entry points
        parts = []
        for k in range(16):                                          # 16 oracle-checked sub-MSMs of 2^20 points
            lo, hi = k << 20, (k + 1) << 20
            pk = c.msm_gens(sc[lo:hi], off=lo)
            assert O.pt_eq(pk, O.msm_affine(gs[lo:hi], sc[lo:hi], threads=T)), f"sub-MSM {k}"
            parts.append(pk)
        assert O.pt_eq(halo.points_sum(np.array(parts)), exp)
        bad, live = halo.check_canaries()
        assert bad == 0
    finally:
        c.close()


def _rs(rng, m):
    a = rng.integers(0, 1 << 64, size=(m, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    return a


def _same_proof(O, a, b):
    assert a.lg_n == b.lg_n and a.hiding == b.hiding
    for i in range(a.lg_n):
        assert O.pt_eq(np.array(a.Ls[i]), np.array(b.Ls[i])), f"L[{i}]"
        assert O.pt_eq(np.array(a.Rs[i]), np.array(b.Rs[i])), f"R[{i}]"
    assert O.pt_eq(np.array(a.U), np.array(b.U)) and list(a.c) == list(b.c)
    if a.hiding:
        assert O.pt_eq(np.array(a.C_bar), np.array(b.C_bar)) and list(a.w_prime) == list(b.w_prime)


def _same_acc(O, a, ao):
    assert O.pt_eq(np.array(a.C_bar), np.array(ao.C_bar))
    assert (a.d, list(a.z), list(a.v), list(a.w)) == (ao.d, list(ao.z), list(ao.v), list(ao.w))
    assert O.pt_eq(np.array(a.U0), np.array(ao.U0)) and bytes(a.h0) == bytes(ao.h0)
    _same_proof(O, a.pi, ao.pi)


def test_ivc_chain_64_steps_2_12_equals_oracle_chain(ctx, oracle):
    """benches/acc.rs:64-98 with k = 64 at n = 2^12: the GPU chain and the oracle chain are run side by side on the same
    draws; every instance proof and every accumulator (C_bar, d, z, v, pi, pi_V) must be equal at every step, the k verifier
    calls and the decider must accept on both sides."""
    from halo_accumulation_b200 import acc, group, pcdl

    O = oracle
    lg, k = 12, 64
    n, d = 1 << lg, (1 << lg) - 1
    T = _threads(O)
    ctx.derive_generators(1 << 16)
    S, H = ctx.get_SH()
    O.set_params(S, H, ctx.get_generators(0, 1 << 16))
    rng = np.random.Generator(np.random.PCG64(3))  # seed 3
    a = ao = None
    qss, qsos, accs, accos = [], [], [], []
    for s in range(k):
        dp = int(rng.integers(d // 2, d))  # benches/acc.rs:16
        p, w, z, wb, q = _rs(rng, dp + 1), _rs(rng, 1)[0], _rs(rng, 1)[0], _rs(rng, 1)[0], _rs(rng, dp)
        Cm = pcdl.commit(ctx, p, d, w)
        v = group.scalar_dot(ctx, p, group.construct_powers(ctx, z, dp + 1))
        pi = pcdl.open(ctx, p, Cm, d, z, w, q, wb)
        Co = O.pcdl_commit(p, d, w, threads=T)
        assert O.pt_eq(Cm, Co) and v.tolist() == O.scalar_dot(p, O.construct_powers(z, dp + 1)).tolist()
        pio = O.pcdl_open(p, Co, d, z, w, q, wb, threads=T)
        _same_proof(O, pi, pio)
        q_new, qo_new = acc.new_instance(Cm, d, z, v, pi), O.make_instance(Co, d, z, v, pio)
        qs = [acc.to_instance(a), q_new] if a is not None else [q_new]
        qso = [O.acc_to_instance(ao), qo_new] if ao is not None else [qo_new]
        h0, w2, q2, wb2 = _rs(rng, 2), _rs(rng, 1)[0], _rs(rng, n - 1), _rs(rng, 1)[0]
        a = acc.prover(ctx, d, qs, h0, w2, q2, wb2)
        ao = O.acc_prover(d, qso, h0, w2, q2, wb2, threads=T)
        _same_acc(O, a, ao)
        qss.append(qs), qsos.append(qso), accs.append(a), accos.append(ao)
    for qs, qso, ac, aco in zip(qss, qsos, accs, accos):  # the fast path: k verifiers ...
        acc.verifier(ctx, d, qs, ac)
        assert O.acc_verifier(d, qso, aco) == 0
    acc.decider(ctx, accs[-1])                               # ... and one decider
    assert O.acc_decider(accos[-1], threads=T) == 0
    assert O.acc_decider(O.Accumulator.from_buffer_copy(bytes(accs[-1])), threads=T) == 0


def test_ivc_chain_2_20_oracle_verifier_and_decider(ctx, oracle):
    """Config 4 at its real size.  The oracle cannot re-run provers at 2^20 in seconds, but its verifier is O(lg n) group
    operations and its decider one 2^20 MSM: after k = 6 GPU accumulation steps the ORACLE must accept steps 0, k/2 and
    k - 1 through acc_verifier and the final accumulator through acc_decider (scripts/ivc_chain.py does the same at k = 64),
    and a corrupted final accumulator must be rejected by both with the same code."""
    from halo_accumulation_b200 import acc, group, pcdl

    O = oracle
    lg, k = 20, 6
    n, d = 1 << lg, (1 << lg) - 1
    T = _threads(O)
    ctx.derive_generators(n)
    try:
        ctx.precompute_generators(0)
        S, H = ctx.get_SH()
        O.set_params(S, H, ctx.get_generators(0, n))
        rng = np.random.Generator(np.random.PCG64(3))
        a = None
        qss, accs = [], []
        for s in range(k):
            dp = int(rng.integers(d // 2, d))
            p, w, z, wb, q = _rs(rng, dp + 1), _rs(rng, 1)[0], _rs(rng, 1)[0], _rs(rng, 1)[0], _rs(rng, dp)
            Cm = pcdl.commit(ctx, p, d, w)
            v = group.scalar_dot(ctx, p, group.construct_powers(ctx, z, dp + 1))
            pi = pcdl.open(ctx, p, Cm, d, z, w, q, wb)
            q_new = acc.new_instance(Cm, d, z, v, pi)
            qs = [acc.to_instance(a), q_new] if a is not None else [q_new]
            a = acc.prover(ctx, d, qs, _rs(rng, 2), _rs(rng, 1)[0], _rs(rng, n - 1), _rs(rng, 1)[0])
            qss.append(qs), accs.append(a)
        for qs, ac in zip(qss, accs):
            acc.verifier(ctx, d, qs, ac)
        acc.decider(ctx, accs[-1])
        as_o = lambda x, T_: T_.from_buffer_copy(bytes(x))
        for s in (0, k // 2, k - 1):
            assert O.acc_verifier(d, [as_o(q, O.Instance) for q in qss[s]], as_o(accs[s], O.Accumulator)) == 0, f"oracle verifier, step {s}"
        assert O.acc_decider(as_o(accs[-1], O.Accumulator), threads=T) == 0, "oracle decider rejects the GPU chain"
        bad = type(accs[-1]).from_buffer_copy(bytes(accs[-1]))
        bad.v[0] ^= 1
        with pytest.raises(acc.Rejected) as e:
            acc.decider(ctx, bad)
        assert e.value.code == O.acc_decider(as_o(bad, O.Accumulator), threads=T)
    finally:
        ctx.derive_generators(1 << 16)
        S, H = ctx.get_SH()
        O.set_params(S, H, ctx.get_generators(0, 1 << 16))


def test_two_contexts_open_concurrently_from_two_threads(halo, oracle):
    """one context per host thread.  Two contexts on the same device run hiding openings (deferred head
    rounds, k_fold_multi with its per-context operation list) and accumulation steps at the same time; each proof equals
    the oracle's."""
    from halo_accumulation_b200 import pcdl

    O = oracle
    n, d = 1 << 12, (1 << 12) - 1
    ctxs = [halo.Context(0, 1 << 12) for _ in range(2)]
    try:
        for c in ctxs:
            c.derive_generators(n)
            c.set_tuning("ipa_defer_rounds", 3)
        S, H = ctxs[0].get_SH()
        O.set_params(S, H, ctxs[0].get_generators(0, n))
        jobs = []
        for t in range(2):
            for r in range(4):
                p = O.random_scalars(n - 7 * t - r, 100 * t + r)
                z, w, wb = O.random_scalars(3, 1000 + 100 * t + r)
                q = O.random_scalars(p.shape[0] - 1, 2000 + 100 * t + r)
                jobs.append((t, p, z, w, q, wb))
        expected = {}
        for j, (t, p, z, w, q, wb) in enumerate(jobs):
            expected[j] = O.pcdl_open(p, O.pcdl_commit(p, d, w, threads=8), d, z, w, q, wb, threads=8)
        got, errs = {}, []
        start = threading.Barrier(2)

        def work(t):
            try:
                start.wait()
                for rep in range(3):
                    for j, (tt, p, z, w, q, wb) in enumerate(jobs):
                        if tt != t:
                            continue
                        Cm = pcdl.commit(ctxs[t], p, d, w)
                        got[(j, rep)] = pcdl.open(ctxs[t], p, Cm, d, z, w, q, wb)
            except Exception as e:  # surfaced below
                errs.append(e)

        th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
        [x.start() for x in th]
        [x.join() for x in th]
        assert not errs, errs
        assert len(got) == 3 * len(jobs)
        for (j, rep), pi in got.items():
            _same_proof(O, pi, expected[j])
    finally:
        for c in ctxs:
            c.close()
        S = H = None
