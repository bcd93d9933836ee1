#!/usr/bin/env python
"""Generates csrc/fp_mul_asm.cuh: the sm_100a Montgomery multiplication / squaring of the two Pasta fields as
straight-line PTX, and proves the instruction list on a Python emulator of the PTX carry-flag semantics first.

Algorithm (K1): CIOS over 32-bit limbs with the partial products split into an "even" array X
(64-bit words at limb positions 0,2,4,6) and an "odd" array Y (words at positions 1,3,5,7) so that every row is
two carry chains of `mad.lo.cc / madc.hi.cc` pairs, each pair fused by ptxas into one IMAD.WIDE.U32(.X).
Round i:   X += a_even * b_i ;  Y += a_odd * b_i ;  m = -X[0]  (since -p^-1 = -1 mod 2^32) ;
           X += m * (1, P2, 0, 0) ;  Y += m * (P1, P3, 0, 2^30) ;  shift one limb: (X, Y) <- (Y, X >> 64) and the
           stray limb X[1] is added into the new X[0] with its carry feeding the next Y chain.
The zero limbs of p cost either an IMAD with a zero multiplier (2 limbs of carry propagation on the FMA pipe) or two
ADDC on the ALU pipe; `--variant` picks the balance (measured on B200, see profiles/).

Usage: python tools/gen_fp_mul.py [--variant N] [--check-only]
"""
import argparse
import os
import random

MASK = 0xFFFFFFFF
FIELDS = {
    "FqParams": 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001,
    "FrParams": 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001,
}


class Prog:
    """A straight-line PTX program over named 32-bit virtual registers."""

    def __init__(self):
        self.ins = []
        self.tmp = 0
        self.zero = set()      # virtual registers known to hold 0 and never materialised
        self.lazy_zero = False  # variant 4: use the literal 0 instead of a zeroed register

    def t(self, hint="t"):
        self.tmp += 1
        return f"{hint}{self.tmp}"

    def emit(self, op, *args):
        if self.lazy_zero:
            if op == "mov" and args[1] == 0:
                self.zero.add(args[0])
                return
            n_pred = 1 if op == "selp" else 0
            src = [0 if (isinstance(a, str) and a in self.zero) else a for a in args[1:len(args) - n_pred]]
            args = (args[0],) + tuple(src) + tuple(args[len(args) - n_pred:] if n_pred else ())
            if op != "setp.ne":
                self.zero.discard(args[0])
        self.ins.append((op,) + args)


def emulate(prog, env):
    """Executes the program on Python ints with an explicit carry flag; mirrors the PTX ISA definitions."""
    cf = 0
    pred = {}

    def val(x):
        return x if isinstance(x, int) else env[x]

    for ins in prog.ins:
        op = ins[0]
        if op in ("mul.lo", "mul.hi"):
            d, a, b = ins[1:]
            pr = val(a) * val(b)
            env[d] = (pr & MASK) if op == "mul.lo" else (pr >> 32) & MASK
        elif op in ("mad.lo.cc", "madc.lo.cc", "madc.hi.cc", "madc.hi", "mad.hi.cc", "madc.lo"):
            d, a, b, c = ins[1:]
            pr = val(a) * val(b)
            part = (pr & MASK) if ".lo" in op else (pr >> 32) & MASK
            s = part + val(c) + (cf if op.startswith("madc") else 0)
            env[d] = s & MASK
            if op.endswith(".cc"):
                cf = s >> 32
        elif op in ("add.cc", "addc.cc", "addc", "add"):
            d, a, b = ins[1:]
            s = val(a) + val(b) + (cf if op.startswith("addc") else 0)
            env[d] = s & MASK
            if op.endswith(".cc"):
                cf = s >> 32
        elif op in ("sub.cc", "subc.cc", "subc", "sub"):
            d, a, b = ins[1:]
            s = val(a) - val(b) - (cf if op.startswith("subc") else 0)  # CF holds the borrow after sub.cc
            env[d] = s & MASK
            if op.endswith(".cc"):
                cf = 1 if s < 0 else 0
        elif op == "shl":
            d, a, k = ins[1:]
            env[d] = (val(a) << k) & MASK
        elif op == "shr":
            d, a, k = ins[1:]
            env[d] = val(a) >> k
        elif op == "not":
            d, a = ins[1:]
            env[d] = (~val(a)) & MASK
        elif op == "shf.r.clamp":  # funnel shift right: lower 32 bits of (hi:lo) >> k
            d, lo, hi, k = ins[1:]
            env[d] = ((((val(hi) << 32) | val(lo)) >> k) & MASK)
        elif op == "shf.l":  # funnel shift left: upper 32 bits of (hi:lo) << k
            d, lo, hi, k = ins[1:]
            env[d] = (((val(hi) << 32) | val(lo)) << k >> 32) & MASK
        elif op == "mov":
            d, a = ins[1:]
            env[d] = val(a)
        elif op == "setp.ne":  # p = (a != b)
            p, a, b = ins[1:]
            pred[p] = val(a) != val(b)
        elif op == "selp":  # d = p ? a : b
            d, a, b, p = ins[1:]
            env[d] = val(a) if pred[p] else val(b)
        else:
            raise ValueError(op)
    return env


def chain_mad(pg, acc, mults, scalar, carry_in, top, fresh_top):
    """acc[2k], acc[2k+1] += mults[k] * scalar as one carry chain; the final carry lands in `top`."""
    first = True
    for k, m in enumerate(mults):
        lo, hi = acc[2 * k], acc[2 * k + 1]
        if first and not carry_in:
            pg.emit("mad.lo.cc", lo, m, scalar, lo)
        else:
            pg.emit("madc.lo.cc", lo, m, scalar, lo)
        pg.emit("madc.hi.cc", hi, m, scalar, hi)
        first = False
    if fresh_top:
        pg.emit("addc", top, 0, 0)
    else:
        pg.emit("addc", top, top, 0)


def chain_pairs(pg, acc, pairs, scalar, carry_in, top):
    """acc[2k], acc[2k+1] += mult * scalar for (k, mult) in pairs (consecutive k up to 3), one carry chain -> top."""
    first = True
    for k, m in pairs:
        lo, hi = acc[2 * k], acc[2 * k + 1]
        pg.emit("madc.lo.cc" if (carry_in or not first) else "mad.lo.cc", lo, m, scalar, lo)
        pg.emit("madc.hi.cc", hi, m, scalar, hi)
        first = False
    pg.emit("addc", top, top, 0)


def build_mul(p, variant, square=False):
    """Returns a Prog computing r = a * b * 2^-256 mod p on registers a0..a7, b0..b7 -> r0..r7."""
    P = [(p >> (32 * i)) & MASK for i in range(8)]
    assert P[0] == 1 and P[4] == P[5] == P[6] == 0 and P[7] == 0x40000000
    P1, P2, P3 = P[1], P[2], P[3]
    pg = Prog()
    pg.lazy_zero = variant == 4
    a = [f"a{i}" for i in range(8)]
    b = [f"a{i}" for i in range(8)] if square else [f"b{i}" for i in range(8)]
    X = [pg.t("x") for _ in range(9)]
    Y = [pg.t("y") for _ in range(9)]
    stray = None
    if square:
        # limbs of 2a: d[j] = (a_j << 1) | (a_{j-1} >> 31); e[j] = a_j << 1 (no incoming bit)
        d = {j: pg.t("d") for j in range(1, 8)}
        e = {j: pg.t("e") for j in range(1, 8)}
        for j in range(1, 8):
            pg.emit("shf.l", d[j], a[j - 1], a[j], 1)
            pg.emit("shl", e[j], a[j], 1)

        def mult(i, j):  # multiplicand at relative position j of row i (j >= i): a_i^2 once, cross terms doubled
            return a[i] if j == i else (e[j] if j == i + 1 else d[j])
    for i in range(8):
        bi = b[i]
        if square and i == 0:
            for k in range(4):
                pg.emit("mul.lo", X[2 * k], mult(0, 2 * k), bi)
                pg.emit("mul.hi", X[2 * k + 1], mult(0, 2 * k), bi)
                pg.emit("mul.lo", Y[2 * k], mult(0, 2 * k + 1), bi)
                pg.emit("mul.hi", Y[2 * k + 1], mult(0, 2 * k + 1), bi)
            pg.emit("mov", X[8], 0)
            pg.emit("mov", Y[8], 0)
        elif square:
            # row i only holds the products a_i * (2a)_j with j >= i (a^2 = sum a_i^2 + 2 sum_{i<j} a_i a_j)
            xp = [(j // 2, mult(i, j)) for j in range(i, 8) if j % 2 == 0]
            yp = [((j - 1) // 2, mult(i, j)) for j in range(i, 8) if j % 2 == 1]
            pg.emit("add.cc", X[0], X[0], stray)
            if yp and yp[0][0] == 0:
                chain_pairs(pg, Y, yp, bi, True, Y[8])  # the stray limb's carry enters Y[0] (relative position 1)
                if xp:
                    chain_pairs(pg, X, xp, bi, False, X[8])
            else:
                # the carry is absorbed by X[1] and rippled up to where the X chain starts
                first_x = 2 * xp[0][0] if xp else 8
                for l in range(1, first_x):
                    pg.emit("addc.cc", X[l], X[l], 0)
                if xp:
                    chain_pairs(pg, X, xp, bi, True, X[8])
                else:
                    pg.emit("addc", X[8], X[8], 0)
                if yp:
                    chain_pairs(pg, Y, yp, bi, False, Y[8])
        elif i == 0:
            for k in range(4):
                pg.emit("mul.lo", X[2 * k], a[2 * k], bi)
                pg.emit("mul.hi", X[2 * k + 1], a[2 * k], bi)
                pg.emit("mul.lo", Y[2 * k], a[2 * k + 1], bi)
                pg.emit("mul.hi", Y[2 * k + 1], a[2 * k + 1], bi)
            pg.emit("mov", X[8], 0)
            pg.emit("mov", Y[8], 0)
        else:
            # the stray limb (old X[1], position 0) joins X[0]; its carry enters the Y chain at position 1
            pg.emit("add.cc", X[0], X[0], stray)
            chain_mad(pg, Y, [a[1], a[3], a[5], a[7]], bi, True, Y[8], False)
            chain_mad(pg, X, [a[0], a[2], a[4], a[6]], bi, False, X[8], False)
        # m = -X[0]
        m = pg.t("m")
        if variant == 4:
            nx = pg.t("n")
            pg.emit("not", nx, X[0])
            pg.emit("add", m, nx, 1)
        else:
            pg.emit("sub", m, 0, X[0])
        # X += m * (1, P2, 0, 0): X[0] cancels to zero
        if variant == 4:
            pg.emit("add.cc", X[0], X[0], 0xFFFFFFFF)  # carry = (X[0] != 0); the limb itself is dropped by the shift
        else:
            pg.emit("add.cc", X[0], X[0], m)
        pg.emit("addc.cc", X[1], X[1], 0)
        pg.emit("madc.lo.cc", X[2], m, P2, X[2])
        pg.emit("madc.hi.cc", X[3], m, P2, X[3])
        if variant in (1, 3):  # zero limbs of p: carry propagation by IMAD with a zero multiplier (FMA pipe)
            pg.emit("madc.lo.cc", X[4], m, 0, X[4])
            pg.emit("madc.hi.cc", X[5], m, 0, X[5])
        else:
            pg.emit("addc.cc", X[4], X[4], 0)
            pg.emit("addc.cc", X[5], X[5], 0)
        if variant == 3:
            pg.emit("madc.lo.cc", X[6], m, 0, X[6])
            pg.emit("madc.hi.cc", X[7], m, 0, X[7])
        else:
            pg.emit("addc.cc", X[6], X[6], 0)
            pg.emit("addc.cc", X[7], X[7], 0)
        pg.emit("addc", X[8], X[8], 0)
        # Y += m * (P1, P3, 0, 2^30)
        pg.emit("mad.lo.cc", Y[0], m, P1, Y[0])
        pg.emit("madc.hi.cc", Y[1], m, P1, Y[1])
        pg.emit("madc.lo.cc", Y[2], m, P3, Y[2])
        pg.emit("madc.hi.cc", Y[3], m, P3, Y[3])
        if variant in (1, 2, 3):
            pg.emit("madc.lo.cc", Y[4], m, 0, Y[4])
            pg.emit("madc.hi.cc", Y[5], m, 0, Y[5])
            pg.emit("madc.lo.cc", Y[6], m, 0x40000000, Y[6])
            pg.emit("madc.hi.cc", Y[7], m, 0x40000000, Y[7])
        elif variant == 4:
            lo30, hi2 = pg.t("s"), pg.t("s")
            pg.emit("shf.l", lo30, 0, m, 30)       # (m:0) << 30 upper word = m << 30
            pg.emit("shf.r.clamp", hi2, m, 0, 2)   # (0:m) >> 2 lower word = m >> 2
            pg.emit("addc.cc", Y[4], Y[4], 0)
            pg.emit("addc.cc", Y[5], Y[5], 0)
            pg.emit("addc.cc", Y[6], Y[6], lo30)
            pg.emit("addc.cc", Y[7], Y[7], hi2)
        else:
            lo30, hi2 = pg.t("s"), pg.t("s")
            pg.emit("shl", lo30, m, 30)
            pg.emit("shr", hi2, m, 2)
            pg.emit("addc.cc", Y[4], Y[4], 0)
            pg.emit("addc.cc", Y[5], Y[5], 0)
            pg.emit("addc.cc", Y[6], Y[6], lo30)
            pg.emit("addc.cc", Y[7], Y[7], hi2)
        pg.emit("addc", Y[8], Y[8], 0)
        # shift one limb: new X = Y ; new Y = X[2..8], 0, 0 ; stray = X[1]
        stray = X[1]
        newY = X[2:9] + [pg.t("y"), pg.t("y")]
        pg.emit("mov", newY[7], 0)
        pg.emit("mov", newY[8], 0)
        X, Y = Y, newY
    # T = X + stray + (Y << 32), < 2p < 2^256
    T = [pg.t("r") for _ in range(8)]
    pg.emit("add.cc", T[0], X[0], stray)
    for k in range(1, 8):
        pg.emit("addc.cc" if k < 7 else "addc", T[k], X[k], Y[k - 1])
    # r = T >= p ? T - p : T
    D = [pg.t("d") for _ in range(8)]
    pg.emit("sub.cc", D[0], T[0], P[0])
    for k in range(1, 8):
        pg.emit("subc.cc", D[k], T[k], P[k])
    bw = pg.t("bw")
    pg.emit("subc", bw, 0, 0)  # 0xffffffff if T < p
    pg.emit("setp.ne", "pb", bw, 0)
    for k in range(8):
        pg.emit("selp", f"r{k}", T[k], D[k], "pb")
    return pg


def check(p, variant, square, iters=3000):
    pg = build_mul(p, variant, square)
    rnd = random.Random(12345)
    rinv = pow(1 << 256, -1, p)
    edge = [0, 1, 2, p - 1, p - 2, (1 << 256) % p, (1 << 255) % p, MASK, 1 << 32, 1 << 254, p >> 1, (p >> 1) + 1, (1 << 64) - 1,
            (1 << 128) - 1, (1 << 224) - 1, (1 << 254) - 1]
    cases = [(x, y) for x in edge for y in edge] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(iters)]
    for x, y in cases:
        if square:
            y = x
        env = {}
        for i in range(8):
            env[f"a{i}"] = (x >> (32 * i)) & MASK
            env[f"b{i}"] = (y >> (32 * i)) & MASK
        emulate(pg, env)
        r = sum(env[f"r{i}"] << (32 * i) for i in range(8))
        assert r == x * y * rinv % p, (hex(x), hex(y), hex(r))
    return pg


def to_ptx(pg, square):
    """One asm statement: outputs %0..%7 = r, inputs %8..%15 = a, %16..%23 = b."""
    opnum = {}
    for i in range(8):
        opnum[f"r{i}"] = f"%{i}"
        opnum[f"a{i}"] = f"%{8 + i}"
        if not square:
            opnum[f"b{i}"] = f"%{16 + i}"
    temps = []

    def reg(x):
        if isinstance(x, int):
            return str(x)
        if x in opnum:
            return opnum[x]
        if x not in temps:
            temps.append(x)
        return x

    lines = []
    for ins in pg.ins:
        op = ins[0]
        if op == "setp.ne":
            lines.append(f"setp.ne.u32 {ins[1]}, {reg(ins[2])}, {reg(ins[3])};")
        elif op == "selp":
            lines.append(f"selp.u32 {reg(ins[1])}, {reg(ins[2])}, {reg(ins[3])}, {ins[4]};")
        elif op in ("shl", "shr"):
            lines.append(f"{op}.b32 {reg(ins[1])}, {reg(ins[2])}, {ins[3]};" if op == "shl" else f"shr.u32 {reg(ins[1])}, {reg(ins[2])}, {ins[3]};")
        elif op == "not":
            lines.append(f"not.b32 {reg(ins[1])}, {reg(ins[2])};")
        elif op == "shf.r.clamp":
            lines.append(f"shf.r.clamp.b32 {reg(ins[1])}, {reg(ins[2])}, {reg(ins[3])}, {ins[4]};")
        elif op == "shf.l":
            lines.append(f"shf.l.wrap.b32 {reg(ins[1])}, {reg(ins[2])}, {reg(ins[3])}, {ins[4]};")
        elif op == "mov":
            lines.append(f"mov.u32 {reg(ins[1])}, {reg(ins[2])};")
        else:
            lines.append(f"{op}.u32 " + ", ".join(reg(x) for x in ins[1:]) + ";")
    decl = [".reg .pred pb;"]
    for k in range(0, len(temps), 12):
        decl.append(".reg .u32 " + ", ".join(temps[k:k + 12]) + ";")
    return decl + lines


def emit_function(name, params, variant, pg, square):
    body = to_ptx(pg, square)
    s = []
    sig = "const uint32_t a[8]" if square else "const uint32_t a[8], const uint32_t b[8]"
    s.append("template <>")
    s.append(f"__device__ __forceinline__ void {name}<{params}, {variant}>(uint32_t r[8], {sig}) {{")
    s.append('    asm("{\\n\\t"')
    for ln in body:
        s.append(f'        "{ln}\\n\\t"')
    s.append('        "}"')
    outs = ", ".join(f'"=r"(r[{i}])' for i in range(8))
    ins = ", ".join(f'"r"(a[{i}])' for i in range(8))
    if not square:
        ins += ", " + ", ".join(f'"r"(b[{i}])' for i in range(8))
    s.append(f"        : {outs}")
    s.append(f"        : {ins});")
    s.append("}")
    return "\n".join(s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check-only", action="store_true")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc", "fp_mul_asm.cuh"))
    args = ap.parse_args()
    out = [
        "// fp_mul_asm.cuh -- GENERATED by tools/gen_fp_mul.py; do not edit.",
        "// Montgomery multiplication / squaring of the Pasta fields as straight-line PTX for sm_100a: CIOS over 32-bit limbs,",
        "// even/odd partial-product arrays, every mad.lo.cc/madc.hi.cc pair fused by ptxas into one IMAD.WIDE.U32(.X).",
        "// Variants differ in how the zero limbs of p propagate carries (IMAD with a zero multiplier vs ADDC); see the generator.",
        "// Every instruction list was executed on the generator's PTX emulator against a * b * 2^-256 mod p before emission.",
        "#pragma once",
        "#include <stdint.h>",
        "namespace halo {",
        "struct FqParams;",
        "struct FrParams;",
        "template <class P, int V> __device__ __forceinline__ void fp_mul_asm(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]);",
        "template <class P, int V> __device__ __forceinline__ void fp_sqr_asm(uint32_t r[8], const uint32_t a[8]);",
    ]
    for variant in (0, 1, 2, 3, 4):
        for params, p in FIELDS.items():
            for square in (False, True):
                pg = check(p, variant, square, iters=1500)
                out.append(emit_function("fp_sqr_asm" if square else "fp_mul_asm", params, variant, pg, square))
        ops = [i[0] for i in pg.ins]
        n_mad = sum(1 for o in ops if o.startswith(("mad", "mul")))
        print(f"variant {variant}: verified on the emulator; {len(ops)} PTX instructions, {n_mad} mad/mul")
    out.append("}  // namespace halo")
    if args.check_only:
        return
    with open(args.out, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", os.path.normpath(args.out))


if __name__ == "__main__":
    main()
