"""SASS evidence for DESIGN.md's instruction accounting.

Disassembles lib/libhalo_b200.so with `cuobjdump -sass` (no GPU needed) and, per kernel, counts instruction classes:
IMAD.WIDE (the 32 x 32 -> 64 multiply-adds: half rate on the integer pipe), other IMAD (IMAD.X / IMAD.MOV / IMAD.IADD .. full
rate), IADD3 (+ .X carry chains, ALU pipe), LDG / STG / LDS / STS, shuffles, branches.  For the single-multiplication
microbenchmark k_fp_mul_tp<1, 0> the loop body (between the backward branch target and the branch) is ONE Montgomery
multiplication, so its counts are "per modmul".  Usage: python scripts/sass_counts.py [out.txt] [--excerpt N]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "halo-accumulation_b200", "lib", "libhalo_b200.so")
KERNELS = {
    "k_fp_mul_tp<1,0> (one generated Montgomery multiplication per loop iteration)": "_ZN4halo11k_fp_mul_tpILi1ELi0EEEviPj",
    "k_pair_bwd<PASS0>": "_ZN4halo10k_pair_bwdILb1EEE",
    "k_pair_bwd<stream>": "_ZN4halo10k_pair_bwdILb0EEE",
    "k_pair_fwd<PASS0>": "_ZN4halo10k_pair_fwdILb1EEE",
    "k_accumulate<indirect>": "_ZN4halo12k_accumulateILb0EEE",
    "k_pair_bwd0 (pass 0, cp.async staging)": "_ZN4halo11k_pair_bwd0E",
    "k_fold_multi (multiplication inlined: folds of < 2^17 outputs)": "_ZN4halo12k_fold_multiE",
    "k_fold_multi_call (multiplication out of line, ipa_fold.cu: folds of >= 2^17 outputs)": "_ZN4halo17k_fold_multi_callE",
    "k_accumulate_quad<4, 4> (msm_small.cu: multiplication out of line)": "_ZN4halo17k_accumulate_quadILi4ELi4EEE",
    "k_reduce_slabs_quad (msm_small.cu: multiplication out of line)": "_ZN4halo19k_reduce_slabs_quadE",
    "k_reduce_slabs": "_ZN4halo14k_reduce_slabsE",
}


def classify(op):
    if op.startswith("IMAD.WIDE"):
        return "IMAD.WIDE*"
    if op.startswith("IMAD.HI"):
        return "IMAD.HI*"
    if op.startswith("IMAD"):
        return "IMAD (other)"
    if op.startswith("IADD3"):
        return "IADD3*"
    for p in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LDC", "SHFL", "BRA", "BAR", "LOP3", "SEL", "ISETP", "SHF", "MOV", "ATOM", "RED", "CALL", "RET", "PRMT", "VOTE", "LEA"):
        if op.startswith(p):
            return p
    return "other"


def functions(sass):
    cur, body = None, []
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if cur:
                yield cur, body
            cur, body = m.group(1), []
        elif cur:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                body.append((int(m.group(1), 16), m.group(2), line.strip()))
    if cur:
        yield cur, body


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = dict(functions(sass))
    lines = [f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a), instruction-class counts per kernel (static)"]
    for title, prefix in KERNELS.items():
        name = next((k for k in funcs if k.startswith(prefix)), None)
        if not name:
            lines.append(f"\n## {title}: not found")
            continue
        body = funcs[name]
        cnt = collections.Counter(classify(op) for _, op, _ in body)
        lines.append(f"\n## {title}\n{name}: {len(body)} instructions")
        lines.append("   " + ", ".join(f"{k} {v}" for k, v in cnt.most_common()))
        if "k_fp_mul_tp" in title:
            # innermost loop: the last backward branch and its target
            bra = re.compile(r"BRA\S*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)")
            back = [(a, int(bra.search(l).group(1), 16)) for a, op, l in body if op.startswith("BRA") and bra.search(l)]
            back = [(a, t) for a, t in back if t < a]
            if back:
                a_br, tgt = back[-1]  # the remainder loop (ptxas also emits 4x / 16x unrolled copies of the same body)
                loop = [(a, op, l) for a, op, l in body if tgt <= a <= a_br]
                c2 = collections.Counter(classify(op) for _, op, _ in loop)
                lines.append(f"   loop body {tgt:#x}..{a_br:#x}: {len(loop)} instructions = ONE modular multiplication (+ loop control)")
                lines.append("   " + ", ".join(f"{k} {v}" for k, v in c2.most_common()))
                lines.append("   excerpt (first 24 instructions of the loop):")
                lines += ["      " + l for _, _, l in loop[:24]]
    text = "\n".join(lines) + "\n"
    if out_path:
        open(out_path, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
