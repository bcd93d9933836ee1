"""Generates tests/golden/consts_golden.npz from the reference's only golden data,
code/src/consts.rs (S, H, GS[0..16384] as Montgomery u64x4 limbs; consts.rs:26-16453).

Run once in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The fixture keeps S, H, a 2 560-point subset of GS verbatim (indices listed in `gs_idx`) and the
SHA-256 of the full 16 384 x 64-byte little-endian limb array, so the complete set stays pinned
without committing 1 MiB of limbs.
"""
import hashlib
import os
import re
import sys

import numpy as np

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/code/src/consts.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "consts_golden.npz")

text = open(SRC).read()
proj = re.findall(r"pub const (S|H): Projective = mk_proj!\(\s*\[([^\]]*)\],\s*\[([^\]]*)\],\s*\[([^\]]*)\]\s*\)", text)
assert [p[0] for p in proj] == ["S", "H"], proj
to_limbs = lambda s: [int(t) for t in s.replace("\n", " ").split(",") if t.strip()]
S = np.array([to_limbs(c) for c in proj[0][1:]], dtype=np.uint64).reshape(12)
H = np.array([to_limbs(c) for c in proj[1][1:]], dtype=np.uint64).reshape(12)
aff = re.findall(r"mk_aff!\(\[([^\]]*)\],\s*\[([^\]]*)\]\)", text)
assert len(aff) == 16384, len(aff)
GS = np.array([to_limbs(x) + to_limbs(y) for x, y in aff], dtype=np.uint64)
assert GS.shape == (16384, 8)
full_sha = hashlib.sha256(GS.astype("<u8").tobytes()).hexdigest()
idx = np.unique(np.concatenate([np.arange(0, 1024), np.arange(1024, 16384, 16), np.arange(16384 - 576, 16384)]))
np.savez_compressed(OUT, S=S, H=H, gs_idx=idx.astype(np.int64), gs=GS[idx], gs_full_sha256=np.array(full_sha), n=np.array(16384))
print("wrote", OUT, "subset", len(idx), "full sha256", full_sha, "bytes", os.path.getsize(OUT))
