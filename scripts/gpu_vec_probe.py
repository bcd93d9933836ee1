"""Achieved HBM bandwidth of the Fr vector kernels (north-star: 'achieved HBM GB/s for the folds')."""
import json, sys
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
ctx = H.Context(0, 1 << 24)
for lg in (20, 24):
    n = 1 << lg
    rows = {
        "fold_scalars (c and z, one round)": (0, 2 * (n // 2) * 96),   # per vector and output element: 2 x 32 B read + 32 B write
        "h_expand": (1, n * 32),                                       # 32 B written per coefficient
        "scalar_dot": (2, 2 * n * 32),                                 # two 32 B reads per element
        "construct_powers": (3, n * 32),
    }
    for name, (kind, bytes_) in rows.items():
        ms = ctx.test_vec_bench(kind, n)
        gbs = bytes_ / ms / 1e6
        print(json.dumps({"kernel": name, "n": f"2^{lg}", "ms": round(ms, 4), "algorithmic_bytes": bytes_, "GB_s": round(gbs, 1),
                          "frac_of_measured_hbm_peak": round(gbs / peak, 3), "peak_GB_s": peak}))
