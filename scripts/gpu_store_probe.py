"""Generator store against on-device derivation: seconds to derive, save, and load n generators."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_accumulation_b200 as H  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << lg
ctx = H.Context(0, n)
ctx.derive_generators(1024)  # warm-up: fixed-base table of the derivation
t = time.perf_counter(); ctx.derive_generators(n); t_derive = time.perf_counter() - t
path = os.path.join(tempfile.gettempdir(), "halo_gens.bin")
t = time.perf_counter(); ctx.save_generators(path); t_save = time.perf_counter() - t
first = ctx.get_generators(n - 4, 4)
ctx.derive_generators(16)
t = time.perf_counter(); ctx.load_generators_file(path); t_load = time.perf_counter() - t
assert (ctx.get_generators(n - 4, 4) == first).all()
t = time.perf_counter(); ctx.load_generators_file(path); t_load2 = time.perf_counter() - t
size = os.path.getsize(path)
os.remove(path)
print(json.dumps({"curve": H._build.CURVE, "n": n, "file_bytes": size, "derive_s": t_derive, "save_s": t_save, "load_s": t_load,
                  "load_again_s": t_load2, "load_GBps": size / t_load2 / 1e9,
                  "note": "file in the box's temp directory (page cache warm after the save); load includes checksum and the on-device on-curve check"}))
ctx.close()
