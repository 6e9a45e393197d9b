"""evals/s of one resident apo_score call (fp32 Form D, K = C/4, no corpus) across problem sizes, one B200.
    python examples/size_sweep.py > sweep.jsonl
Shows where the call stops being launch-bound and reaches the HBM stream rate."""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("senweaver-ide_b200")
eng = pkg.Engine(0)
for C, T in [(4, 1000), (16, 10_000), (64, 10_000), (64, 100_000), (256, 100_000), (64, 1_000_000), (256, 1_000_000), (1024, 1_000_000),
             (256, 4_000_000), (256, 10_000_000)]:
    eng.dims_generate(0x5EED0042, 0, C, 0, T, 300)
    K = max(1, C // 4)
    eng.score(C, K)
    reps = 200 if C * T < 10**8 else 10
    t0 = time.perf_counter()
    for _ in range(reps):
        r = eng.score(C, K)
    wall = (time.perf_counter() - t0) / reps
    rt = eng.score(C, K, timing=True)                 # per-stage CUDA events on request (a small call does not record them by default)
    print(json.dumps({"C": C, "T": T, "bytes": 36 * C * T, "wall_us_per_call": wall * 1e6, "device_ms": rt.timing.total_ms,
                      "evals_per_s_wall": C * T / wall, "GBps_wall": 36 * C * T / wall / 1e9, "launches": r.timing.launches}), flush=True)
