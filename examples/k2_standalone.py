"""Device time of the stand-alone corpus scan (K2 + the K3 tail) over 10 M records — what sessions, host streaming and small
calls pay when the scan cannot ride inside a scoring launch.  APO_K2_GRIDSTRIDE=1 selects the old grid-stride kernel."""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("senweaver-ide_b200")
eng = pkg.Engine(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
eng.dims_generate(0x5EED0003, 0, 4, 0, 4096, 300)
eng.corpus_generate(0x5EED0003, 0, T, 300)
ms = []
for i in range(12):
    eng.score_begin(4)
    eng.score_accumulate(0)
    r = eng.score_finish(4, 2, corpus=True)
    if i >= 2:
        ms.append(r.timing.corpus_ms)
print(json.dumps({"records": T, "kernel": "grid-stride" if os.environ.get("APO_K2_GRIDSTRIDE") else "tma-tiles", "corpus_ms": float(np.median(ms)),
                  "GBps": 32.0 * T / (float(np.median(ms)) * 1e-3) / 1e9, "bad": int(r.report.bad), "pat": [int(r.report.pat[p].count) for p in range(6)]}))
