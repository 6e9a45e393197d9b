import importlib, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("senweaver-ide_b200")
eng = pkg.Engine(0)
C, T = 64, 1_000_000
eng.dims_generate(0x5EED0002, 0, C, 0, T, 300)
host = np.stack([eng.dims_download(c, 0, T) for c in range(C)])          # pageable numpy
pin = pkg.host_empty((C, T, 9), np.float32)
pin[:] = host
for name, buf, env in (("pageable, driver staging (APO_TUNE_NO_STAGING)", host, "1"), ("pageable, library staging", host, None), ("pinned (apo_host_alloc)", pin, None)):
    eng.set_tuning(pkg.TUNE_NO_STAGING if env else 0)
    eng.score_host(buf, 16)
    t0 = time.perf_counter()
    for _ in range(5):
        r = eng.score_host(buf, 16)
    dt = (time.perf_counter() - t0) / 5
    print(name, f"{dt*1e3:.1f} ms/step  {C*T/dt/1e9:.2f} G evals/s  {C*T*36/dt/1e9:.1f} GB/s")
