"""Form T probe: resident K1t and the host-streaming call on C x T evaluations of the build's generator, beside K1q on the same
tensor.  python examples/tuple_probe.py [C] [T]   (one GPU; prints one JSON line)"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("senweaver-ide_b200")


def main():
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 2_500_000
    K, seed = max(1, C // 4), 0x5EED0001
    nthreads = min(16, len(os.sched_getaffinity(0)))
    eng = pkg.Engine(0)
    eng.dims_generate_compact(seed, 0, C, 0, T, 300)
    eng.corpus_generate(seed, 0, T, 300)
    ref = [eng.score(C, K, corpus=True) for _ in range(4)][-1]
    kq = ref.timing.reward_ms
    ref_sums = eng.debug_partials(C)
    book, d2book = eng.dims_codebook(), eng.dims_d2book()
    pc, pd = pkg.host_empty((C, T), np.uint32), pkg.host_empty((C, T), np.uint16)
    for c in range(C):
        eng.dims_packed_download(c, 0, T, out=(pc[c], pd[c]))
    tl, th = pkg.host_empty((C, T), np.uint16), pkg.host_empty((C, T), np.uint8)
    t0 = time.perf_counter()
    _, _, tbook = pkg.tuple_encode_host(pc, pd, nthreads=nthreads, out=(tl, th))
    enc_s = time.perf_counter() - t0
    idx = tl[0].astype(np.uint32) | (th[0].astype(np.uint32) << 16)
    hot_cover = float((idx < 24576).mean())
    eng.tuples_upload(tl, th, tbook, book, d2book)
    res = [eng.score(C, K, source=pkg.SRC_TUPLES, corpus=True) for _ in range(8)]
    k1t = float(np.mean([x.timing.reward_ms for x in res[3:]]))
    tot = float(np.mean([x.timing.total_ms for x in res[3:]]))
    same = eng.debug_partials(C) == ref_sums and np.array_equal(res[-1].scores, ref.scores) and np.array_equal(res[-1].topk, ref.topk)
    hrec = pkg.host_empty((T,), pkg.RECORD_DTYPE)
    hrec[:] = eng.corpus_download(0, T)

    def step():
        eng.corpus_upload(hrec)
        return eng.score_host_tuples(tl, th, tbook, book, d2book, K, corpus=True)

    step()
    a = time.perf_counter()
    for _ in range(3):
        r = step()
    e2e_ms = (time.perf_counter() - a) * 1e3 / 3
    same_e2e = eng.debug_partials(C) == ref_sums and np.array_equal(r.topk, ref.topk)
    print(json.dumps({
        "C": C, "T": T, "distinct_evaluations": int(tbook[0].size), "head_coverage_cand0": hot_cover, "host_encode_s": round(enc_s, 2),
        "encode_threads": nthreads, "k1q_ms": kq, "k1q_Gevals": C * T / kq / 1e6, "k1t_ms": k1t, "k1t_Gevals": C * T / k1t / 1e6,
        "k1t_GBps": 3.0 * C * T / k1t / 1e6, "tuple_step_ms": tot, "resident_identical": bool(same), "e2e_ms": e2e_ms,
        "e2e_Gevals": C * T / e2e_ms / 1e6, "e2e_h2d_GBps": (3.0 * C * T + 32.0 * T) / e2e_ms / 1e6, "e2e_identical": bool(same_e2e)}))
    eng.close()


if __name__ == "__main__":
    main()
