"""Randomised parity soak (not collected by pytest): N trials over every resident form and kernel variant against the oracle.

    python tests/soak_parity.py [trials] [seed]

Each trial draws a shape (C, T), a window, K, a kernel variant, a data source — Form D generated / random-with-NaNs, the
compact layout (Form Q), raw records (Form R) or packed records (Form R16) — and, half of the time, a corpus for the
report; it then requires the exact integer partial sums, counts, top-K and (with a corpus) the pattern counts / examples /
tallies to equal the oracle's.  Prints one summary line; exits non-zero on the first mismatch."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402

pkg = importlib.import_module("senweaver-ide_b200")


def main(trials: int, seed: int):
    e = pkg.Engine(0)
    rng = np.random.default_rng(seed)
    kinds = {"fp32": 0, "fp32-nan": 0, "compact": 0, "records": 0, "records16": 0, "host-packed": 0, "host-compact": 0}
    evals = 0
    for trial in range(trials):
        C = int(rng.integers(1, 24))
        T = int(rng.choice([1, 3, 4, 127, 128, 129, 2559, 2560, 2561, 5119, 5120, 5121, 7777, 20_011]))
        K = int(rng.integers(0, C + 1))
        kind = str(rng.choice(list(kinds)))
        ap = int(rng.integers(0, 1025))
        s0, c0, t0 = int(rng.integers(1, 1 << 30)), int(rng.integers(0, 50)), int(rng.integers(0, 1 << 20))
        align = 8 if kind == "compact" else 4
        first = int(rng.integers(0, T // align + 1)) * align if T >= align and rng.random() < 0.4 else 0
        count = int(rng.integers(1, T - first + 1)) if first < T and rng.random() < 0.4 else 0
        if kind.startswith("host-"):
            first = count = 0                                   # the host-streaming calls take the whole tensor
        e.set_tuning(pkg.TUNE_NO_FUSE if rng.random() < 0.3 else 0)      # stand-alone K2 / K3 launches (TMA tiles from 8192 records) or the fused tail
        hi = first + count if count else T
        with_corpus = rng.random() < 0.5
        if with_corpus:
            Tc = int(rng.choice([1, 50, 1000, 4097, 9001]))
            recs = orc.gen_records(s0 ^ 0x55, orc.STREAM_CORPUS, 0, 1, 0, Tc, ap, 2).reshape(-1)
            e.corpus_upload(recs)
        if kind in ("records", "records16"):
            roll = orc.gen_records(s0, orc.STREAM_ROLLOUT, c0, C, t0, T, ap, 4)
            if kind == "records16":
                e.rollouts16_upload(pkg.pack16(roll))
            else:
                e.rollouts_upload(roll)
            variant = int(rng.integers(0, 5))
            r = e.score(C, K, source=1, corpus=with_corpus, variant=variant, first=first, count=count)
            exp = orc.score_records_fx(roll[:, first:hi])
            ref_s, _ = orc.score_records(roll[:, first:hi])
        else:
            if kind == "fp32-nan":
                dims = rng.choice(np.array([-1, -0.8, -0.5, -0.3, -0.2, 0, 0.3, 0.5, 0.8, 1], np.float32), (C, T, 9))
                dims[:, :, 2] = rng.uniform(-1, 1, (C, T)).astype(np.float32)
                dims[rng.random(dims.shape) < rng.random()] = np.nan
            else:
                dims = orc.gen_dims(s0, c0, C, t0, T, ap, 4)
            variant = 0
            if kind == "host-packed":
                pc, pd, book, d2book = pkg.packed_encode_host(dims, nthreads=2)
                r = e.score_host_packed(pc, pd, book, d2book, K, corpus=with_corpus)
            elif kind == "host-compact":
                q8, d2, li, book = pkg.compact_encode_host(dims, nthreads=2)
                r = e.score_host_compact(q8, d2, li, book, K, corpus=with_corpus)
            else:
                e.dims_upload(dims)
                if kind == "compact":
                    e.dims_compact()
                    variant = int(rng.integers(0, 7))
                else:
                    variant = int(rng.integers(0, 5))
                r = e.score(C, K, corpus=with_corpus, variant=variant, first=first, count=count)
            exp = orc.score_dims_fx(dims[:, first:hi])
            ref_s, _ = orc.score_dims(dims[:, first:hi])
        got = e.debug_partials(C)
        ctx = (trial, kind, C, T, K, variant, first, count, with_corpus)
        assert got == exp, ("partials", ctx)
        assert np.array_equal(r.topk, orc.topk(ref_s, K)), ("topk", ctx)
        if with_corpus:
            ref = orc.report(recs)
            rep = r.report
            assert (rep.total, rep.good, rep.bad, rep.none, rep.withReward) == (ref.total, ref.good, ref.bad, ref.none, ref.withReward), ("tallies", ctx)
            for p in range(6):
                assert (rep.pat[p].count, rep.pat[p].flag, list(rep.pat[p].examples)) == \
                       (ref.pat[p].count, ref.pat[p].flag, list(ref.pat[p].examples)), ("pattern", p, ctx)
                if ref.pat[p].flag:                                   # severity is defined for emitted patterns only (APO:650-741)
                    assert rep.pat[p].severity == ref.pat[p].severity, ("severity", p, ctx)
            for i in range(9):
                assert rep.dim[i].count == ref.dim[i].count, ("dimcount", i, ctx)
        kinds[kind] += 1
        evals += C * (hi - first)
    print(f"soak ok: {trials} trials (seed {seed}), {evals} evaluations, by source {kinds}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
