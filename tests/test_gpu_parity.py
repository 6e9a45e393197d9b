"""Parity of the sm_100a path against the CPU oracle, through the C ABI (-m gpu).

Bars (BASELINE.json north_star / SURVEY 8d): top-K indices, pattern flags/counts/examples,
tallies, per-evaluation dims/finalReward and the integer partial sums are BIT-EXACT;
score[c] / avgReward / per-dimension means are within 1e-5 relative of the oracle's
sequential binary64 sums (abs(d) <= 1e-5 * max(abs(ref), 1e-6)); in practice they agree to
~1e-13, which the tests also assert as a tighter secondary bound.
"""
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL, EPS = 1e-5, 1e-6


def close(a, b, rtol=RTOL):
    return abs(a - b) <= rtol * max(abs(b), EPS)


def assert_scores(res, ref_scores, ref_counts, tight=1e-11):
    assert np.array_equal(res.counts, ref_counts)
    for c, (a, b) in enumerate(zip(res.scores, ref_scores)):
        if math.isinf(b):
            assert a == b, c
        else:
            assert close(a, b), (c, a, b)
            assert abs(a - b) <= tight * max(abs(b), 1.0), (c, a, b)


# ------------------------------------------------------------------ single-trace path
def test_reward_batch_golden_bit_exact(engine, orc):
    with open(os.path.join(GOLD, "reward_cases.json")) as f:
        g = json.load(f)
    cases = list(g["kats"].values()) + g["random"]
    recs = np.frombuffer(b"".join(bytes.fromhex(c["record_hex"]) for c in cases), orc.RECORD_DTYPE)
    dims, masks, finals = engine.reward_batch(recs)
    for i, c in enumerate(cases):
        for k in range(9):
            exp = math.nan if c["dims"][k] == "nan" else float.fromhex(c["dims"][k])
            assert (math.isnan(exp) and math.isnan(dims[i, k])) or dims[i, k] == exp, (i, k)
            assert bool(masks[i] & (1 << k)) == (c["dims"][k] != "nan")
        assert finals[i] == float.fromhex(c["finalReward"]), i


def test_reward_batch_generated_bit_exact(engine, orc):
    recs = orc.gen_records(0xABCDEF, orc.STREAM_ROLLOUT, 9, 1, 0, 30000, 500, 8).reshape(-1)
    dims, masks, finals = engine.reward_batch(recs)
    for i in range(0, len(recs), 7):
        d, m, fr = orc.reward_one(recs[i])
        assert m == masks[i]
        assert np.array_equal(np.isnan(d), np.isnan(dims[i])) and np.array_equal(d[~np.isnan(d)], dims[i][~np.isnan(d)])
        assert (fr is None and math.isnan(finals[i])) or fr == finals[i]


# ------------------------------------------------------------------ generators agree bit for bit
def test_generators_match_oracle(engine, orc):
    seed, ap = 0x5EED0003, 300
    engine.corpus_generate(seed, 1000, 5000, ap)
    assert engine.corpus_download(0, 5000).tobytes() == orc.gen_records(seed, orc.STREAM_CORPUS, 0, 1, 1000, 5000, ap, 8).tobytes()
    engine.rollouts_generate(seed, 3, 4, 96, 3001, ap)
    ref = orc.gen_records(seed, orc.STREAM_ROLLOUT, 3, 4, 96, 3001, ap, 8)
    for c in range(4):
        assert engine.rollouts_download(c, 0, 3001).tobytes() == ref[c].tobytes()
    engine.dims_generate(seed, 3, 4, 96, 3001, ap)
    refd = orc.gen_dims(seed, 3, 4, 96, 3001, ap, 8)
    for c in range(4):
        assert engine.dims_download(c, 0, 3001).tobytes() == refd[c].tobytes()
    with open(os.path.join(GOLD, "generator_case.json")) as f:
        g = json.load(f)
    engine.dims_generate(g["seed"], 5, 1, 1000, 64, g["agent_permille"])
    assert engine.dims_download(0, 0, 64).tobytes().hex() == g["dims_c5_t1000"]


# ------------------------------------------------------------------ K1 + top-K, Form D
@pytest.mark.parametrize("C,T,K", [(4, 1000, 4), (3, 1, 1), (5, 4099, 2), (7, 1023, 7), (64, 20001, 16), (2, 1025, 1)])
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
def test_score_dims_matches_oracle(engine, orc, C, T, K, variant):
    dims = orc.gen_dims(0x5EED0000 + C, 0, C, 0, T, 300, 8)
    engine.dims_upload(dims)
    res = engine.score(C, K, variant=variant)
    ref_s, ref_n = orc.score_dims(dims)
    assert_scores(res, ref_s, ref_n)
    assert np.array_equal(res.topk, orc.topk(ref_s, K))
    # the integer partial sums are exact: no tolerance
    sums, counts = engine.debug_partials(C)
    esums, ecounts = orc.score_dims_fx(dims)
    assert sums == esums and counts == ecounts


def test_score_dims_nan_patterns_and_null_candidates(engine, orc):
    rng = np.random.default_rng(5)
    C, T = 6, 3000
    dims = rng.uniform(-1, 1, (C, T, 9)).astype(np.float32)
    dims[rng.random((C, T, 9)) < 0.35] = np.nan          # arbitrary presence masks (all 512 possible)
    dims[2] = np.nan                                      # candidate with no non-null evaluation
    dims[4] = dims[1]                                     # exact duplicate -> tie -> lower index wins
    dims[:, 100:140, :] = np.nan                          # null evaluations
    engine.dims_upload(dims)
    res = engine.score(C, C)
    ref_s, ref_n = orc.score_dims(dims)
    assert_scores(res, ref_s, ref_n)
    assert res.scores[2] == -np.inf and res.counts[2] == 0
    assert res.scores[1] == res.scores[4]
    assert np.array_equal(res.topk, orc.topk(ref_s, C))
    assert list(res.topk).index(1) < list(res.topk).index(4)
    assert res.topk[-1] == 2
    sums, counts = engine.debug_partials(C)
    esums, ecounts = orc.score_dims_fx(dims)
    assert sums == esums and counts == ecounts


def test_score_windows_add_up_exactly(engine, orc):
    C, T = 5, 50000
    dims = orc.gen_dims(77, 0, C, 0, T, 300, 8)
    engine.dims_upload(dims)
    engine.score(C, 1)
    total, tcount = engine.debug_partials(C)
    acc, cnt = [0] * C, [0] * C
    for first, count in [(0, 12000), (12000, 20004), (32004, 17996)]:
        engine.score(C, 1, first=first, count=count)
        s, n = engine.debug_partials(C)
        es, en = orc.score_dims_fx(dims[:, first:first + count])
        assert s == es and n == en
        acc = [a + b for a, b in zip(acc, s)]
        cnt = [a + b for a, b in zip(cnt, n)]
    assert acc == total and cnt == tcount


def test_recip_variant_within_tolerance(engine, orc):
    C, T = 8, 30000
    dims = orc.gen_dims(99, 0, C, 0, T, 300, 8)
    engine.dims_upload(dims)
    res = engine.score(C, 4, recip=True)
    ref_s, ref_n = orc.score_dims(dims)
    assert_scores(res, ref_s, ref_n)
    assert np.array_equal(res.topk, orc.topk(ref_s, 4))


def test_custom_weights(engine, orc):
    w = np.array([0.3, 0.1, 0.05, 0.05, 0.1, 0.1, 0.1, 0.1, 0.1])
    dims = orc.gen_dims(5, 0, 4, 0, 2000, 300, 8)
    try:
        engine.set_weights(w)
        engine.dims_upload(dims)
        res = engine.score(4, 2)
        ref_s, ref_n = orc.score_dims(dims, w=w)
        assert_scores(res, ref_s, ref_n)
        sums, _ = engine.debug_partials(4)
        assert sums == orc.score_dims_fx(dims, w=w)[0]
    finally:
        engine.set_weights(orc.weights())


def test_score_host_streaming_matches_resident(engine, orc):
    C, T = 16, 70001
    dims = orc.gen_dims(123, 0, C, 0, T, 300, 8)
    r = engine.score_host(dims, 4)
    sums, counts = engine.debug_partials(C)
    esums, ecounts = orc.score_dims_fx(dims)
    assert sums == esums and counts == ecounts
    ref_s, _ = orc.score_dims(dims)
    assert np.array_equal(r.topk, orc.topk(ref_s, 4))


# ------------------------------------------------------------------ K1r, Form R per evaluation
@pytest.mark.parametrize("ap", [0, 1024, 512])
def test_score_rollouts_matches_oracle(engine, orc, ap):
    C, T = 6, 9001
    recs = orc.gen_records(0x5EED0004, orc.STREAM_ROLLOUT, 0, C, 0, T, ap, 8)
    engine.rollouts_upload(recs)
    res = engine.score(C, 3, source=1)
    ref_s, ref_n = orc.score_records(recs)
    assert_scores(res, ref_s, ref_n)
    assert np.array_equal(res.topk, orc.topk(ref_s, 3))
    sums, counts = engine.debug_partials(C)
    esums, ecounts = orc.score_records_fx(recs)
    assert sums == esums and counts == ecounts


# ------------------------------------------------------------------ K2 detect6 + report
def check_report(rep, ref):
    assert (rep.total, rep.good, rep.bad, rep.none) == (ref.total, ref.good, ref.bad, ref.none)
    assert rep.goodRate == ref.goodRate
    for m in range(5):
        assert list(rep.byMode[m]) == list(ref.byMode[m])
        assert rep.byModeGoodRate[m] == ref.byModeGoodRate[m]
    assert rep.withReward == ref.withReward
    if ref.withReward:
        assert close(rep.avgReward, ref.avgReward) and abs(rep.avgReward - ref.avgReward) < 1e-12
    else:
        assert math.isnan(rep.avgReward)
    for i in range(9):
        assert rep.dim[i].count == ref.dim[i].count
        assert close(rep.dim[i].sum, ref.dim[i].sum) and close(rep.dim[i].avg, ref.dim[i].avg)
        assert (rep.dim[i].low_flag, rep.dim[i].sugg_flag) == (ref.dim[i].low_flag, ref.dim[i].sugg_flag)
        if ref.dim[i].low_flag:
            assert rep.dim[i].low_severity == ref.dim[i].low_severity
        if ref.dim[i].sugg_flag:
            assert rep.dim[i].sugg_priority == ref.dim[i].sugg_priority
    for p in range(6):
        assert (rep.pat[p].count, rep.pat[p].flag) == (ref.pat[p].count, ref.pat[p].flag), p
        assert list(rep.pat[p].examples) == list(ref.pat[p].examples), p
        if ref.pat[p].flag:
            assert rep.pat[p].severity == ref.pat[p].severity
    assert (rep.toolCalls, rep.toolSucc, rep.toolFail) == (ref.toolCalls, ref.toolSucc, ref.toolFail)
    assert (math.isnan(rep.toolSuccessRate) and math.isnan(ref.toolSuccessRate)) or rep.toolSuccessRate == ref.toolSuccessRate


@pytest.mark.parametrize("T,base", [(1000, 0), (1, 5), (100003, 1 << 33), (37, 0)])
def test_corpus_report_matches_oracle(engine, orc, T, base):
    recs = orc.gen_records(0x5EED0003, orc.STREAM_CORPUS, 0, 1, base, T, 300, 8).reshape(-1)
    engine.corpus_upload(recs, idx_base=base)
    dims = orc.gen_dims(1, 0, 2, 0, 64, 300, 2)
    engine.dims_upload(dims)
    res = engine.score(2, 1, corpus=True)
    check_report(res.report, orc.report(recs, idx_base=base))
    assert np.array_equal(res.topk, orc.topk(orc.score_dims(dims)[0], 1))


def test_corpus_without_bad_feedback_has_no_patterns(engine, orc):
    recs = orc.gen_records(3, orc.STREAM_CORPUS, 0, 1, 0, 5000, 300, 8).reshape(-1).copy()
    recs["feedback"][recs["feedback"] == 2] = 0              # APO:641 early-out
    engine.corpus_upload(recs)
    engine.dims_upload(orc.gen_dims(1, 0, 1, 0, 8, 300, 1))
    res = engine.score(1, 1, corpus=True)
    ref = orc.report(recs)
    check_report(res.report, ref)
    assert all(res.report.pat[p].count == 0 and res.report.pat[p].flag == 0 for p in range(6))


def test_corpus_golden_report(engine, orc):
    with open(os.path.join(GOLD, "reward_cases.json")) as f:
        g = json.load(f)
    recs = np.frombuffer(b"".join(bytes.fromhex(c["record_hex"]) for c in g["random"]), orc.RECORD_DTYPE)
    engine.corpus_upload(recs)
    engine.dims_upload(orc.gen_dims(1, 0, 1, 0, 8, 300, 1))
    res = engine.score(1, 1, corpus=True)
    check_report(res.report, orc.report(recs))


# ------------------------------------------------------------------ top-K at larger C
@pytest.mark.parametrize("C,K", [(1024, 256), (1000, 1000), (300, 1), (4096, 64)])
def test_radix_topk_many_candidates(engine, orc, C, K):
    rng = np.random.default_rng(C)
    dims = np.full((C, 8, 9), np.nan, np.float32)
    vals = rng.integers(-50, 50, C).astype(np.float32) / 64.0     # many exact ties
    dims[:, :, 0] = vals[:, None]
    dims[::17] = np.nan                                           # -inf candidates
    engine.dims_upload(dims)
    res = engine.score(C, K)
    ref_s, ref_n = orc.score_dims(dims)
    assert_scores(res, ref_s, ref_n)
    assert np.array_equal(res.topk, orc.topk(ref_s, K))


def test_error_behaviour(engine, apo):
    dims = np.zeros((2, 16, 9), np.float32)
    engine.dims_upload(dims)
    with pytest.raises(apo.ApoError):
        engine.score(2, 1, first=2)                       # window start not a multiple of 4
    with pytest.raises(apo.ApoError):
        engine.score(2, 3)                                # K > C
    res = engine.score(2, 2)                              # the handle stays usable after an error
    assert list(res.topk) == [0, 1]


def test_chunked_scoring_matches_single_call(engine, orc):
    """begin / accumulate per candidate chunk and record window / finish == one apo_score."""
    C, T, K, seed = 10, 30_000, 4, 0x5EED0005
    dims = orc.gen_dims(seed, 0, C, 0, T, 300, 8)
    engine.dims_upload(dims)
    engine.corpus_upload(orc.gen_records(seed, orc.STREAM_CORPUS, 0, 1, 0, T, 300, 8).reshape(-1))
    ref = engine.score(C, K, corpus=True)
    rsums, rcounts = engine.debug_partials(C)
    engine.score_begin(C)
    for c0, cn in [(0, 4), (4, 4), (8, 2)]:                     # candidate chunks of <= 4
        engine.dims_upload(dims[c0:c0 + cn])
        for first, count in [(0, 10_000), (10_000, 20_000)]:     # two record windows each
            engine.score_accumulate(c0, first=first, count=count)
    res = engine.score_finish(C, K, corpus=True)
    sums, counts = engine.debug_partials(C)
    assert sums == rsums and counts == rcounts
    assert np.array_equal(res.scores, ref.scores) and np.array_equal(res.topk, ref.topk)
    assert res.report.bad == ref.report.bad and res.report.avgReward == ref.report.avgReward
    assert res.timing.launches == 6 + 1
    engine.score_begin(C)
    engine.dims_upload(dims[0:4])
    with pytest.raises(Exception):
        engine.score_accumulate(8)                                # [8,12) exceeds C_total = 10


def test_branch_free_divisions_are_correctly_rounded(engine, orc):
    """d2 = succ/total*2-1 uses a Newton/Markstein quotient and d5 compares dur > thr*total instead
    of dividing (csrc/apo_device.cuh); both must equal the oracle's IEEE divisions bit for bit."""
    rng = np.random.default_rng(11)
    n = 400_000
    recs = np.zeros(n, orc.RECORD_DTYPE)
    total = np.concatenate([rng.integers(1, 50, n // 4), rng.integers(1, 1 << 16, n // 4),
                            rng.integers(1, 1 << 32, n // 4, dtype=np.uint64), (1 << 32) - rng.integers(1, 1000, n // 4)]).astype(np.uint64)
    succ = (rng.random(n) * (total + 1)).astype(np.uint64).clip(0, total)
    recs["toolCalls"], recs["toolSucc"] = total.astype(np.uint32), succ.astype(np.uint32)
    recs["toolFail"] = (total - succ).astype(np.uint32)
    # durations straddling the three thresholds of TCS:725-727, to the fp32 ulp
    thr = rng.choice([1000.0, 3000.0, 10000.0], n)
    base = (thr * total.astype(np.float64)).astype(np.float32)
    step = rng.integers(-2, 3, n)
    dur = base.copy()
    for k in (1, 2):
        dur = np.where(step >= k, np.nextafter(dur, np.float32(np.inf)), dur)
        dur = np.where(step <= -k, np.nextafter(dur, np.float32(0)), dur)
    recs["toolDurMs"] = dur
    recs["flags"] = 0x0A
    dims, masks, finals = engine.reward_batch(recs)
    ref = np.empty((n, 9)); refm = np.empty(n, np.uint32)
    for i in range(0, n, 1):
        if i % 40 and i > 2000:          # the oracle loop is Python-slow: full check on 2000, then every 40th
            continue
        d, m = orc.reward_dims(recs[i])
        assert m == masks[i]
        assert d[2] == dims[i, 2], (i, int(succ[i]), int(total[i]))
        assert (np.isnan(d[5]) and np.isnan(dims[i, 5])) or d[5] == dims[i, 5], (i, float(dur[i]), int(total[i]))
    # and the whole batch through the exact-sum path (C loop, every record)
    engine.rollouts_upload(recs.reshape(1, n))
    engine.score(1, 1, source=1)
    sums, counts = engine.debug_partials(1)
    esums, ecounts = orc.score_records_fx(recs.reshape(1, n))
    assert sums == esums and counts == ecounts


# ------------------------------------------------------------------ edge cases of the boundary
def test_degenerate_shapes(engine, orc):
    # one candidate, one record, K = 0 / K = C
    dims = orc.gen_dims(2, 0, 1, 0, 1, 300, 1)
    engine.dims_upload(dims)
    r = engine.score(1, 0)
    assert r.topk.shape == (0,) and r.counts[0] in (0, 1)
    r = engine.score(1, 1)
    assert list(r.topk) == [0]
    # empty window
    engine.dims_upload(orc.gen_dims(2, 0, 3, 0, 64, 300, 1))
    r = engine.score(3, 3, first=64, count=0)            # count 0 at the end = empty remainder
    assert list(r.counts) == [0, 0, 0] and all(np.isneginf(r.scores)) and list(r.topk) == [0, 1, 2]
    # empty corpus: report stays zeroed, scoring unaffected
    engine.corpus_upload(np.empty(0, orc.RECORD_DTYPE))
    r = engine.score(3, 1, corpus=True)
    assert r.report.total == 0 and r.report.pat[0].count == 0


def test_special_values(engine, orc):
    """Signed zeros, subnormals, large magnitudes inside the documented domain, infinities outside it."""
    C, T = 2, 256
    dims = np.zeros((C, T, 9), np.float32)
    dims[0, :, 0] = -0.0
    dims[0, :, 1] = np.float32(1e-45)                     # subnormal fp32
    dims[0, :, 2] = np.nan
    dims[1, :, :] = np.linspace(-500, 500, T * 9, dtype=np.float32).reshape(T, 9)
    engine.dims_upload(dims)
    r = engine.score(C, C)
    ref_s, ref_n = orc.score_dims(dims)
    assert_scores(r, ref_s, ref_n)
    sums, counts = engine.debug_partials(C)
    assert (sums, counts) == orc.score_dims_fx(dims)


def test_zero_weight_dimensions(engine, orc):
    w = np.array([0.5, 0.0, 0.25, 0.0, 0.0, 0.25, 0.0, 0.0, 0.0])
    rng = np.random.default_rng(3)
    dims = rng.uniform(-1, 1, (3, 2000, 9)).astype(np.float32)
    dims[rng.random(dims.shape) < 0.4] = np.nan
    try:
        engine.set_weights(w)
        engine.dims_upload(dims)
        r = engine.score(3, 3)
        # rows whose present dims all carry weight 0 have totalWeight == 0 -> finalReward null (TCS:784)
        ref_s, ref_n = orc.score_dims(dims, w=w)
        assert np.array_equal(r.counts, ref_n)
        assert_scores(r, ref_s, ref_n)
        assert np.array_equal(r.topk, orc.topk(ref_s, 3))
    finally:
        engine.set_weights(orc.weights())


def test_attach_caller_owned_device_buffer(engine, orc):
    import torch
    C, T, pitch = 3, 1000, 1024
    dims = orc.gen_dims(8, 0, C, 0, T, 300, 2)
    buf = torch.full((C, pitch, 9), float("nan"), dtype=torch.float32, device="cuda")
    buf[:, :T] = torch.from_numpy(dims).cuda()
    torch.cuda.synchronize()
    engine.dims_attach(buf.data_ptr(), C, T, pitch)
    r = engine.score(C, 2)
    sums, counts = engine.debug_partials(C)
    assert (sums, counts) == orc.score_dims_fx(dims)
    stream = torch.cuda.Stream()
    engine.set_stream(stream.cuda_stream)                  # caller-provided stream
    try:
        r2 = engine.score(C, 2)
    finally:
        engine.set_stream(0)
    assert np.array_equal(r.scores, r2.scores) and np.array_equal(r.topk, r2.topk)
    engine.dims_upload(dims)                               # back to engine-owned storage


def test_random_shapes_property(engine, orc):
    """Randomised shapes / windows / variants / K: exact sums, scores, top-K vs the oracle."""
    rng = np.random.default_rng(2026)
    for trial in range(40):
        C = int(rng.integers(1, 40))
        T = int(rng.choice([1, 2, 3, 5, 127, 128, 129, 2559, 2560, 2561, 5000, 12345]))
        K = int(rng.integers(0, C + 1))
        variant = int(rng.integers(0, 5))
        if rng.random() < 0.5:
            dims = orc.gen_dims(int(rng.integers(1, 1 << 30)), int(rng.integers(0, 100)), C, int(rng.integers(0, 1 << 20)), T, int(rng.integers(0, 1025)), 4)
        else:
            dims = rng.uniform(-1, 1, (C, T, 9)).astype(np.float32)
            dims[rng.random(dims.shape) < rng.random()] = np.nan
        engine.dims_upload(dims)
        first = int(rng.integers(0, T // 4 + 1)) * 4 if T >= 4 and rng.random() < 0.5 else 0
        count = int(rng.integers(0, T - first + 1)) if rng.random() < 0.5 else 0
        r = engine.score(C, K, variant=variant, first=first, count=count)
        sub = dims[:, first:first + count] if count else dims[:, first:]
        ref_s, ref_n = orc.score_dims(sub) if sub.shape[1] else (np.full(C, -np.inf), np.zeros(C, np.uint64))
        assert_scores(r, ref_s, ref_n)
        assert np.array_equal(r.topk, orc.topk(ref_s, K)), (trial, C, T, K)
        if sub.shape[1]:
            sums, counts = engine.debug_partials(C)
            assert (sums, counts) == orc.score_dims_fx(sub)


# ------------------------------------------------------------------ Form R16 (packed) and host-record streaming
def test_rollouts16_matches_oracle(engine, orc, apo):
    C, T, seed = 5, 12_345, 0x5EED0006
    recs = orc.gen_records(seed, orc.STREAM_ROLLOUT, 2, C, 64, T, 400, 8)
    engine.rollouts16_generate(seed, 2, C, 64, T, 400)
    packed = apo.pack16(recs)
    for c in range(C):
        assert engine.rollouts16_download(c, 0, T).tobytes() == packed[c].tobytes()
    res = engine.score(C, 2, source=1)
    unpacked = orc.unpack16(packed)
    ref_s, ref_n = orc.score_records(unpacked)
    assert_scores(res, ref_s, ref_n)
    sums, counts = engine.debug_partials(C)
    assert (sums, counts) == orc.score_records_fx(unpacked) == orc.score_records_fx(recs)
    engine.rollouts16_upload(packed)                               # uploaded == generated
    engine.score(C, 2, source=1)
    assert engine.debug_partials(C) == (sums, counts)


@pytest.mark.parametrize("row", [32, 16])
def test_score_host_records_streaming(engine, orc, apo, row):
    C, T = 12, 90_001
    recs = orc.gen_records(0x5EED0007, orc.STREAM_ROLLOUT, 0, C, 0, T, 300, 8)
    host = recs if row == 32 else apo.pack16(recs)
    r = engine.score_host_records(host, 3)
    sums, counts = engine.debug_partials(C)
    assert (sums, counts) == orc.score_records_fx(recs)
    assert np.array_equal(r.topk, orc.topk(orc.score_records(recs)[0], 3))


def test_rollouts_custom_weights_use_rebuilt_tables(engine, orc):
    """K1r reads value*weight products from a table built per weight vector (apo_abi.cu build_luts)."""
    w = np.array([0.2, 0.15, 0.1, 0.05, 0.07, 0.03, 0.2, 0.1, 0.1])
    recs = orc.gen_records(0x5EED0008, orc.STREAM_ROLLOUT, 0, 4, 0, 20_000, 512, 8)
    try:
        engine.set_weights(w)
        engine.rollouts_upload(recs)
        r = engine.score(4, 2, source=1)
        sums, counts = engine.debug_partials(4)
        assert (sums, counts) == orc.score_records_fx(recs, w=w)
        assert np.array_equal(r.topk, orc.topk(orc.score_records(recs, w=w)[0], 2))
    finally:
        engine.set_weights(orc.weights())


def test_incremental_scoring_session(engine, orc):
    """One session, records arriving in three batches: after each batch finish() must equal a from-scratch
    score of everything seen so far (exact integers), and the corpus report must not double count."""
    C, T, seed = 6, 24_000, 0x5EED0009
    dims = orc.gen_dims(seed, 0, C, 0, T, 300, 8)
    recs = orc.gen_records(seed, orc.STREAM_CORPUS, 0, 1, 0, T, 300, 8).reshape(-1)
    engine.dims_upload(dims)
    engine.score_begin(C)
    for lo, hi in [(0, 8_000), (8_000, 16_000), (16_000, 24_000)]:
        engine.score_accumulate(0, first=lo, count=hi - lo)
        engine.corpus_upload(recs[:hi])
        r = engine.score_finish(C, 3, corpus=True)
        assert engine.debug_partials(C) == orc.score_dims_fx(dims[:, :hi])
        ref = orc.report(recs[:hi])
        assert (r.report.total, r.report.bad, r.report.withReward) == (ref.total, ref.bad, ref.withReward)
        assert [r.report.pat[p].count for p in range(6)] == [ref.pat[p].count for p in range(6)]
        assert np.array_equal(r.topk, orc.topk(orc.score_dims(dims[:, :hi])[0], 3))


def test_fused_and_standalone_corpus_paths_agree(engine, orc, apo):
    """The corpus scan rides inside the scoring kernel when it can hide behind it (large C*T), otherwise it is
    the stand-alone K2 launch; both must produce the identical report, for Form D, Form Q and Form R."""
    seed, C, T, Tc = 0x5EED000B, 24, 300_000, 40_000          # 24 x 300k x 36 B = 0.26 GB: fused (scan of 40k records hides)
    recs = orc.gen_records(seed, orc.STREAM_CORPUS, 0, 1, 0, Tc, 300, 8).reshape(-1)
    ref = orc.report(recs)
    engine.corpus_upload(recs)
    for layout in ("fp32", "compact", "records"):
        if layout == "records":
            engine.rollouts_generate(seed, 0, C, 0, T, 300)
            src = 1
        else:
            (engine.dims_generate if layout == "fp32" else engine.dims_generate_compact)(seed, 0, C, 0, T, 300)
            src = 0
        a = engine.score(C, 4, source=src, corpus=True)
        assert a.timing.launches == 1, layout                   # K1 + K2 + K3 in one launch
        engine.set_tuning(apo.TUNE_NO_FUSE)
        b = engine.score(C, 4, source=src, corpus=True)
        engine.set_tuning(0)
        assert b.timing.launches == 2
        check_report(a.report, ref)
        check_report(b.report, ref)
        assert np.array_equal(a.scores, b.scores) and np.array_equal(a.topk, b.topk)
        c = engine.score(C, 4, source=src)                      # no corpus request: still one launch (empty scan + last-CTA finalise)
        assert c.timing.launches == 1 and np.array_equal(c.scores, a.scores) and np.array_equal(c.topk, a.topk)


def test_plain_c_client_runs(engine, apo, tmp_path):
    """examples/score_demo.c (pure C11 against the .so) on the GPU: top-K sorted, error convention honoured."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "score_demo")
    lib_dir = os.path.dirname(apo.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "score_demo.c"),
                           "-L", lib_dir, "-lapo_b200", f"-Wl,-rpath,{lib_dir}", "-lm", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "layout=2" in r.stdout and "K > C -> rc=-1" in r.stdout


def test_scores_are_closer_to_the_exact_mean_than_the_sequential_sum(engine, orc):
    """the fixed-point accumulation is at least as accurate as the reference's own
    sequential binary64 sum.  Exact rational mean of the per-evaluation finalRewards as the yardstick."""
    from fractions import Fraction
    C, T = 2, 120_000
    dims = orc.gen_dims(0x5EED000C, 0, C, 0, T, 300, 8)
    engine.dims_upload(dims)
    res = engine.score(C, 1)
    seq, counts = orc.score_dims(dims)
    w = orc.weights()
    import ctypes
    for c in range(C):
        exact = Fraction(0)
        out = ctypes.c_double()
        for t in range(T):
            row = np.ascontiguousarray(dims[c, t])
            if orc.lib().orc_final_reward_f32(row.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p), ctypes.addressof(out)):
                exact += Fraction(out.value)
        exact /= int(counts[c])
        err_gpu = abs(Fraction(float(res.scores[c])) - exact)
        err_seq = abs(Fraction(float(seq[c])) - exact)
        assert err_gpu <= err_seq + Fraction(1, 2**60), (float(err_gpu), float(err_seq))
        assert err_gpu < Fraction(1, 2**50)


def test_rollout_records_at_every_threshold_boundary(engine, orc):
    """K1r reads four dimensions from [mode][counter] tables, tokens from a (tokens-1)/1000 bucket table and the
    tool_success_rate quotient from tabulated reciprocals when a warp's records all have < 64 tool calls: sweep every
    counter across its thresholds +-1 (both threshold sets), the 63/64 switch, saturated and degenerate records."""
    rng = np.random.default_rng(99)
    C, T = 4, 24_000
    tokens = np.array([0, 1, 999, 1000, 1001, 1999, 2000, 2001, 4999, 5000, 5001, 9999, 10000, 10001, 14999, 15000, 15001, 29999, 30000, 30001,
                       31000, 31001, 65535, 10**6, 2**32 - 1], np.uint32)
    calls = np.array([0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 11, 15, 16, 25, 26, 27, 62, 63, 64, 65, 100, 1000, 65535, 2**32 - 1], np.uint32)
    recs = np.zeros((C, T), orc.RECORD_DTYPE)
    recs["feedback"] = rng.integers(0, 3, (C, T))
    recs["flags"] = rng.integers(0, 4, (C, T)) | 8 | (rng.integers(0, 2, (C, T)) << 4)
    recs["flags"][rng.random((C, T)) < 0.03] &= ~np.uint8(8)                        # a few never-scored records
    recs["mode"] = rng.integers(0, 5, (C, T))
    recs["userMsgs"] = rng.integers(0, 12, (C, T))
    recs["asstMsgs"] = rng.integers(0, 12, (C, T))
    recs["toolCalls"] = calls[rng.integers(0, len(calls), (C, T))]
    recs["toolCalls"][0, :8000] = rng.integers(0, 64, 8000)                         # whole warps below 64: the tabulated-reciprocal path
    fail = (rng.random((C, T)) * (np.minimum(recs["toolCalls"], 8) + 1)).astype(np.uint32)
    recs["toolFail"] = np.minimum(fail, recs["toolCalls"])
    recs["toolSucc"] = recs["toolCalls"] - recs["toolFail"]
    odd = rng.random((C, T)) < 0.02                                                 # malformed: succ + fail != total
    recs["toolSucc"][odd] = rng.integers(0, 200, odd.sum())
    recs["llmCalls"] = rng.integers(0, 12, (C, T))
    recs["tokens"] = tokens[rng.integers(0, len(tokens), (C, T))]
    per = np.array([0, 1, 999.5, 1000, 1000.0001, 2999, 3000, 3001, 9999, 10000, 10001, 1e7], np.float64)
    recs["toolDurMs"] = (per[rng.integers(0, len(per), (C, T))] * np.maximum(recs["toolCalls"], 1)).astype(np.float32)
    engine.rollouts_upload(recs)
    engine.score(C, 2, source=1)
    assert engine.debug_partials(C) == orc.score_records_fx(recs)
    # the same records through the single-trace path (reward_dims) agree with the oracle per record
    flat = recs[1, :4000]
    dims, masks, finals = engine.reward_batch(flat)
    for i in range(0, 4000, 7):
        d, m = orc.reward_dims(flat[i])
        assert m == masks[i] and np.array_equal(np.nan_to_num(d, nan=9.0), np.nan_to_num(dims[i], nan=9.0))


def test_host_streaming_pageable_and_pinned_inputs_agree(engine, orc, apo):
    """apo_score_host reads page-locked memory (apo_host_alloc) in place and gathers pageable memory through its pinned
    staging buffers on host threads; both must give the resident result (several chunks, ragged last chunk, C = 1)."""
    for C, T in ((5, 4_700_001), (1, 300_007)):          # 846 MB = 4 windows of 256 MB: staging buffers are reused
        dims = orc.gen_dims(0x5EED0077, 1, C, 0, T, 300, 8)
        exp = orc.score_dims_fx(dims)
        engine.score_host(dims, 1)                                        # pageable numpy
        assert engine.debug_partials(C) == exp
        pinned = apo.host_empty(dims.shape, np.float32)
        pinned[:] = dims
        engine.score_host(pinned, 1)
        assert engine.debug_partials(C) == exp
        del pinned
    recs = orc.gen_records(0x5EED0078, orc.STREAM_ROLLOUT, 0, 3, 0, 700_001, 300, 8)
    engine.score_host_records(recs, 2)                                    # 32-byte rows, pageable
    assert engine.debug_partials(3) == orc.score_records_fx(recs)


def test_large_pageable_uploads_go_through_staging(engine, orc, apo):
    """apo_dims_upload / apo_rollouts_upload / apo_corpus_upload of >= 64 MB pageable arrays are gathered through the pinned
    staging buffers in column slices (three slices for the 605 MB case); the resident bytes must be exactly the input."""
    C, T = 4, 4_200_001
    dims = orc.gen_dims(0x5EED0079, 0, C, 0, T, 300, 8)
    engine.dims_upload(dims)
    for c, (a, n) in ((0, (0, 5000)), (3, (T - 7001, 7001)), (2, (2_000_000, 4096))):
        assert engine.dims_download(c, a, n).tobytes() == dims[c, a:a + n].tobytes()
    engine.score(C, 1)
    staged = engine.debug_partials(C)
    engine.set_tuning(apo.TUNE_NO_STAGING)
    engine.dims_upload(dims)
    engine.score(C, 1)
    engine.set_tuning(0)
    assert engine.debug_partials(C) == staged == orc.score_dims_fx(dims)
    recs = orc.gen_records(0x5EED007A, orc.STREAM_ROLLOUT, 0, 3, 0, 700_001, 300, 8)          # 67 MB
    engine.rollouts_upload(recs)
    assert engine.rollouts_download(2, 699_000, 1001).tobytes() == recs[2, 699_000:].tobytes()
    corpus = orc.gen_records(0x5EED007B, orc.STREAM_CORPUS, 0, 1, 0, 2_100_003, 300, 8).reshape(-1)   # 67 MB
    engine.corpus_upload(corpus)
    assert engine.corpus_download(2_100_000, 3).tobytes() == corpus[2_100_000:].tobytes()
    assert engine.corpus_download(0, 4096).tobytes() == corpus[:4096].tobytes()
