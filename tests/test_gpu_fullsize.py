"""BASELINE.json-size runs (-m gpu), checked through size-independent properties and
oracle spot windows — the oracle cannot sweep 2.56 G evaluations in test time.

  * config 2 (64 x 1M) and config 3 (256 x 10M, 92 GB resident) are generated on device
    by the integer-only generator the oracle restates bit for bit;
  * oracle windows: exact integer partial sums over sampled record windows;
  * additivity: partial sums of disjoint windows add up to the whole, exactly;
  * tie-break: a duplicated candidate scores identically and the lower index wins;
  * determinism: two runs (different kernel variants) give identical integer sums.
"""
from importlib import import_module

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_config(engine, orc, seed, C, T, K, windows):
    engine.dims_generate(seed, 0, C, 0, T, 300)
    engine.corpus_generate(seed, 0, T, 300)
    res = engine.score(C, K, corpus=True)
    total, tcount = engine.debug_partials(C)
    # determinism across kernel variants
    res1 = engine.score(C, K, variant=1)
    t1, c1 = engine.debug_partials(C)
    assert t1 == total and c1 == tcount and np.array_equal(res1.topk, res.topk)
    # oracle spot windows (exact)
    for first, count, cands in windows:
        engine.score(C, 1, first=first, count=count)
        s, n = engine.debug_partials(C)
        for c in cands:
            d = orc.gen_dims(seed, c, 1, first, count, 300, 8)
            es, en = orc.score_dims_fx(d)
            assert s[c] == es[0] and n[c] == en[0], (first, c)
    # additivity over a 3-way split
    a, b = (T // 3) // 4 * 4, (2 * T // 3) // 4 * 4
    acc = [0] * C
    cnt = [0] * C
    for first, count in [(0, a), (a, b - a), (b, T - b)]:
        engine.score(C, 1, first=first, count=count)
        s, n = engine.debug_partials(C)
        acc = [x + y for x, y in zip(acc, s)]
        cnt = [x + y for x, y in zip(cnt, n)]
    assert acc == total and cnt == tcount
    # scores and top-K follow from the exact sums
    sh = import_module("senweaver-ide_b200").sharding
    exp_scores = sh.scores_from_partials(total, tcount)
    assert np.array_equal(exp_scores, res.scores)
    assert np.array_equal(sh.topk_indices(exp_scores, K), res.topk)
    rep = res.report
    assert rep.total == T and rep.good + rep.bad + rep.none == T
    assert sum(int(rep.byMode[m][0]) for m in range(5)) == T
    for p in range(6):
        ex = [x for x in rep.pat[p].examples if x >= 0]
        assert ex == sorted(ex) and len(ex) == min(3, rep.pat[p].count)
    return res


def test_config2_64x1M(engine, orc):
    res = run_config(engine, orc, 0x5EED0002, 64, 1_000_000, 16,
                     [(0, 4096, [0, 63]), (500_000, 2048, [17]), (999_000, 1000, [5, 40])])
    # whole-corpus report against the oracle (1M records is seconds of CPU work)
    recs = orc.gen_records(0x5EED0002, orc.STREAM_CORPUS, 0, 1, 0, 1_000_000, 300, 8).reshape(-1)
    ref = orc.report(recs)
    rep = res.report
    assert (rep.good, rep.bad, rep.none, rep.withReward) == (ref.good, ref.bad, ref.none, ref.withReward)
    for p in range(6):
        assert rep.pat[p].count == ref.pat[p].count and rep.pat[p].flag == ref.pat[p].flag
        assert list(rep.pat[p].examples) == list(ref.pat[p].examples)
    for i in range(9):
        assert rep.dim[i].count == ref.dim[i].count
        assert abs(rep.dim[i].avg - ref.dim[i].avg) <= 1e-5 * max(abs(ref.dim[i].avg), 1e-6)
    assert abs(rep.avgReward - ref.avgReward) <= 1e-5 * max(abs(ref.avgReward), 1e-6)


def test_config3_256x10M_resident(engine, orc):
    run_config(engine, orc, 0x5EED0003, 256, 10_000_000, 64,
               [(0, 2048, [0, 255]), (5_000_000, 2048, [128]), (9_999_000, 1000, [77])])


def test_duplicate_candidates_tie_break(engine, orc):
    engine.dims_generate(0x5EED0002, 7, 1, 0, 1_000_000, 300)
    row = engine.dims_download(0, 0, 1_000_000)
    engine.dims_upload(np.stack([row, row, row]))
    res = engine.score(3, 3)
    assert res.scores[0] == res.scores[1] == res.scores[2] and list(res.topk) == [0, 1, 2]


def test_config3_compact_layout_is_bit_identical(engine, orc):
    """256 x 10M: the 35.8 GB Form Q recoding must reproduce the 92 GB Form D result exactly."""
    C, T, seed = 256, 10_000_000, 0x5EED0003
    engine.dims_generate(seed, 0, C, 0, T, 300)
    ref = engine.score(C, 64)
    exp = engine.debug_partials(C)
    engine.dims_generate_compact(seed, 0, C, 0, T, 300)
    assert engine.dims_layout() == 2
    res = engine.score(C, 64)
    assert engine.debug_partials(C) == exp
    assert np.array_equal(res.scores, ref.scores) and np.array_equal(res.topk, ref.topk)
    # lossless: a decoded slice equals the oracle's fp32 generator output bit for bit
    d = orc.gen_dims(seed, 200, 1, 9_000_000, 4096, 300, 8)[0]
    got = engine.dims_download(200, 9_000_000, 4096)
    nan = np.isnan(d)
    assert np.array_equal(nan, np.isnan(got)) and np.array_equal(d[~nan].view(np.uint32), got[~nan].view(np.uint32))


@pytest.mark.parametrize("name,agent_permille", [("all-normal", 0), ("all-agent", 1024), ("mixed", 512)])
def test_config4_128x5M_records_thresholds(engine, orc, apo, name, agent_permille):
    """BASELINE configs[3]: 128 x 5 M per-(candidate, record) trace records (Form R, 20.5 GB), dims derived on the device with
    the normal / agent threshold sets (TCS:702-704, 711-713, 734, 741-743, 756) — all-normal, all-agent and a 50/50 mix.
    Oracle spot windows (exact integers), exact additivity over a 3-way split, variant determinism, top-K from the sums."""
    C, T, K, seed = 128, 5_000_000, 32, 0x5EED0004
    engine.rollouts_generate(seed, 0, C, 0, T, agent_permille)
    res = engine.score(C, K, source=apo.SRC_ROLLOUTS)
    total, tcount = engine.debug_partials(C)
    res1 = engine.score(C, K, source=apo.SRC_ROLLOUTS, variant=1)
    assert engine.debug_partials(C) == (total, tcount) and np.array_equal(res1.topk, res.topk)
    for first, count, cands in [(0, 4096, [0, 127]), (2_500_000, 2048, [64]), (4_999_000, 1000, [5, 101])]:
        engine.score(C, 1, source=apo.SRC_ROLLOUTS, first=first, count=count)
        s, n = engine.debug_partials(C)
        for c in cands:
            recs = orc.gen_records(seed, orc.STREAM_ROLLOUT, c, 1, first, count, agent_permille, 8)
            es, en = orc.score_records_fx(recs)
            assert s[c] == es[0] and n[c] == en[0], (name, first, c)
            modes = set(int(m) for m in np.unique(recs["mode"]))
            if agent_permille == 0:
                assert 2 not in modes
            elif agent_permille == 1024:
                assert modes == {2}
            else:
                assert 2 in modes and len(modes) > 1
    a, b = (T // 3) // 4 * 4, (2 * T // 3) // 4 * 4
    acc, cnt = [0] * C, [0] * C
    for first, count in [(0, a), (a, b - a), (b, T - b)]:
        engine.score(C, 1, source=apo.SRC_ROLLOUTS, first=first, count=count)
        s, n = engine.debug_partials(C)
        acc = [x + y for x, y in zip(acc, s)]
        cnt = [x + y for x, y in zip(cnt, n)]
    assert acc == total and cnt == tcount
    exp_scores = apo.sharding.scores_from_partials(total, tcount)
    assert np.array_equal(exp_scores, res.scores)
    assert np.array_equal(apo.sharding.topk_indices(exp_scores, K), res.topk)


def test_near_tie_topk_order_is_defined_by_the_exact_sums(engine, orc, apo):
    """Two candidates whose exact means differ by less than 1e-15 (top-K is unpinned by the reference; the
    reference's own sequential binary64 mean drifts by ~1e-14 at this size, so it cannot resolve such a gap).  The engine's
    order is defined as: score = the exact rational mean rounded once to binary64 (sum -> binary64, / count), descending,
    ties -> lower index.  Candidate 1 is candidate 0 with ONE evaluation's tool_success_rate raised by one fp32 ulp."""
    from fractions import Fraction
    T = 400_000
    base = orc.gen_dims(0x5EED0C0D, 3, 1, 0, T, 300, 8)[0]
    t_star = int(np.flatnonzero(~np.isnan(base[:, 2]))[0])
    base[t_star, 2] = np.float32(1e-3)                        # a small value: one fp32 ulp is 1.2e-10, the mean moves by ~5e-17
    bumped = base.copy()
    bumped[t_star, 2] = np.nextafter(base[t_star, 2], np.float32(2.0))
    dims = np.stack([base, bumped, base])                     # 2 == 0 exactly (tie), 1 is a hair better
    engine.dims_upload(dims)
    res = engine.score(3, 3)
    sums, counts = engine.debug_partials(3)
    assert (sums, counts) == orc.score_dims_fx(dims)          # the integers are the oracle's, exactly
    assert sums[1] > sums[0] == sums[2] and counts[0] == counts[1]
    gap = Fraction(sums[1] - sums[0], 1 << 52) / counts[0]
    assert 0 < gap < Fraction(1, 10**15)
    exp = apo.sharding.scores_from_partials(sums, counts)
    assert np.array_equal(res.scores, exp)
    assert list(res.topk) == list(apo.sharding.topk_indices(exp, 3))
    if exp[1] > exp[0]:
        assert list(res.topk) == [1, 0, 2]                    # the gap survives the rounding to binary64: the better candidate wins
    else:
        assert exp[1] == exp[0] and list(res.topk) == [0, 1, 2]   # it does not: a binary64 tie, lower index first
    # the sequential binary64 oracle (the reference's arithmetic) agrees to 1e-11 but need not resolve the gap
    seq, _ = orc.score_dims(dims)
    assert np.all(np.abs(seq - exp) <= 1e-11)
