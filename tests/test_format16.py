"""Form R16 (16-byte packed record): the product's host pack/unpack helpers against the oracle's
independent unpacking, and the claim of include/apo_b200.h that packing leaves every output of
the path unchanged (dims, finalReward, the six pattern predicates, tallies)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def rec_strategy():
    return st.tuples(st.integers(0, 2), st.booleans(), st.booleans(), st.booleans(), st.integers(0, 4),
                     st.integers(0, 65535), st.integers(0, 65535), st.integers(0, 65535), st.integers(0, 65535),
                     st.integers(0, 2**32 - 1), st.integers(0, 2**32 - 1), st.floats(0, 1e9, width=32, allow_nan=False))


def build(orc, t):
    fb, err, ended, valid, mode, um, am, calls, fail, llm, tok, dur = t
    fail = min(fail, calls)
    r = np.zeros(1, orc.RECORD_DTYPE)
    r["feedback"], r["mode"], r["userMsgs"], r["asstMsgs"] = fb, mode, um, am
    r["flags"] = (1 if err else 0) | (2 if ended else 0) | (8 if valid else 0) | (16 if fail > 0 else 0)
    r["toolCalls"], r["toolFail"], r["toolSucc"], r["llmCalls"], r["tokens"] = calls, fail, calls - fail, llm, tok
    r["toolDurMs"] = dur if calls else 0.0
    return r


@settings(max_examples=400, deadline=None)
@given(rec_strategy())
def test_pack16_preserves_every_output_of_the_path(apo, orc, t):
    apo.build_library()
    r = build(orc, t)
    back = orc.unpack16(apo.pack16(r))               # product pack -> oracle unpack (independent restatements)
    assert np.array_equal(apo.unpack16(apo.pack16(r)), back)
    d0, m0, f0 = orc.reward_one(r[0])
    d1, m1, f1 = orc.reward_one(back[0])
    assert m0 == m1 and f0 == f1
    assert np.array_equal(np.nan_to_num(d0, nan=9.0), np.nan_to_num(d1, nan=9.0))
    ra, rb = orc.report(r), orc.report(back)
    assert [ra.pat[p].count for p in range(6)] == [rb.pat[p].count for p in range(6)]
    assert (ra.good, ra.bad, ra.none, ra.withReward) == (rb.good, rb.bad, rb.none, rb.withReward)
    assert list(ra.byMode[int(r["mode"][0])]) == list(rb.byMode[int(r["mode"][0])])


def test_pack16_rejects_unrepresentable(apo, orc):
    apo.build_library()
    r = np.zeros(3, orc.RECORD_DTYPE)
    r["toolCalls"], r["toolSucc"], r["toolFail"] = [5, 70000, 4], [4, 70000, 1], [1, 0, 1]   # 2nd too large, 3rd succ+fail != calls
    assert apo.pack16(r[:1]).shape == (1,)
    for bad in (r[1:2], r[2:3]):
        with pytest.raises(apo.ApoError):
            apo.pack16(bad)


def test_generator_records_roundtrip(apo, orc):
    apo.build_library()
    recs = orc.gen_records(9, orc.STREAM_ROLLOUT, 0, 2, 0, 5000, 500, 4)
    back = orc.unpack16(apo.pack16(recs))
    exp = recs.copy()
    exp["tokens"] = np.minimum(recs["tokens"], 65535)                # the only saturating field the generator reaches
    assert np.array_equal(back, exp)
    assert orc.score_records_fx(back) == orc.score_records_fx(recs)  # and it does not change a single evaluation


def test_compact_host_encoder_is_lossless(apo, orc):
    """apo_compact_encode_host (host-format code, no GPU): decoding the three planes with the codebook gives back the fp32 bit
    patterns of Form D, codes are dense and ordered by value, absent = 255 / +0.0 / cleared mask bit."""
    import numpy as np
    dims = orc.gen_dims(0x5EED00C8, 2, 5, 100, 20_011, 400, 4)
    q8, d2, li, book = apo.compact_encode_host(dims, nthreads=3)
    book = book.reshape(8, 256)
    dim_of = [0, 1, 3, 4, 5, 6, 7, 8]
    for j, dim in enumerate(dim_of):
        used = book[j][book[j] != 0xFFFFFFFF]
        vals = used.view(np.float32)
        assert np.all(np.diff(vals) >= 0) and len(set(used.tolist())) == len(used)           # dense, ordered by value
        code = ((q8 >> np.uint64(8 * j)) & np.uint64(255)).astype(np.int64)
        absent = np.isnan(dims[:, :, dim])
        assert np.array_equal(code == 255, absent)
        dec = book[j][np.where(absent, 0, code)]
        assert np.array_equal(dec[~absent], dims[:, :, dim][~absent].view(np.uint32))
    a2 = np.isnan(dims[:, :, 2])
    assert np.array_equal(d2[~a2].view(np.uint32), dims[:, :, 2][~a2].view(np.uint32)) and np.all(d2[a2] == 0)
    mask = np.zeros(dims.shape[:2], np.uint32)
    for i in range(9):
        mask |= (~np.isnan(dims[:, :, i])).astype(np.uint32) << i
    assert np.array_equal(li, (((mask >> 5) | (mask << 4)) & 511).astype(np.uint16))
    one = apo.compact_encode_host(dims, nthreads=1)
    assert all(np.array_equal(x, y) for x, y in zip(one, (q8, d2, li, book.reshape(-1))))      # thread count does not matter


def test_packed_host_encoder_is_lossless(apo, orc):
    """apo_packed_encode_host (host-format code, no GPU): nibble j of pc indexes codebook[j], pd indexes d2book; decoding
    returns the fp32 bit patterns of Form D; 15 / 4095 mean absent; any thread count gives the same bytes."""
    import numpy as np
    dims = orc.gen_dims(0x5EED00CA, 1, 4, 50, 30_011, 350, 4)
    pc, pd, book, d2book = apo.packed_encode_host(dims, nthreads=3)
    book = book.reshape(8, 256)
    for j, dim in enumerate([0, 1, 3, 4, 5, 6, 7, 8]):
        code = ((pc >> np.uint32(4 * j)) & np.uint32(15)).astype(np.int64)
        absent = np.isnan(dims[:, :, dim])
        assert np.array_equal(code == 15, absent)
        assert np.array_equal(book[j][np.where(absent, 0, code)][~absent], dims[:, :, dim][~absent].view(np.uint32))
    a2 = np.isnan(dims[:, :, 2])
    assert np.array_equal(pd == 4095, a2)
    assert np.array_equal(d2book[np.where(a2, 0, pd.astype(np.int64))][~a2], dims[:, :, 2][~a2].view(np.uint32))
    used = d2book[d2book != 0xFFFFFFFF].view(np.float32)
    assert np.all(np.diff(used) >= 0) and len(used) < 4095
    one = apo.packed_encode_host(dims, nthreads=1)
    assert all(np.array_equal(x, y) for x, y in zip(one, (pc, pd, book.reshape(-1), d2book)))


def test_tuple_host_encoder_is_lossless(apo, orc):
    """apo_tuple_encode_host (host-format code, no GPU): every evaluation becomes the 24-bit index (tl | th << 16) of its Form P pair
    in the dictionary; entries are distinct, ordered most frequent first (ties by key); any thread count gives the same bytes; a
    tensor with more distinct evaluations than the capacity is refused with the count reported."""
    import numpy as np
    dims = orc.gen_dims(0x5EED00CB, 2, 6, 10, 70_003, 350, 4)
    pc, pd, _, _ = apo.packed_encode_host(dims, nthreads=3)
    tl, th, (tpc, tpd) = apo.tuple_encode_host(pc, pd, nthreads=3)
    idx = tl.astype(np.int64) | (th.astype(np.int64) << 16)
    n = len(tpc)
    assert 1000 < n < pc.size and idx.max() == n - 1
    assert np.array_equal(tpc[idx], pc) and np.array_equal(tpd[idx], pd)                 # lossless
    keys = tpc.astype(np.uint64) | (tpd.astype(np.uint64) << np.uint64(32))
    assert len(np.unique(keys)) == n                                                      # distinct entries
    cnt = np.bincount(idx.ravel(), minlength=n)
    assert cnt.min() >= 1 and np.all(np.diff(cnt) <= 0)                                   # most frequent first
    ties = np.diff(cnt) == 0
    assert np.all(keys[1:][ties] > keys[:-1][ties])                                       # ties by key: deterministic
    one = apo.tuple_encode_host(pc, pd, nthreads=1)
    assert np.array_equal(one[0], tl) and np.array_equal(one[1], th) and np.array_equal(one[2][0], tpc) and np.array_equal(one[2][1], tpd)
    with pytest.raises(apo.ApoError) as ei:
        apo.tuple_encode_host(pc, pd, cap=1000)
    assert ei.value.code == -3
    reported = int(str(ei.value).split(":")[1].split()[0])
    assert 1000 < reported <= n                                                           # stops early: a lower bound of the distinct count
    # the whole range of a 16-bit low plane and a non-zero high plane
    big_pc = np.arange(70_000, dtype=np.uint32).reshape(1, -1)
    big_pd = np.full(big_pc.shape, 4095, np.uint16)
    l, h, (bp, bd) = apo.tuple_encode_host(big_pc, big_pd)
    i2 = l.astype(np.int64) | (h.astype(np.int64) << 16)
    assert len(bp) == 70_000 and h.max() == 1 and np.array_equal(bp[i2], big_pc)
    # empty record axis
    l, h, (bp, bd) = apo.tuple_encode_host(np.empty((3, 0), np.uint32), np.empty((3, 0), np.uint16))
    assert l.shape == (3, 0) and len(bp) == 0
