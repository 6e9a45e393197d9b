"""Form Q — the compact resident layout (14 B per evaluation) and its kernel K1q (-m gpu).
Lossless: decoding returns the original fp32 bit patterns, and the integer partial sums equal
those of the fp32 kernel and of the oracle with no tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    nan = np.isnan(a)
    return np.array_equal(nan, np.isnan(b)) and np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32))


@pytest.mark.parametrize("C,T", [(4, 1000), (3, 1), (7, 5121), (16, 40_003), (2, 5120)])
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6])
def test_compact_matches_fp32_and_oracle(engine, orc, C, T, variant):
    dims = orc.gen_dims(0x5EED0010 + C, 3, C, 100, T, 400, 8)
    engine.dims_upload(dims)
    ref = engine.score(C, min(C, 3))
    rs, rc = engine.debug_partials(C)
    assert engine.dims_layout() == 1
    engine.dims_compact()
    assert engine.dims_layout() == 2
    for c in range(C):
        assert same_bits(engine.dims_download(c, 0, T), dims[c])          # lossless
    res = engine.score(C, min(C, 3), variant=variant)
    assert engine.debug_partials(C) == (rs, rc) == orc.score_dims_fx(dims)
    assert np.array_equal(res.scores, ref.scores) and np.array_equal(res.topk, ref.topk)


def test_compact_generate_and_upload_variants(engine, orc):
    C, T, seed = 9, 33_333, 0x5EED0011
    dims = orc.gen_dims(seed, 0, C, 64, T, 300, 8)
    exp = orc.score_dims_fx(dims)
    engine.dims_generate_compact(seed, 0, C, 64, T, 300)
    assert engine.dims_layout() == 2
    engine.score(C, 2)
    assert engine.debug_partials(C) == exp
    assert same_bits(engine.dims_download(4, 10, 500), dims[4, 10:510])
    engine.dims_upload_compact(dims)
    engine.score(C, 2)
    assert engine.debug_partials(C) == exp
    # windows and chunked accumulation work on the compact layout too
    engine.score(C, 1, first=1000, count=20_000)
    assert engine.debug_partials(C) == orc.score_dims_fx(dims[:, 1000:21_000])
    engine.score_begin(C + 2)
    engine.score_accumulate(2)
    r = engine.score_finish(C + 2, 3)
    assert r.counts[0] == 0 and r.counts[1] == 0 and engine.debug_partials(C + 2)[0][2:] == exp[0]


def test_compact_all_presence_masks_and_custom_weights(engine, orc):
    rng = np.random.default_rng(17)
    C, T = 5, 20_000
    levels = np.array([-1.0, -0.8, -0.5, -0.3, -0.2, 0.0, -0.0, 0.3, 0.5, 0.8, 1.0], np.float32)
    dims = levels[rng.integers(0, len(levels), (C, T, 9))]
    dims[:, :, 2] = rng.uniform(-1, 1, (C, T)).astype(np.float32)        # d2 stays free-form fp32
    dims[rng.random(dims.shape) < 0.3] = np.nan
    dims[3] = np.nan
    w = np.array([0.3, 0.1, 0.05, 0.05, 0.1, 0.1, 0.1, 0.1, 0.1])          # total weight with an all-ones significand
    try:
        wz = np.array([0.0, 0.4, 0.05, 0.0, 0.1, 0.1, 0.15, 0.1, 0.1])        # zero weights: all-null masks exist
        for weights in (orc.weights(), w, wz):
            engine.set_weights(weights)
            engine.dims_upload(dims)
            engine.dims_compact()
            for variant in (0, 4, 6):                                       # default mixed lookup / product tables only / all fp32
                r = engine.score(C, C, variant=variant)
                assert engine.debug_partials(C) == orc.score_dims_fx(dims, w=weights), variant
                assert np.array_equal(r.topk, orc.topk(orc.score_dims(dims, w=weights)[0], C))
        # changing the weights after compaction rebuilds the product tables
        engine.set_weights(orc.weights())
        engine.score(C, 1)
        assert engine.debug_partials(C) == orc.score_dims_fx(dims)
    finally:
        engine.set_weights(orc.weights())


def test_non_categorical_data_keeps_form_d(engine, orc, apo):
    rng = np.random.default_rng(1)
    dims = rng.uniform(-1, 1, (3, 4000, 9)).astype(np.float32)           # > 255 distinct values per dimension
    engine.dims_upload(dims)
    with pytest.raises(apo.ApoError):
        engine.dims_compact()
    assert engine.dims_layout() == 1
    engine.score(3, 1)
    assert engine.debug_partials(3) == orc.score_dims_fx(dims)
    with pytest.raises(apo.ApoError):
        engine.dims_upload_compact(dims)
    assert engine.dims_layout() == 0
    # exactly 255 distinct values still fit, 256 do not
    base = np.full((1, 512, 9), np.nan, np.float32)
    base[0, :255, 0] = np.arange(255, dtype=np.float32) / 256
    engine.dims_upload_compact(base)
    assert engine.dims_layout() == 2
    base[0, 255, 0] = 0.999
    with pytest.raises(apo.ApoError):
        engine.dims_upload_compact(base)


def test_compact_full_size_config2(engine, orc):
    C, T, seed = 64, 1_000_000, 0x5EED0002
    engine.dims_generate(seed, 0, C, 0, T, 300)
    ref = engine.score(C, 16)
    exp = engine.debug_partials(C)
    engine.dims_generate_compact(seed, 0, C, 0, T, 300)
    res = engine.score(C, 16)
    assert engine.debug_partials(C) == exp
    assert np.array_equal(res.scores, ref.scores) and np.array_equal(res.topk, ref.topk)


def test_mixed_lookup_falls_back_when_the_prefix_table_is_too_large(engine, orc):
    """40 x 40 distinct (d0, d1) values exceed the 1024-entry prefix table: the mixed-lookup variants fall back and stay exact."""
    rng = np.random.default_rng(23)
    C, T = 3, 30_000
    lv = np.linspace(-1, 1, 40).astype(np.float32)
    dims = np.full((C, T, 9), np.nan, np.float32)
    dims[:, :, 0] = lv[rng.integers(0, 40, (C, T))]
    dims[:, :, 1] = lv[rng.integers(0, 40, (C, T))]
    dims[:, :, 6] = lv[rng.integers(0, 5, (C, T))]
    dims[rng.random(dims.shape) < 0.1] = np.nan
    engine.dims_upload(dims)
    engine.dims_compact()
    exp = orc.score_dims_fx(dims)
    for variant in (0, 5):
        engine.score(C, 1, variant=variant)
        assert engine.debug_partials(C) == exp
    # 31 x 31 (+ absent) just fits
    dims[:, :, 0] = lv[rng.integers(0, 31, (C, T))]
    dims[:, :, 1] = lv[rng.integers(0, 31, (C, T))]
    dims[rng.random(dims.shape) < 0.1] = np.nan
    engine.dims_upload(dims)
    engine.dims_compact()
    engine.score(C, 1, variant=5)
    assert engine.debug_partials(C) == orc.score_dims_fx(dims)


def test_host_compact_wire_format_streams_to_the_same_integers(engine, orc, apo):
    """Form Q as a PCIe wire format (14 B / evaluation): apo_score_host_compact over host planes must give the integers of the
    Form D tensor; the host encoder (apo_compact_encode_host) must emit exactly the planes and the codebook of the device
    transcoder; a resident Form Q tensor with another codebook keeps working after the streaming call."""
    seed, C, T = 0x5EED00C7, 9, 150_001
    dims = orc.gen_dims(seed, 0, C, 0, T, 300, 8)
    exp = orc.score_dims_fx(dims)
    q8, d2, li, book = apo.compact_encode_host(dims, nthreads=4)
    engine.dims_upload_compact(dims)
    assert np.array_equal(engine.dims_codebook(), book)
    for c in (0, 4, 8):
        a, b, l = engine.dims_compact_download(c, 0, T)
        assert np.array_equal(a, q8[c]) and np.array_equal(b.view(np.uint32), d2[c].view(np.uint32)) and np.array_equal(l, li[c])
    res = engine.score(C, 3)
    assert engine.debug_partials(C) == exp
    # another tensor stays resident (different candidate qualities -> usually another codebook order) while we stream this one
    other = orc.gen_dims(seed + 1, 40, 3, 0, 40_000, 900, 8)
    engine.dims_upload_compact(other)
    exp_other = orc.score_dims_fx(other)
    recs = orc.gen_records(seed, orc.STREAM_CORPUS, 0, 1, 0, 30_000, 300, 8).reshape(-1)
    engine.corpus_upload(recs)
    for host in ("pageable", "pinned"):
        planes = (q8, d2, li)
        if host == "pinned":
            planes = tuple(apo.host_empty(p.shape, p.dtype) for p in (q8, d2, li))
            for dst, src in zip(planes, (q8, d2, li)):
                dst[:] = src
        r = engine.score_host_compact(*planes, book, 3, corpus=True)
        assert engine.debug_partials(C) == exp
        assert np.array_equal(r.scores, res.scores) and np.array_equal(r.topk, res.topk)
        assert r.report.bad == orc.report(recs).bad
    engine.score(3, 1)                                   # the resident tensor's lookup tables were restored
    assert engine.debug_partials(3) == exp_other
    # not categorical -> the encoder refuses (Form D must be kept)
    noisy = dims.copy()
    noisy[0, :300, 3] = np.linspace(-1, 1, 300, dtype=np.float32)
    with pytest.raises(apo.ApoError):
        apo.compact_encode_host(noisy)


def test_packed_wire_format_six_bytes_per_evaluation(engine, orc, apo):
    """Form P (6 B / evaluation): the host encoder and the device export of a resident tensor agree bit for bit, streaming it
    (apo_score_host_packed: H2D -> k_unpack_p -> K1q) gives the integers of the Form D tensor, and data that does not fit the
    4-bit / 12-bit codes is refused."""
    seed, C, T = 0x5EED00C9, 7, 120_007
    dims = orc.gen_dims(seed, 3, C, 1000, T, 400, 8)
    exp = orc.score_dims_fx(dims)
    pc, pd, book, d2book = apo.packed_encode_host(dims, nthreads=4)
    engine.dims_upload_compact(dims)
    assert np.array_equal(engine.dims_codebook(), book) and np.array_equal(engine.dims_d2book(), d2book)
    for c in (0, 3, 6):
        a, b = engine.dims_packed_download(c, 0, T)
        assert np.array_equal(a, pc[c]) and np.array_equal(b, pd[c])
    a, b = engine.dims_packed_download(2, 5000, 4097)                     # a window
    assert np.array_equal(a, pc[2, 5000:9097]) and np.array_equal(b, pd[2, 5000:9097])
    recs = orc.gen_records(seed, orc.STREAM_CORPUS, 0, 1, 0, 20_000, 300, 8).reshape(-1)
    engine.corpus_upload(recs)
    ref = engine.score(C, 3)
    r = engine.score_host_packed(pc, pd, book, d2book, 3, corpus=True)
    assert engine.debug_partials(C) == exp
    assert np.array_equal(r.scores, ref.scores) and np.array_equal(r.topk, ref.topk) and r.report.bad == orc.report(recs).bad
    assert r.timing.launches >= 3                                          # unpack + K1q per window, then the corpus scan / finalize
    ppc, ppd = apo.host_empty(pc.shape, np.uint32), apo.host_empty(pd.shape, np.uint16)
    ppc[:], ppd[:] = pc, pd
    engine.score_host_packed(ppc, ppd, book, d2book, 3)
    assert engine.debug_partials(C) == exp
    engine.score(C, 1)                                                     # the resident tensor is untouched by the streaming calls
    assert engine.debug_partials(C) == exp
    # 16 distinct values in a coded dimension: Form Q still works, Form P refuses (encoder and export)
    wide = dims.copy()
    wide[0, :16, 4] = np.arange(16, dtype=np.float32) / 32
    with pytest.raises(apo.ApoError):
        apo.packed_encode_host(wide)
    engine.dims_upload_compact(wide)
    with pytest.raises(apo.ApoError):
        engine.dims_packed_download(0, 0, 100)
    # more than 4095 distinct tool_success_rate values
    many = dims.copy()
    many[1, :5000, 2] = np.linspace(-1, 1, 5000, dtype=np.float32)
    with pytest.raises(apo.ApoError):
        apo.packed_encode_host(many)


def test_tuple_dictionary_form_three_bytes_per_evaluation(engine, orc, apo):
    """Form T (3 B / evaluation): dictionary indices + one finalReward per distinct evaluation (k_tuple_values, the operations
    of K1q) + K1t's gather-and-add give the integers of the Form D tensor — streamed from host memory and resident, with the
    dictionary tail beyond the shared-memory head, ragged rows, windows, sessions and changed weights."""
    seed, C, T = 0x5EED00D1, 8, 300_011                                   # > 24576 distinct evaluations: head + tail of the table
    dims = orc.gen_dims(seed, 3, C, 1000, T, 400, 8)
    exp = orc.score_dims_fx(dims)
    pc, pd, book, d2book = apo.packed_encode_host(dims, nthreads=4)
    tl, th, tbook = apo.tuple_encode_host(pc, pd, nthreads=4)
    assert len(tbook[0]) > 24576
    recs = orc.gen_records(seed, orc.STREAM_CORPUS, 0, 1, 0, 20_000, 300, 8).reshape(-1)
    engine.corpus_upload(recs)
    engine.dims_upload(dims)
    ref = engine.score(C, 3, corpus=True)
    assert engine.debug_partials(C) == exp
    # -- streamed from host memory (pageable, then page-locked)
    r = engine.score_host_tuples(tl, th, tbook, book, d2book, 3, corpus=True)
    assert engine.debug_partials(C) == exp
    assert np.array_equal(r.scores, ref.scores) and np.array_equal(r.counts, ref.counts) and np.array_equal(r.topk, ref.topk)
    assert r.report.bad == ref.report.bad and r.report.total == ref.report.total
    ptl, pth = apo.host_empty(tl.shape, np.uint16), apo.host_empty(th.shape, np.uint8)
    ptl[:], pth[:] = tl, th
    engine.score_host_tuples(ptl, pth, tbook, book, d2book, 3)
    assert engine.debug_partials(C) == exp
    # -- resident
    engine.tuples_upload(tl, th, tbook, book, d2book)
    r = engine.score(C, 3, source=apo.SRC_TUPLES, corpus=True)
    assert engine.debug_partials(C) == exp
    assert np.array_equal(r.scores, ref.scores) and np.array_equal(r.topk, ref.topk) and r.report.bad == ref.report.bad
    r = engine.score(C, 3, source=apo.SRC_TUPLES, recip=True)
    rr = engine.score(C, 3, recip=True)
    assert np.array_equal(r.scores, rr.scores)
    engine.score(C, 1, source=apo.SRC_TUPLES, first=8000, count=100_003)
    assert engine.debug_partials(C) == orc.score_dims_fx(dims[:, 8000:108_003])
    with pytest.raises(apo.ApoError):
        engine.score(C, 1, source=apo.SRC_TUPLES, first=4, count=100)       # windows start on multiples of 8
    engine.score_begin(C + 2)
    engine.score_accumulate(2, source=apo.SRC_TUPLES, count=150_000)
    engine.score_accumulate(2, source=apo.SRC_TUPLES, first=150_000)
    r = engine.score_finish(C + 2, 3)
    assert r.counts[0] == 0 and r.counts[1] == 0 and engine.debug_partials(C + 2)[0][2:] == exp[0]
    # -- another weight vector: the table is rebuilt per call
    w = np.array([0.3, 0.05, 0.1, 0.05, 0.1, 0.1, 0.1, 0.1, 0.1])
    engine.set_weights(w)
    engine.score(C, 2, source=apo.SRC_TUPLES)
    got = engine.debug_partials(C)
    engine.score(C, 2)
    assert got == engine.debug_partials(C) == orc.score_dims_fx(dims, w=w)
    engine.set_weights(orc.weights())
    # -- small dictionary (everything in shared memory), tiny tensors, empty record axis
    for Cs, Ts in ((3, 1), (2, 7), (5, 8192), (4, 8193), (1, 20_000)):
        d = orc.gen_dims(seed + Ts, 0, Cs, 0, Ts, 300, 8)
        p1, p2, b1, b2 = apo.packed_encode_host(d)
        t1, t2, tb = apo.tuple_encode_host(p1, p2)
        engine.score_host_tuples(t1, t2, tb, b1, b2, 1)
        assert engine.debug_partials(Cs) == orc.score_dims_fx(d)
        engine.tuples_upload(t1, t2, tb, b1, b2)
        engine.score(Cs, 1, source=apo.SRC_TUPLES)
        assert engine.debug_partials(Cs) == orc.score_dims_fx(d)
    # -- an index outside the dictionary is refused, not read
    bad = tl.copy()
    bad[2, 77] = 0xFFFF
    badh = th.copy()
    badh[2, 77] = 0xFF
    with pytest.raises(apo.ApoError):
        engine.score_host_tuples(bad, badh, tbook, book, d2book, 1)
    engine.score_host_tuples(tl, th, tbook, book, d2book, 1)               # and the handle stays usable
    assert engine.debug_partials(C) == exp
