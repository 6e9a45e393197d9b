"""Record-axis sharding over >= 2 B200s through the C ABI's own NCCL join (-m gpu; skipped on
a 1-GPU box).  Every rank must return the same scores / top-K / report as the unsharded
single-GPU call, and the joined integer partials must be bit-identical to it."""
import os
import socket
import sys
from importlib import import_module

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C, T, K, SEED = 12, 400_003, 5, 0x5EED0005


def worker(rank, world, uid_path, out_dir, layout):
    sys.path.insert(0, ROOT)
    pkg = import_module("senweaver-ide_b200")
    torch.cuda.set_device(rank)
    eng = pkg.Engine(rank)
    uid = open(uid_path, "rb").read()
    eng.comm_init(world, rank, uid)
    if layout.endswith("nccl-join"):
        eng.set_tuning(pkg.TUNE_NCCL_JOIN)          # the fallback join: ncclAllReduce + k_finalize
    mode = eng.comm_join_mode()
    first, last = pkg.sharding.shard_range(T, world, rank)
    (eng.dims_generate_compact if layout in ("compact", "tuples") else eng.dims_generate)(SEED, 0, C, first, last - first, 300)
    eng.corpus_generate(SEED, first, last - first, 300)
    if layout == "tuples":                          # Form T: every rank its own dictionary over its shard; the joined integers are the same
        n = last - first
        book, d2book = eng.dims_codebook(), eng.dims_d2book()
        pc, pd = np.empty((C, n), np.uint32), np.empty((C, n), np.uint16)
        for c in range(C):
            eng.dims_packed_download(c, 0, n, out=(pc[c], pd[c]))
        tl, th, tbook = pkg.tuple_encode_host(pc, pd, nthreads=2)
        eng.tuples_upload(tl, th, tbook, book, d2book)
        streamed = eng.score_host_tuples(tl, th, tbook, book, d2book, K, corpus=True)       # a joined host-streaming call
        for _ in range(3):
            res = eng.score(C, K, source=pkg.SRC_TUPLES, corpus=True)
        assert np.array_equal(streamed.scores, res.scores) and np.array_equal(streamed.topk, res.topk)
    else:
        for _ in range(3):                          # consecutive joined calls alternate between the two export slots
            res = eng.score(C, K, corpus=True)
    sums, counts = eng.debug_partials(C)            # after the join: the summed integers
    rep = res.report
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), scores=res.scores, counts=res.counts, topk=res.topk,
             sums=np.array([str(s) for s in sums]), pat=np.array([[rep.pat[p].count, *rep.pat[p].examples] for p in range(6)]),
             tallies=np.array([rep.total, rep.good, rep.bad, rep.none, rep.withReward]), avg=np.array([rep.avgReward]),
             launches=np.array([res.timing.launches]), mode=np.array([mode]))
    eng.close()


@pytest.mark.parametrize("layout", ["fp32", "compact", "fp32-nccl-join", "tuples"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_score_equals_single_gpu(tmp_path, engine, orc, world, layout):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    pkg = import_module("senweaver-ide_b200")
    uid_path = tmp_path / "uid.bin"
    uid_path.write_bytes(pkg.Engine.comm_unique_id())
    mp.spawn(worker, args=(world, str(uid_path), str(tmp_path), layout), nprocs=world, join=True)
    engine.dims_generate(SEED, 0, C, 0, T, 300)
    engine.corpus_generate(SEED, 0, T, 300)
    ref = engine.score(C, K, corpus=True)
    rsums, rcounts = engine.debug_partials(C)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert [int(s) for s in z["sums"]] == rsums
        assert np.array_equal(z["counts"], ref.counts) and np.array_equal(z["scores"], ref.scores)
        assert np.array_equal(z["topk"], ref.topk)
        assert list(z["tallies"]) == [ref.report.total, ref.report.good, ref.report.bad, ref.report.none, ref.report.withReward]
        assert z["avg"][0] == ref.report.avgReward
        for p in range(6):
            assert list(z["pat"][p]) == [ref.report.pat[p].count, *ref.report.pat[p].examples]
        if layout.endswith("nccl-join"):
            assert z["mode"][0] == 1 and z["launches"][0] in (3, 4)   # K1 (+ corpus scan when it can hide behind it, else K2), ncclAllReduce, K3
        elif layout == "tuples":
            assert z["mode"][0] == 2 and z["launches"][0] == 3        # k_tuple_values, K1t, K2 with the peer-memory join in its tail
        else:
            assert z["mode"][0] == 2 and z["launches"][0] in (1, 2)   # peer-memory join inside the scoring launch (or inside K2's tail)
    # and the single-GPU result itself is pinned to the oracle on a window
    d = orc.gen_dims(SEED, 3, 1, 0, 50_000, 300, 8)
    engine.score(C, 1, first=0, count=50_000)
    s, n = engine.debug_partials(C)
    es, en = orc.score_dims_fx(d)
    assert s[3] == es[0] and n[3] == en[0]


def test_single_process_two_handles_on_threads(engine, orc):
    """The Electron main process is ONE process: two handles (one per GPU) driven by two threads of the same process
    must join through the library's communicator exactly like two ranks in two processes."""
    import threading
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    pkg = import_module("senweaver-ide_b200")
    uid = pkg.Engine.comm_unique_id()
    world, out, errs = 2, {}, []

    def run(rank):
        try:
            eng = pkg.Engine(rank)
            eng.comm_init(world, rank, uid)
            first, last = pkg.sharding.shard_range(T, world, rank)
            eng.dims_generate(SEED, 0, C, first, last - first, 300)
            eng.corpus_generate(SEED, first, last - first, 300)
            res = eng.score(C, K, corpus=True)
            out[rank] = (res.scores.copy(), res.topk.copy(), eng.debug_partials(C), res.report.bad, [res.report.pat[p].count for p in range(6)])
            eng.close()
        except Exception as e:                      # surfaced below: a failing thread must not hang the other in NCCL silently
            errs.append((rank, repr(e)))

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not errs, errs
    assert all(not t.is_alive() for t in th)
    engine.dims_generate(SEED, 0, C, 0, T, 300)
    engine.corpus_generate(SEED, 0, T, 300)
    ref = engine.score(C, K, corpus=True)
    rp = engine.debug_partials(C)
    for r in range(world):
        s, tk, parts, bad, pats = out[r]
        assert np.array_equal(s, ref.scores) and np.array_equal(tk, ref.topk) and parts == rp
        assert bad == ref.report.bad and pats == [ref.report.pat[p].count for p in range(6)]


def chunked_worker(rank, world, uid_path, out_dir):
    """BASELINE configs[4] in miniature: C_total candidates in candidate-chunks that are regenerated between passes
    (score_begin / accumulate / finish), the record axis sharded over the ranks and joined once in finish."""
    sys.path.insert(0, ROOT)
    pkg = import_module("senweaver-ide_b200")
    torch.cuda.set_device(rank)
    eng = pkg.Engine(rank)
    eng.comm_init(world, rank, open(uid_path, "rb").read())
    Ct, Cc, Tg = 40, 8, 600_000
    first, last = pkg.sharding.shard_range(Tg, world, rank)
    eng.corpus_generate(SEED, first, last - first, 300)
    eng.score_begin(Ct)
    for c0 in range(0, Ct, Cc):
        eng.dims_generate(SEED, c0, Cc, first, last - first, 300)
        half = ((last - first) // 2) // 8 * 8
        eng.score_accumulate(c0, count=half)                 # and each chunk in two record windows
        eng.score_accumulate(c0, first=half)
    res = eng.score_finish(Ct, 10, corpus=True)
    res2 = eng.score_finish(Ct, 10, corpus=True)             # repeatable finish: a second join of the same partials
    sums, counts = eng.debug_partials(Ct)
    assert np.array_equal(res.scores, res2.scores) and np.array_equal(res.topk, res2.topk) and res.report.bad == res2.report.bad
    np.savez(os.path.join(out_dir, f"c{rank}.npz"), scores=res.scores, topk=res.topk, sums=np.array([str(s) for s in sums]),
             counts=np.array(counts), bad=np.array([res.report.bad]), pat=np.array([res.report.pat[p].count for p in range(6)]))
    eng.close()


@pytest.mark.parametrize("world", [2, 8])
def test_config5_shape_chunked_sharded_session(tmp_path, engine, orc, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    pkg = import_module("senweaver-ide_b200")
    uid_path = tmp_path / "uid.bin"
    uid_path.write_bytes(pkg.Engine.comm_unique_id())
    mp.spawn(chunked_worker, args=(world, str(uid_path), str(tmp_path)), nprocs=world, join=True)
    Ct, Tg = 40, 600_000
    engine.dims_generate(SEED, 0, Ct, 0, Tg, 300)
    engine.corpus_generate(SEED, 0, Tg, 300)
    ref = engine.score(Ct, 10, corpus=True)
    rsums, rcounts = engine.debug_partials(Ct)
    es, en = orc.score_generated_fx(SEED, [0, 17, 39], 0, Tg, 300, nthreads=8)    # and the single-GPU call against the oracle
    assert [rsums[c] for c in (0, 17, 39)] == es and [rcounts[c] for c in (0, 17, 39)] == en
    for r in range(world):
        z = np.load(tmp_path / f"c{r}.npz")
        assert [int(s) for s in z["sums"]] == rsums and list(z["counts"]) == rcounts
        assert np.array_equal(z["scores"], ref.scores) and np.array_equal(z["topk"], ref.topk)
        assert z["bad"][0] == ref.report.bad and list(z["pat"]) == [ref.report.pat[p].count for p in range(6)]
