"""N > 1 host logic on CPU: world_size-2 (and 3) `gloo` processes shard the record axis, build the
packed int64 partial vector (senweaver-ide_b200/sharding.py, layout of csrc/apo_device.cuh) from the
oracle's exact per-shard sums, join with ONE all_reduce(sum) — the same single collective the GPU
path issues through NCCL — and must reproduce the unsharded result bit for bit."""
import os
import socket
import sys
from importlib import import_module

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C, T, K, SEED = 6, 10_003, 3, 0x5EED0002


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def tuple_partials(pkg, orc, dims):
    """Form T on this shard, as the engine scores it: the shard's OWN dictionary, one reward per distinct evaluation, then sums of
    table entries per candidate (csrc/apo_tuple.cu) — dictionaries differ from rank to rank, the integers they sum to do not."""
    pc, pd, book, d2book = pkg.packed_encode_host(dims, nthreads=2)
    tl, th, (tpc, tpd) = pkg.tuple_encode_host(pc, pd, nthreads=2)
    n = len(tpc)
    entries = np.full((n, 9), np.nan, np.float32)
    for j, dim in enumerate([0, 1, 3, 4, 5, 6, 7, 8]):
        code = ((tpc >> np.uint32(4 * j)) & np.uint32(15)).astype(np.int64)
        entries[code != 15, dim] = book.reshape(8, 256)[j][code[code != 15]].view(np.float32)
    entries[tpd != 4095, 2] = d2book[tpd[tpd != 4095].astype(np.int64)].view(np.float32)
    fx, counted = orc.score_dims_fx(entries.reshape(n, 1, 9)) if n else ([], [])
    idx = tl.astype(np.int64) | (th.astype(np.int64) << 16)
    sums, counts = [], []
    for c in range(dims.shape[0]):
        occ = np.bincount(idx[c], minlength=n)
        sums.append(sum(int(o) * int(v) for o, v in zip(occ, fx) if o))
        counts.append(sum(int(o) * int(k) for o, k in zip(occ, counted) if o))
    return sums, counts


def local_vector(sh, orc, rank, world, layout="fp32"):
    first, last = sh.shard_range(T, world, rank)
    vec = np.zeros(sh.acc_words(C, world), np.int64)
    dims = orc.gen_dims(SEED, 0, C, first, last - first, 300, 2)
    sums, counts = tuple_partials(import_module("senweaver-ide_b200"), orc, dims) if layout == "tuples" else orc.score_dims_fx(dims)
    sh.pack_candidate_partials(vec, sums, counts)
    recs = orc.gen_records(SEED, orc.STREAM_CORPUS, 0, 1, first, last - first, 300, 2).reshape(-1)
    rep = orc.report(recs, idx_base=first)
    base = sh.ACC_PER_CAND * C
    vec[base + sh.CORP_TALLY: base + sh.CORP_TALLY + 3] = (rep.good, rep.bad, rep.none)
    for p in range(6):
        # raw per-shard match counts (the bad==0 early-out is applied after the join)
        bad = recs["feedback"] == 2
        hit = [bad & ((recs["flags"] & 1) != 0), bad & ((recs["flags"] & 0x10) != 0), bad & (recs["tokens"] > 10000),
               bad & (recs["llmCalls"] > 2), bad & (recs["userMsgs"] >= 4), bad & (recs["toolDurMs"].astype(np.float64) > 15000)][p]
        vec[base + sh.CORP_PAT + p] = int(hit.sum())
        idx = np.flatnonzero(hit)[:3] + first
        for k, v in enumerate(idx):
            vec[base + sh.CORP_EX + 18 * rank + 3 * p + k] = int(v) + 1
    vec[base + sh.CORP_NREC] = last - first
    return vec


def worker(rank, world, port, out_dir, layout="fp32"):
    sys.path.insert(0, ROOT)
    import oracle as orc
    sh = import_module("senweaver-ide_b200").sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    t = torch.from_numpy(local_vector(sh, orc, rank, world, layout))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)              # the single join
    np.save(os.path.join(out_dir, f"joined_{rank}.npy"), t.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,layout", [(2, "fp32"), (3, "fp32"), (2, "tuples")])
def test_sharded_join_equals_unsharded(tmp_path, orc, world, layout):
    sh = import_module("senweaver-ide_b200").sharding
    port = free_port()
    mp.spawn(worker, args=(world, port, str(tmp_path), layout), nprocs=world, join=True)
    joined = [np.load(tmp_path / f"joined_{r}.npy") for r in range(world)]
    for r in range(1, world):
        assert np.array_equal(joined[0], joined[r])        # every rank finalizes from identical input
    vec = joined[0]
    # unsharded reference
    dims = orc.gen_dims(SEED, 0, C, 0, T, 300, 4)
    esums, ecounts = orc.score_dims_fx(dims)
    sums, counts = sh.unpack_candidate_partials(vec, C)
    assert sums == esums and counts == ecounts             # bit-identical for any rank count
    scores = sh.scores_from_partials(sums, counts)
    ref_s, _ = orc.score_dims(dims)
    assert np.allclose(scores, ref_s, rtol=1e-5, atol=1e-12)
    assert np.array_equal(sh.topk_indices(scores, K), orc.topk(ref_s, K))
    recs = orc.gen_records(SEED, orc.STREAM_CORPUS, 0, 1, 0, T, 300, 4).reshape(-1)
    rep = orc.report(recs)
    base = sh.ACC_PER_CAND * C
    assert tuple(vec[base + sh.CORP_TALLY: base + sh.CORP_TALLY + 3]) == (rep.good, rep.bad, rep.none)
    assert vec[base + sh.CORP_NREC] == T
    ex = sh.merge_examples(vec, C, world)
    for p in range(6):
        assert vec[base + sh.CORP_PAT + p] == rep.pat[p].count
        assert ex[p] == [x for x in rep.pat[p].examples if x >= 0]


def test_shard_ranges_cover_and_align():
    sh = import_module("senweaver-ide_b200").sharding
    for Tn in (0, 1, 3, 4, 1000, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            cuts = [sh.shard_range(Tn, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == Tn
            for (a, b), (c, d) in zip(cuts, cuts[1:]):
                assert b == c and b % 8 == 0
            assert all(a <= b for a, b in cuts)


def test_limb_roundtrip():
    sh = import_module("senweaver-ide_b200").sharding
    for v in (0, 1, -1, 2**52, -(2**52), 10**24, -(10**24) + 12345, (1 << 100) - 1):
        l0, l1, l2 = sh.split_limbs(v)
        assert 0 <= l0 < 2**32 and 0 <= l1 < 2**32 and sh.join_limbs(l0, l1, l2) == v
