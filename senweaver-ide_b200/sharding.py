"""Record-axis sharding and the packed partial vector (host-side arithmetic only).

The record axis T is cut into contiguous shards, one per rank (one process per GPU).  Each
rank produces the int64 partial vector laid out in csrc/apo_device.cuh; because every entry
is an integer (fixed-point sums in 2^-52 units kept as three limbs, counts, and per-rank
example slots), ONE elementwise int64 sum — ncclAllReduce on the GPUs, gloo in the CPU
tests — joins the shards, and the joined vector is bit-identical for any rank count.
"""
from __future__ import annotations

import numpy as np

ACC_PER_CAND = 4
CORP_REWARD, CORP_DIM, CORP_TALLY, CORP_MODE, CORP_PAT, CORP_TOOL, CORP_NREC, CORP_EX, CORP_FIXED = 0, 4, 40, 43, 58, 64, 67, 68, 68
FX_BITS = 52


def acc_words(C: int, nranks: int) -> int:
    return ACC_PER_CAND * C + CORP_FIXED + 18 * nranks


def shard_range(T: int, nranks: int, rank: int, align: int = 8) -> tuple[int, int]:
    """[first, last) of rank's contiguous shard; interior boundaries are multiples of `align`
    (8 = the window granularity of the compact layout, a multiple of the 4 the fp32 layouts need)."""
    assert 0 <= rank < nranks
    def cut(r):
        if r >= nranks:
            return T
        return min(T, (T * r // nranks) // align * align)
    return cut(rank), cut(rank + 1)


def split_limbs(v: int) -> tuple[int, int, int]:
    """Signed big integer -> (l0, l1, l2) with v == l2*2^64 + l1*2^32 + l0, 0 <= l0,l1 < 2^32."""
    l0 = v & 0xFFFFFFFF
    l1 = (v >> 32) & 0xFFFFFFFF
    l2 = v >> 64
    return l0, l1, l2


def join_limbs(l0: int, l1: int, l2: int) -> int:
    return (int(l2) << 64) + (int(l1) << 32) + int(l0)


def pack_candidate_partials(vec: np.ndarray, sums_fx: list[int], counts: list[int]) -> None:
    for c, (s, n) in enumerate(zip(sums_fx, counts)):
        l0, l1, l2 = split_limbs(int(s))
        vec[ACC_PER_CAND * c: ACC_PER_CAND * c + 4] = (l0, l1, l2, int(n))


def unpack_candidate_partials(vec: np.ndarray, C: int) -> tuple[list[int], list[int]]:
    sums, counts = [], []
    for c in range(C):
        l0, l1, l2, n = (int(x) for x in vec[ACC_PER_CAND * c: ACC_PER_CAND * c + 4])
        sums.append(join_limbs(l0, l1, l2))
        counts.append(n)
    return sums, counts


def scores_from_partials(sums_fx: list[int], counts: list[int]) -> np.ndarray:
    """score[c] = (sum / 2^52) / count, -inf when count == 0 (same rounding as the K3 kernel:
    the integer sum is rounded once to binary64, scaled exactly, then divided)."""
    out = np.empty(len(sums_fx), np.float64)
    for c, (s, n) in enumerate(zip(sums_fx, counts)):
        out[c] = -np.inf if n == 0 else (float(s) / float(1 << FX_BITS)) / float(n)
    return out


def topk_indices(scores: np.ndarray, K: int) -> np.ndarray:
    """score descending, ties -> lower index."""
    order = sorted(range(len(scores)), key=lambda c: (-scores[c], c))
    return np.array(order[:K], np.int32)


def merge_examples(vec: np.ndarray, C: int, nranks: int) -> list[list[int]]:
    """First three example indices per pattern, walking ranks in shard order."""
    base = ACC_PER_CAND * C + CORP_EX
    out = []
    for p in range(6):
        got = []
        for r in range(nranks):
            for k in range(3):
                v = int(vec[base + 18 * r + 3 * p + k])
                if v > 0 and len(got) < 3:
                    got.append(v - 1)
        out.append(got)
    return out
