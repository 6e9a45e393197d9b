"""Form T on the CPU (no GPU): the claim the dictionary layout rests on — "one finalReward per DISTINCT evaluation, then a sum of
table entries, gives the integers of the per-evaluation path" — checked with the oracle, and the integer packing K1t uses for its
table entries (value + counted << 58, several entries added before the split) checked exhaustively at its bounds."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

VALID_SHIFT = 58          # csrc/apo_kernels.h KT_VALID_SHIFT
GROUP_MAX = 16            # csrc/apo_tuple.cu: KT_GROUP * KT_STEPS <= 16


def decode_entries(tbook, book, d2book):
    """dictionary entries (Form P pairs) -> fp32 dims [n][9], NaN = absent"""
    tpc, tpd = tbook
    book = book.reshape(8, 256)
    out = np.full((len(tpc), 9), np.nan, np.float32)
    for j, dim in enumerate([0, 1, 3, 4, 5, 6, 7, 8]):
        code = ((tpc >> np.uint32(4 * j)) & np.uint32(15)).astype(np.int64)
        present = code != 15
        out[present, dim] = book[j][code[present]].view(np.float32)
    p2 = tpd != 4095
    out[p2, 2] = d2book[tpd[p2].astype(np.int64)].view(np.float32)
    return out


@pytest.mark.parametrize("C,T,agent", [(5, 40_003, 300), (3, 9_000, 1000), (2, 1, 0)])
def test_sum_of_dictionary_entries_equals_the_per_evaluation_sums(apo, orc, C, T, agent):
    dims = orc.gen_dims(0x5EED00E1 + C, 1, C, 77, T, agent, 4)
    pc, pd, book, d2book = apo.packed_encode_host(dims, nthreads=2)
    tl, th, tbook = apo.tuple_encode_host(pc, pd, nthreads=2)
    n = len(tbook[0])
    entries = decode_entries(tbook, book, d2book)
    idx = tl.astype(np.int64) | (th.astype(np.int64) << 16)
    assert np.array_equal(np.nan_to_num(entries[idx], nan=9.0), np.nan_to_num(dims, nan=9.0))          # lossless, entry by entry
    fx, counted = orc.score_dims_fx(entries.reshape(n, 1, 9))                                        # one evaluation per "candidate"
    assert all(abs(v) < (1 << 53) for v in fx) and all(k in (0, 1) for k in counted)                  # what k_tuple_values requires
    exp_s, exp_n = orc.score_dims_fx(dims)
    for c in range(C):
        occ = np.bincount(idx[c], minlength=n)
        assert sum(int(o) * int(v) for o, v in zip(occ, fx) if o) == exp_s[c]
        assert sum(int(o) * int(k) for o, k in zip(occ, counted) if o) == exp_n[c]


def split(s):
    """K1t: value sum and count of a group of packed entries (arithmetic shift, as the kernel's int64 code)"""
    k = (s + (1 << (VALID_SHIFT - 1))) >> VALID_SHIFT
    return s - (k << VALID_SHIFT), k


@settings(max_examples=300, deadline=None)
@given(st.lists(st.tuples(st.integers(-(1 << 53) + 1, (1 << 53) - 1), st.booleans()), min_size=1, max_size=GROUP_MAX))
def test_packed_entries_split_exactly(entries):
    s = sum(v + (int(k) << VALID_SHIFT) for v, k in entries)
    assert -(1 << 63) <= s < (1 << 63)                               # fits the kernel's int64
    assert split(s) == (sum(v for v, _ in entries), sum(int(k) for _, k in entries))


def test_packed_entries_split_at_the_bounds():
    lim = (1 << 53) - 1
    for v in (lim, -lim):
        for k in (0, 1):
            for g in (1, 8, GROUP_MAX):
                s = g * (v + (k << VALID_SHIFT))
                assert -(1 << 63) <= s < (1 << 63) and split(s) == (g * v, g * k)
