"""ctypes/numpy front end of the CPU oracle (oracle/apo_oracle.c).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never from the product package.
Parity pinned by tests/golden/ref_*.json (reference method texts executed by oracle/ts_harness; see apo_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libapo_oracle.so")

NDIM, NPAT, NMODE = 9, 6, 5
STREAM_CORPUS, STREAM_ROLLOUT = 1, 2

RECORD_DTYPE = np.dtype([
    ("feedback", "u1"), ("flags", "u1"), ("mode", "u1"), ("durClass", "u1"),
    ("userMsgs", "<u2"), ("asstMsgs", "<u2"),
    ("toolCalls", "<u4"), ("toolSucc", "<u4"), ("toolFail", "<u4"),
    ("llmCalls", "<u4"), ("tokens", "<u4"), ("toolDurMs", "<f4"),
])
assert RECORD_DTYPE.itemsize == 32
RECORD16_DTYPE = np.dtype([
    ("hdr", "<u2"), ("userMsgs", "u1"), ("asstMsgs", "u1"), ("toolCalls", "<u2"), ("toolFail", "<u2"),
    ("llmCalls", "u1"), ("durClass", "u1"), ("tokens", "<u2"), ("toolDurMs", "<f4"),
])
assert RECORD16_DTYPE.itemsize == 16


class Pattern(C.Structure):
    _fields_ = [("count", C.c_uint64), ("flag", C.c_uint8), ("severity", C.c_uint8), ("examples", C.c_int64 * 3)]


class DimStat(C.Structure):
    _fields_ = [("sum", C.c_double), ("count", C.c_uint64), ("avg", C.c_double),
                ("low_flag", C.c_uint8), ("low_severity", C.c_uint8), ("sugg_flag", C.c_uint8), ("sugg_priority", C.c_uint8)]


class Report(C.Structure):
    _fields_ = [("total", C.c_uint64), ("good", C.c_uint64), ("bad", C.c_uint64), ("none", C.c_uint64),
                ("goodRate", C.c_double), ("byMode", (C.c_uint64 * 3) * NMODE), ("byModeGoodRate", C.c_double * NMODE),
                ("withReward", C.c_uint64), ("rewardSum", C.c_double), ("avgReward", C.c_double),
                ("dim", DimStat * NDIM), ("pat", Pattern * NPAT),
                ("toolCalls", C.c_uint64), ("toolSucc", C.c_uint64), ("toolFail", C.c_uint64),
                ("toolSuccessRate", C.c_double)]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("apo_oracle.c", "apo_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libapo_oracle.so"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u32, u64, dp, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_int
        L.orc_reward_dims.restype = u32
        L.orc_reward_dims.argtypes = [vp, dp]
        L.orc_final_reward.restype = i32
        L.orc_final_reward.argtypes = [dp, u32, dp, dp]
        L.orc_final_reward_f32.restype = i32
        L.orc_final_reward_f32.argtypes = [vp, dp, dp]
        L.orc_score_dims.argtypes = [vp, u32, u64, u64, dp, dp, vp]
        L.orc_score_records.argtypes = [vp, u32, u64, u64, dp, dp, vp]
        L.orc_score_dims_mt.argtypes = [vp, u32, u64, u64, dp, dp, vp, i32]
        L.orc_score_records_mt.argtypes = [vp, u32, u64, u64, dp, dp, vp, i32]
        L.orc_score_dims_fx.argtypes = [vp, u32, u64, u64, dp, vp, vp, vp]
        L.orc_score_records_fx.argtypes = [vp, u32, u64, u64, dp, vp, vp, vp]
        L.orc_unpack16.argtypes = [vp, u64, vp]
        L.orc_unpack16.restype = None
        L.orc_topk.argtypes = [dp, u32, u32, vp]
        L.orc_report_build.argtypes = [vp, u64, u64, dp, C.POINTER(Report)]
        L.orc_report_build_mt.argtypes = [vp, u64, u64, dp, C.POINTER(Report), i32]
        L.orc_report_generated.argtypes = [u64, u64, u64, u64, u32, dp, C.POINTER(Report), i32]
        L.orc_score_generated_fx.argtypes = [u64, vp, u32, u64, u64, u32, dp, vp, vp, vp, i32]
        L.orc_gen_record.argtypes = [u64, u32, u32, u64, u32, vp]
        L.orc_gen_dims_row.argtypes = [u64, u32, u64, u32, vp]
        L.orc_gen_dims.argtypes = [u64, u32, u32, u64, u64, u64, u32, vp, i32]
        L.orc_gen_records.argtypes = [u64, u32, u32, u32, u64, u64, u64, u32, vp, i32]
        for f in (L.orc_score_dims_fx, L.orc_score_records_fx, L.orc_score_dims, L.orc_score_records, L.orc_score_dims_mt, L.orc_score_records_mt, L.orc_topk,
                  L.orc_report_build, L.orc_report_build_mt, L.orc_report_generated, L.orc_score_generated_fx, L.orc_gen_record, L.orc_gen_dims_row, L.orc_gen_dims, L.orc_gen_records):
            f.restype = None
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def weights() -> np.ndarray:
    return np.array((C.c_double * NDIM).in_dll(lib(), "orc_weights"), dtype=np.float64)


def reward_dims(rec: np.ndarray):
    """One Form-R record -> (dims[9] float64 with NaN = absent, mask)."""
    rec = np.ascontiguousarray(rec, dtype=RECORD_DTYPE).reshape(1)
    d = np.empty(NDIM, np.float64)
    m = lib().orc_reward_dims(_p(rec), _p(d))
    return d, int(m)


def final_reward(dims: np.ndarray, mask: int, w: np.ndarray | None = None):
    w = weights() if w is None else np.ascontiguousarray(w, np.float64)
    dims = np.ascontiguousarray(dims, np.float64)
    out = C.c_double()
    ok = lib().orc_final_reward(_p(dims), mask, _p(w), C.addressof(out))
    return out.value if ok else None


def reward_one(rec: np.ndarray, w: np.ndarray | None = None):
    """(dims, mask, finalReward|None) for one record, honouring the VALID flag."""
    d, m = reward_dims(rec)
    rec = np.ascontiguousarray(rec, dtype=RECORD_DTYPE).reshape(1)
    if not (int(rec["flags"][0]) & 0x08):
        return d, m, None
    return d, m, final_reward(d, m, w)


def score_dims(dims: np.ndarray, T: int | None = None, w=None, nthreads: int = 0):
    """dims float32 [C][pitch][9] -> (scores f64[C], counts u64[C])."""
    dims = np.ascontiguousarray(dims, np.float32)
    Cn, pitch, nd = dims.shape
    assert nd == NDIM
    T = pitch if T is None else T
    w = weights() if w is None else np.ascontiguousarray(w, np.float64)
    s = np.empty(Cn, np.float64)
    n = np.empty(Cn, np.uint64)
    if nthreads > 0:
        lib().orc_score_dims_mt(_p(dims), Cn, T, pitch, _p(w), _p(s), _p(n), nthreads)
    else:
        lib().orc_score_dims(_p(dims), Cn, T, pitch, _p(w), _p(s), _p(n))
    return s, n


def score_records(recs: np.ndarray, T: int | None = None, w=None, nthreads: int = 0):
    recs = np.ascontiguousarray(recs, RECORD_DTYPE)
    Cn, pitch = recs.shape
    T = pitch if T is None else T
    w = weights() if w is None else np.ascontiguousarray(w, np.float64)
    s = np.empty(Cn, np.float64)
    n = np.empty(Cn, np.uint64)
    if nthreads > 0:
        lib().orc_score_records_mt(_p(recs), Cn, T, pitch, _p(w), _p(s), _p(n), nthreads)
    else:
        lib().orc_score_records(_p(recs), Cn, T, pitch, _p(w), _p(s), _p(n))
    return s, n


def score_dims_fx(dims: np.ndarray, T: int | None = None, w=None):
    """Exact per-candidate integer sums sum_t rint(fr*2^52) (Python ints) and counts."""
    dims = np.ascontiguousarray(dims, np.float32)
    Cn, pitch, _ = dims.shape
    T = pitch if T is None else T
    w = weights() if w is None else np.ascontiguousarray(w, np.float64)
    lo = np.empty(Cn, np.uint64); hi = np.empty(Cn, np.int64); n = np.empty(Cn, np.uint64)
    lib().orc_score_dims_fx(_p(dims), Cn, T, pitch, _p(w), _p(lo), _p(hi), _p(n))
    return [(int(h) << 64) + int(l) for l, h in zip(lo, hi)], [int(x) for x in n]


def score_records_fx(recs: np.ndarray, T: int | None = None, w=None):
    recs = np.ascontiguousarray(recs, RECORD_DTYPE)
    Cn, pitch = recs.shape
    T = pitch if T is None else T
    w = weights() if w is None else np.ascontiguousarray(w, np.float64)
    lo = np.empty(Cn, np.uint64); hi = np.empty(Cn, np.int64); n = np.empty(Cn, np.uint64)
    lib().orc_score_records_fx(_p(recs), Cn, T, pitch, _p(w), _p(lo), _p(hi), _p(n))
    return [(int(h) << 64) + int(l) for l, h in zip(lo, hi)], [int(x) for x in n]


def unpack16(recs16: np.ndarray) -> np.ndarray:
    recs16 = np.ascontiguousarray(recs16, RECORD16_DTYPE)
    out = np.empty(recs16.shape, RECORD_DTYPE)
    lib().orc_unpack16(_p(recs16), recs16.size, _p(out))
    return out


def topk(scores: np.ndarray, K: int) -> np.ndarray:
    scores = np.ascontiguousarray(scores, np.float64)
    out = np.empty(K, np.int32)
    lib().orc_topk(_p(scores), scores.shape[0], K, _p(out))
    return out


def report(recs: np.ndarray, idx_base: int = 0, w=None, nthreads: int = 0) -> Report:
    """nthreads == 0: single thread in exact reference order (the parity oracle); > 0: per-slice partials on the
    thread pool (same integers and examples, binary64 sums merged per slice)."""
    recs = np.ascontiguousarray(recs, RECORD_DTYPE).reshape(-1)
    w = weights() if w is None else np.ascontiguousarray(w, np.float64)
    r = Report()
    if nthreads > 0:
        lib().orc_report_build_mt(_p(recs), recs.shape[0], idx_base, _p(w), C.byref(r), nthreads)
    else:
        lib().orc_report_build(_p(recs), recs.shape[0], idx_base, _p(w), C.byref(r))
    return r


def report_generated(seed: int, t0: int, T: int, idx_base: int | None = None, agent_permille: int = 300, w=None,
                     nthreads: int = 8) -> Report:
    """Report over records [t0, t0+T) of the generator's corpus stream, generated on the fly (never materialised)."""
    w = weights() if w is None else np.ascontiguousarray(w, np.float64)
    r = Report()
    lib().orc_report_generated(seed, t0, T, t0 if idx_base is None else idx_base, agent_permille, _p(w), C.byref(r), nthreads)
    return r


def score_generated_fx(seed: int, cands, t0: int, T: int, agent_permille: int = 300, w=None, nthreads: int = 8):
    """Exact integer sums / counts (as score_dims_fx) of the listed candidates over generated Form D evaluations."""
    cands = np.ascontiguousarray(cands, np.uint32)
    w = weights() if w is None else np.ascontiguousarray(w, np.float64)
    n = cands.shape[0]
    lo = np.empty(n, np.uint64); hi = np.empty(n, np.int64); cnt = np.empty(n, np.uint64)
    lib().orc_score_generated_fx(seed, _p(cands), n, t0, T, agent_permille, _p(w), _p(lo), _p(hi), _p(cnt), nthreads)
    return [(int(h) << 64) + int(l) for l, h in zip(lo, hi)], [int(x) for x in cnt]


def gen_records(seed: int, stream: int, c0: int, Cn: int, t0: int, T: int, agent_permille: int = 300,
                nthreads: int = 8) -> np.ndarray:
    out = np.empty((Cn, T), RECORD_DTYPE)
    lib().orc_gen_records(seed, stream, c0, Cn, t0, T, T, agent_permille, _p(out), nthreads)
    return out


def gen_dims(seed: int, c0: int, Cn: int, t0: int, T: int, agent_permille: int = 300, nthreads: int = 8) -> np.ndarray:
    out = np.empty((Cn, T, NDIM), np.float32)
    lib().orc_gen_dims(seed, c0, Cn, t0, T, T, agent_permille, _p(out), nthreads)
    return out
