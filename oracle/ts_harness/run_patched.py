#!/usr/bin/env python
"""Runs the PATCHED reference services — what a senweaver-ide checkout looks like after ts/patches/*.ed — end to end on the CPU.

Test infrastructure (like everything under oracle/): the method texts of the patched traceCollectorService.ts / apoService.ts
(`_computeRewardSignals`, `_refreshEngineStats`, `getStats`; `_buildReport`, `_analyzePatterns`, `_evaluateBeam`, `_applyBeamUpdate` + the
untouched `_extractMode`, `_generateLocalSuggestions`, `getStats`) and the functions of ts/traceRecordCodec.ts are executed unmodified by
oracle/ts_harness/minijs.py; `IApoScoringService` is a stand-in that answers `rewardBatch` / `score` with the byte blocks the
C ABI returns (apo_reward_batch's dims / masks / finals, struct apo_corpus_report), computed here by the ORACLE — so the whole
TypeScript data flow (trace -> 32-byte record -> service call -> result block -> decoded numbers -> report object) runs, and its
output can be held against the fixtures the UNPATCHED reference produced (tests/golden/ref_*.json).

Promises are settled synchronously (minijs.SyncPromise): the data flow is checked, not the scheduling.
tests/test_ts_patched_services.py drives this; it needs the reference checkout (the patches apply to it).
"""
from __future__ import annotations

import os
import re
import struct
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "ts", "patches"))
import make_patches as mp  # noqa: E402
import minijs as js  # noqa: E402
import run_reference as rr  # noqa: E402

import oracle  # noqa: E402

CODEC = os.path.join(ROOT, "ts", "traceRecordCodec.ts")


# ------------------------------------------------------------------------------------------------ host objects for the codec
class ByteBuf:
    """an ArrayBuffer"""
    def __init__(self, n_or_bytes):
        self.b = bytearray(n_or_bytes)


def uint8array(this, arg=0, offset=0, length=js.undefined):
    buf = arg if isinstance(arg, ByteBuf) else ByteBuf(int(js.to_number(arg)))
    off = int(offset)
    n = len(buf.b) - off if length is js.undefined else int(length)
    copy = js.NativeFunction(lambda this: uint8array(None, ByteBuf(buf.b[off:off + n])))       # .slice(): a copy with its own ArrayBuffer

    def set_(this, src, at=0):                                                                 # .set(typedArray, offset)
        data = bytes(src["buffer"].b[src["byteOffset"]:src["byteOffset"] + src["byteLength"]])
        at = int(at)
        if at < 0 or at + len(data) > n:
            raise js.JSThrow("RangeError: offset is out of bounds")
        buf.b[off + at:off + at + len(data)] = data
        return js.undefined
    return js.JSObject(buffer=buf, byteOffset=off, byteLength=n, length=n, slice=copy, set=js.NativeFunction(set_))


def typed_array(fmt: str, size: int):
    """new Float64Array(arrayBuffer) / new Int32Array(arrayBuffer): little endian like every platform the IDE runs on"""
    def ctor(this, buf):
        assert isinstance(buf, ByteBuf) and len(buf.b) % size == 0
        return js.JSArray(struct.unpack(f"<{len(buf.b) // size}{fmt}", bytes(buf.b)))
    return js.NativeFunction(ctor)


def data_view(this, buf, offset=0, length=js.undefined):
    assert isinstance(buf, ByteBuf), "DataView over something that is not an ArrayBuffer"
    base = int(offset)
    size = len(buf.b) - base if length is js.undefined else int(length)

    def rng(off, n):
        o = int(off)
        if o < 0 or o + n > size:
            raise js.JSThrow("RangeError: offset is outside the bounds of the DataView")
        return base + o

    def getter(fmt, n):
        return js.NativeFunction(lambda this, off, little=False: struct.unpack(("<" if js.truthy(little) else ">") + fmt, buf.b[rng(off, n):rng(off, n) + n])[0])

    def setter(fmt, n, conv):
        def put(this, off, val, little=False):
            o = rng(off, n)
            buf.b[o:o + n] = struct.pack(("<" if js.truthy(little) else ">") + fmt, conv(js.to_number(val)))
            return js.undefined
        return js.NativeFunction(put)
    wrap = lambda bits: (lambda x: int(x) & ((1 << bits) - 1))
    return js.JSObject(getUint8=getter("B", 1), getUint32=getter("I", 4), getFloat64=getter("d", 8), getBigUint64=getter("Q", 8),
                       getBigInt64=getter("q", 8), setUint8=setter("B", 1, wrap(8)), setUint16=setter("H", 2, wrap(16)),
                       setUint32=setter("I", 4, wrap(32)), setFloat32=setter("f", 4, float))


def vsbuffer(data: bytes):
    """VSBuffer.wrap(new Uint8Array(data))"""
    return js.JSObject(buffer=uint8array(None, ByteBuf(data)))


def vsbuffer_bytes(v) -> bytes:
    u8 = v["buffer"]
    return bytes(u8["buffer"].b[u8["byteOffset"]:u8["byteOffset"] + u8["byteLength"]])


def codec_function(src: str, name: str) -> str:
    start = src.index(f"export function {name}(")
    depth, k = 0, src.index("(", start)
    while True:
        depth += {"(": 1, ")": -1}.get(src[k], 0)
        if depth == 0:
            break
        k += 1
    depth, j = 0, src.index("{", k)
    while True:
        c = src[j]
        if c in "'\"":
            j = src.index(c, j + 1)
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return src[start + len("export function "):j + 1]


# ------------------------------------------------------------------------------------------------ the engine's stand-in
class OracleScoring:
    """IApoScoringService as the patched services see it; the numbers come from the oracle (in the IDE: from the B200 engine)."""
    def __init__(self):
        self.calls = {"rewardBatch": 0, "score": 0}

    def reward_batch(self, this, records):
        self.calls["rewardBatch"] += 1
        recs = np.frombuffer(vsbuffer_bytes(records), dtype=oracle.RECORD_DTYPE)
        dims = np.empty((len(recs), 9), np.float64)
        masks = np.empty(len(recs), np.uint32)
        finals = np.empty(len(recs), np.float64)
        for i, r in enumerate(recs):
            d, m, f = oracle.reward_one(r)
            dims[i], masks[i], finals[i] = d, m, (np.nan if f is None else f)
        return js.SyncPromise(js.JSObject(dims=vsbuffer(dims.tobytes()), masks=vsbuffer(masks.tobytes()), finals=vsbuffer(finals.tobytes())))

    def score(self, this, dims, C, T, corpus=js.undefined, K=0):
        self.calls["score"] += 1
        C, T, K = int(C), int(T), int(K)
        recs = np.frombuffer(vsbuffer_bytes(corpus), dtype=oracle.RECORD_DTYPE) if corpus is not js.undefined and corpus is not None else np.empty(0, oracle.RECORD_DTYPE)
        rep = oracle.report(recs)
        d = np.frombuffer(vsbuffer_bytes(dims), dtype=np.float32)
        assert d.size == C * T * 9, "dims block is not float32[C][T][9]"
        scores, counts = oracle.score_dims(d.reshape(C, T, 9))
        topk = oracle.topk(scores, K) if K else np.empty(0, np.int32)
        return js.SyncPromise(js.JSObject(scores=vsbuffer(np.asarray(scores, "<f8").tobytes()), counts=vsbuffer(np.asarray(counts, "<u8").tobytes()),
                                          topk=vsbuffer(np.asarray(topk, "<i4").tobytes()), report=vsbuffer(bytes(rep))))


# ------------------------------------------------------------------------------------------------ patched services
class Patched(rr.Reference):
    def __init__(self, ref_root: str):
        self.tmp = tempfile.TemporaryDirectory()
        for rel in (rr.TCS_REL, rr.APO_REL):
            path = os.path.join(self.tmp.name, rel)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "w", encoding="utf-8") as f:
                f.write(mp.patched_text(ref_root, rel))
        super().__init__(self.tmp.name)                 # extracts the (now patched) methods
        self.scoring = OracleScoring()
        g = self.interp.g.vars
        src = open(CODEC, encoding="utf-8").read()
        consts = "\n".join(m.group(0).replace("export ", "") for m in re.finditer(
            r"^(?:export )?const (?:APO_RECORD_BYTES|F_ERRORS|MODE_CODE|U32_MAX|SEV|MODES|u64)\b.*$", src, re.M))
        consts += "\n" + re.search(r"^export (const DIM_NAMES = \[.*?\] as const;)", src, re.M | re.S).group(1)
        g.update({"DataView": js.NativeFunction(data_view), "Uint8Array": js.NativeFunction(uint8array),
                  "Float64Array": typed_array("d", 8), "Int32Array": typed_array("i", 4),
                  "VSBuffer": js.JSObject(wrap=js.NativeFunction(lambda this, u8: js.JSObject(buffer=u8))),
                  "queueMicrotask": js.NativeFunction(lambda this, f: self.interp.call(f, js.undefined, [])),
                  "EMPTY_DIMS": vsbuffer(np.full(36, np.nan, np.float32).tobytes())})
        for st in js.Parser(consts).parse_program()[1]:
            self.interp.exec(st, self.interp.g, js.undefined)
        for name in ("encodeTraceRecord", "encodeTraceRecords", "decodeReward", "decodeCorpusReport"):
            g[name] = self.interp.make_method(codec_function(src, name), js.undefined)[1]
        svc = js.JSObject(rewardBatch=js.NativeFunction(self.scoring.reward_batch), score=js.NativeFunction(self.scoring.score))
        fire = js.JSObject(fire=js.NativeFunction(lambda this, *a: js.undefined))
        self.tcs.update(_scoring=svc, _engineStats=None, _dirty=False, _onDidChangeState=fire,
                        _saveToStorage=js.NativeFunction(lambda this: js.undefined))
        self.tcs["_refreshEngineStats"] = self._method(self.tcs_text, "TCS", "_refreshEngineStats", self.tcs)
        self.adopted = []                               # prompts handed to _applyBeamBestPrompt (APO:1161: auto-apply of a new best)
        self.apo.update(_scoring=svc, _dirty=False, _onDidChangeState=fire, _config=js.JSObject(beamWidth=4, beamRounds=3),
                        _applyBeamBestPrompt=js.NativeFunction(lambda this, p: self.adopted.append(js.from_js(p)) or js.undefined))
        for name in ("_applyBeamUpdate", "_evaluateBeam"):
            self.apo[name] = self._method(self.apo_text, "APO", name, self.apo)

    def evaluate_beam(self, candidates: list, dims: np.ndarray):
        """patched APOService._evaluateBeam(candidates, evaluations float32[C][T][9], T) -> the beam state it leaves"""
        C, T, _ = dims.shape
        self.interp.call(self.apo["_evaluateBeam"], self.apo, [js.to_js(candidates), vsbuffer(np.ascontiguousarray(dims, "<f4").tobytes()), T])
        return js.from_js(self.apo["_beamState"])

    def stats(self, traces):
        self.tcs["_traces"] = js.JSMap((t["id"], js.to_js(t)) for t in traces)
        self.interp.call(self.tcs["_refreshEngineStats"], self.tcs, [])      # what the constructor / every scored trace triggers
        return super().stats(traces)


def build_outputs(ref_root: str):
    """The same cases as run_reference.build_fixtures, through the patched services."""
    from oracle import ts_transcription as ts
    P = Patched(ref_root)
    cases = []
    for name, tup in rr.golden_inputs():
        t = ts.make_trace(*tup)
        s = P.compute_reward_signals(t)["summary"]
        cases.append({"name": name, "dims": [{"name": d["name"], "value": rr.hexf(d["value"])} for d in s["rewardDimensions"]],
                      "finalReward": rr.hexf(s["finalReward"])})
    reports = {}
    for cname, idx in rr.corpus_index().items():
        traces = rr.corpus_traces(cname, idx)
        res = P.build_report(traces)
        res["stats"] = P.stats(res["traces"])
        del res["traces"]
        reports[cname] = rr.hexify(res)
    return cases, reports, dict(P.scoring.calls), P.lines
