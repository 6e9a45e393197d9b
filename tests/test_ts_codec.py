"""ts/traceRecordCodec.ts executed, not only read: `encodeTraceRecord` — the function the patched TraceCollectorService / APOService
call to turn a ConversationTrace into the 32-byte Form R record — runs unmodified in the in-repo TypeScript-subset interpreter
(oracle/ts_harness/minijs.py, the one that executes the reference's own methods) over every trace of the reference pin's inputs,
against a DataView that writes into a bytearray.  The bytes must equal what the Python twins produce (the transcription's
encode_record and the package's encode_trace, which the GPU tests feed to the engine)."""
import importlib
import json
import os
import re
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ts_harness"))
import minijs as js  # noqa: E402

from oracle import ts_transcription as tr  # noqa: E402

SRC = open(os.path.join(ROOT, "ts", "traceRecordCodec.ts"), encoding="utf-8").read()


def function_text(name: str) -> str:
    """`export function name(...) ... { body }` -> method text `name(...) ... { body }` (braces matched outside strings)."""
    start = SRC.index(f"export function {name}(")
    depth, k = 0, SRC.index("(", start)
    while True:                                         # the parameter list
        depth += {"(": 1, ")": -1}.get(SRC[k], 0)
        if depth == 0:
            break
        k += 1
    i = SRC.index("{", k)                                # return annotations here are plain names, no object types
    depth, j = 0, i
    while True:
        c = SRC[j]
        if c in "'\"":
            j = SRC.index(c, j + 1)
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return SRC[start + len("export function "):j + 1]


class View:
    """DataView over a bytearray: the setters encodeTraceRecord uses (little endian where it says so)."""
    def __init__(self, n):
        self.b = bytearray(n)

    def obj(self):
        def setter(fmt_le, fmt_be, size, conv):
            def put(this, off, val, little=False):
                v = conv(js.to_number(val))
                self.b[int(off):int(off) + size] = struct.pack(fmt_le if js.truthy(little) or size == 1 else fmt_be, v)
                return js.undefined
            return js.NativeFunction(put)
        wrap = lambda bits: (lambda x: int(x) & ((1 << bits) - 1))              # ToUint8 / 16 / 32 of an integral double
        return js.JSObject(setUint8=setter("<B", ">B", 1, wrap(8)), setUint16=setter("<H", ">H", 2, wrap(16)),
                           setUint32=setter("<I", ">I", 4, wrap(32)), setFloat32=setter("<f", ">f", 4, float))


def data_view(this, buf, offset=0, length=js.undefined):
    """new DataView(arrayBuffer, byteOffset?, byteLength?) over Python bytes: the getters the decoders use"""
    raw = bytes(buf)[int(offset):] if length is js.undefined else bytes(buf)[int(offset):int(offset) + int(length)]

    def getter(fmt, size):
        def get(this, off, little=False):
            o = int(off)
            if o < 0 or o + size > len(raw):
                raise js.JSThrow("RangeError: offset is outside the bounds of the DataView")
            return struct.unpack(("<" if js.truthy(little) else ">") + fmt, raw[o:o + size])[0]
        return js.NativeFunction(get)
    return js.JSObject(getUint8=getter("B", 1), getUint32=getter("I", 4), getFloat64=getter("d", 8),
                       getBigUint64=getter("Q", 8), getBigInt64=getter("q", 8))


def vsbuffer(b: bytes, pad: int = 0):
    """VSBuffer-like: .buffer is a Uint8Array view {buffer, byteOffset} into a larger ArrayBuffer"""
    return js.JSObject(buffer=js.JSObject(buffer=b"\xAA" * pad + b, byteOffset=pad, byteLength=len(b)))


@pytest.fixture(scope="module")
def codec():
    interp = js.Interp({"DataView": js.NativeFunction(data_view)})
    consts = "\n".join(m.group(0) for m in re.finditer(r"^const (?:F_ERRORS|MODE_CODE|U32_MAX|SEV|MODES|u64)\b.*$", SRC, re.M))
    consts += "\n" + re.search(r"^export (const DIM_NAMES = \[.*?\] as const;)", SRC, re.M | re.S).group(1)
    assert all(w in consts for w in ("F_FAILSPAN", "designer", "0xFFFFFFFF", "conversation_efficiency", "getBigUint64", "'high'"))
    # module-level constants live in the interpreter's global scope, like the reference harness's helpers
    for st in js.Parser(consts).parse_program()[1]:
        interp.exec(st, interp.g, js.undefined)
    fns = {name: interp.make_method(function_text(name), js.undefined)[1] for name in ("encodeTraceRecord", "decodeReward", "decodeCorpusReport")}
    return interp, fns


def encode_ts(codec, trace, *extra):
    interp, fns = codec
    view = View(40)
    interp.call(fns["encodeTraceRecord"], js.undefined, [js.to_js(trace), view.obj(), 4, *extra])
    assert bytes(view.b[:4]) == b"\0\0\0\0" and bytes(view.b[36:]) == b"\0\0\0\0"        # writes its 32 bytes at the offset, nothing else
    return bytes(view.b[4:36])


def golden_traces():
    """the 431 trace tuples of the reference pin (tests/golden/ref_inputs.json), as ConversationTrace objects: unscored and scored"""
    doc = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_inputs.json")))
    out = []
    for _, tup in doc["tuples"]:
        out.append(tr.make_trace(*tup))
        scored = tr.make_trace(*tup)
        tr.compute_reward_signals(scored)
        out.append(scored)
    return out


def test_ts_encoder_equals_the_python_twins_on_every_pinned_trace(codec):
    tc = importlib.import_module("senweaver-ide_b200.trace_collector")
    traces = golden_traces()
    assert len(traces) >= 800
    for t in traces:
        got = encode_ts(codec, t)
        assert got == tr.encode_record(t), t["summary"]
        assert got == tc.encode_trace(t).tobytes()
    # the explicit `valid` argument (what _computeRewardSignals passes while finalReward is still null)
    t = json.loads(json.dumps(traces[0]))
    t["summary"]["finalReward"] = None
    a, b = encode_ts(codec, t), encode_ts(codec, t, True)
    assert a[1] & 0x08 == 0 and b[1] & 0x08 == 0x08 and a[:1] + a[2:] == b[:1] + b[2:]


def test_ts_encoder_edge_cases(codec):
    """saturation, unknown chat modes, fractional durations at the thresholds (durClass is decided on the double)"""
    base = tr.make_trace("good", False, True, 4, 3, 1, 12000.5, 2, 1000, 3, 2, "agent")
    for mut in ({"totalToolCalls": 2 ** 33, "toolCallsSucceeded": 2 ** 33}, {"totalTokens": 4294967295}, {"totalToolDurationMs": 15000.000000000002},
                {"totalToolDurationMs": 4000.0000000000005, "totalToolCalls": 4}, {"totalToolDurationMs": 0}, {"totalToolDurationMs": 2 ** 25 + 1.5}):
        t = json.loads(json.dumps(base))
        t["summary"].update(mut)
        assert encode_ts(codec, t) == tr.encode_record(t), mut
    for mode in ("normal", "gather", "designer", "somethingElse", "", None):
        t = json.loads(json.dumps(base))
        t["metadata"] = None if mode is None else {"chatMode": mode}
        assert encode_ts(codec, t) == tr.encode_record(t), mode
    t = json.loads(json.dumps(base))
    t["spans"] = [{"type": "user_message", "data": {}}] * 70000
    assert encode_ts(codec, t) == tr.encode_record(t)


def test_ts_decoders_read_the_result_blocks_like_the_ctypes_mirror(codec, orc):
    """decodeCorpusReport on the 784 bytes of a real apo_corpus_report (the oracle's struct has the same layout) and decodeReward on
    apo_reward_batch-shaped blocks: every number the patched _buildReport / _computeRewardSignals would read."""
    interp, fns = codec
    recs = orc.gen_records(0x5EED00F1, orc.STREAM_CORPUS, 0, 1, 0, 5000, 400, 2).reshape(-1)
    rep = orc.report(recs, idx_base=100)
    raw = bytes(rep)
    assert len(raw) == 784
    out = js.from_js(interp.call(fns["decodeCorpusReport"], js.undefined, [vsbuffer(raw, pad=16)]))
    assert (out["total"], out["good"], out["bad"], out["none"], out["withReward"]) == (rep.total, rep.good, rep.bad, rep.none, rep.withReward)
    assert out["goodRate"] == rep.goodRate and out["avgReward"] == rep.avgReward
    assert (out["toolCalls"], out["toolSucc"], out["toolFail"], out["toolSuccessRate"]) == (rep.toolCalls, rep.toolSucc, rep.toolFail, rep.toolSuccessRate)
    modes = ["unknown", "normal", "agent", "gather", "designer"]
    assert set(out["byMode"]) == {m for k, m in enumerate(modes) if rep.byMode[k][0]}
    for k, m in enumerate(modes):
        if rep.byMode[k][0]:
            assert out["byMode"][m] == {"total": rep.byMode[k][0], "good": rep.byMode[k][1], "bad": rep.byMode[k][2], "goodRate": rep.byModeGoodRate[k]}
    names = ["user_feedback", "task_completion", "tool_success_rate", "tool_call_reliability", "tool_call_efficiency",
             "tool_duration_efficiency", "response_efficiency", "token_efficiency", "conversation_efficiency"]
    sev = [None, "medium", "high"]
    for i, n in enumerate(names):
        d = rep.dim[i]
        if d.count:
            assert out["rewardByDimension"][n] == {"sum": d.sum, "count": d.count, "avg": d.avg,
                                                   "low": ("high" if d.low_severity == 2 else "medium") if d.low_flag else None,
                                                   "suggest": ("high" if d.sugg_priority == 2 else "medium") if d.sugg_flag else None}
        else:
            assert n not in out["rewardByDimension"]
    assert sev[1] == "medium"
    for p in range(6):
        q = rep.pat[p]
        assert out["patterns"][p] == {"emitted": bool(q.flag), "frequency": q.count, "severity": ["low", "medium", "high"][q.severity],
                                      "examples": [int(x) for x in q.examples if x >= 0]}
    assert any(pp["examples"] and min(pp["examples"]) >= 100 for pp in out["patterns"])          # indices carry the idx_base
    # an empty corpus: nothing with a reward, no tool calls -> the nulls of APO:550-552 / TCS:624
    empty = js.from_js(interp.call(fns["decodeCorpusReport"], js.undefined, [vsbuffer(bytes(orc.report(recs[:0])))]))
    assert empty["avgReward"] is None and empty["toolSuccessRate"] is None and empty["byMode"] == {} and empty["rewardByDimension"] == {}
    # decodeReward: record i of a batch
    sample = recs[:64]
    dims = np.empty((64, 9), np.float64)
    masks = np.empty(64, np.uint32)
    finals = np.empty(64, np.float64)
    for i, r in enumerate(sample):
        d, m, f = orc.reward_one(r)
        dims[i], masks[i], finals[i] = d, m, (np.nan if f is None else f)
    for i in (0, 7, 63):
        got = js.from_js(interp.call(fns["decodeReward"], js.undefined, [vsbuffer(dims.tobytes(), 8), vsbuffer(masks.tobytes()), vsbuffer(finals.tobytes(), 24), i]))
        want = [{"name": n, "value": float(dims[i, k])} for k, n in enumerate(names) if int(masks[i]) >> k & 1]
        assert got["rewardDimensions"] == want
        assert got["finalReward"] == (None if np.isnan(finals[i]) else float(finals[i]))
    null_final = js.from_js(interp.call(fns["decodeReward"], js.undefined, [vsbuffer(dims.tobytes()), vsbuffer(masks.tobytes()), vsbuffer(np.full(64, np.nan).tobytes())]))
    assert null_final["finalReward"] is None
