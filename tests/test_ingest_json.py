"""Ingest: the reference's persisted trace array -> Form R through the C ABI's host-side
format code (apo_records_from_json; no GPU involved, so this runs in the CPU suite).  The checker is the
independent transcription's encode_record over the same trace objects, plus a committed golden fixture."""
import json
import os
import random
from importlib import import_module

import numpy as np
import pytest

import importlib.util

from oracle import ts_transcription as ts

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
_mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mg)
random_tuple = _mg.random_tuple


@pytest.fixture(scope="module")
def pkg():
    return import_module("senweaver-ide_b200")


def corpus(seed, n):
    rng = random.Random(seed)
    traces, persisted = [], []
    for i in range(n):
        t = ts.make_trace(*random_tuple(rng))
        t["threadId"] = f"th-{i}"
        if rng.random() < 0.8:
            ts.compute_reward_signals(t)
        p = ts.persisted_form(t, rng, i)
        traces.append(p)
        persisted.append(p)
    return traces, persisted


def expected(traces):
    return b"".join(ts.encode_record(t) for t in traces)


@pytest.mark.parametrize("dump", [
    dict(separators=(",", ":")),                       # what JSON.stringify emits
    dict(indent=2), dict(indent="\t", ensure_ascii=False), dict(separators=(" ,\n", " :\r\n "), ensure_ascii=False)])
def test_ingest_matches_transcription(pkg, dump):
    traces, persisted = corpus(0xA11CE, 400)
    text = json.dumps(persisted, **dump)
    recs = pkg.records_from_json(text)
    assert recs.shape[0] == 400 and recs.tobytes() == expected(traces)
    # the python-side encoder of the service mirror agrees as well
    enc = import_module("senweaver-ide_b200.trace_collector").encode_trace
    assert b"".join(enc(t).tobytes() for t in traces) == recs.tobytes()


def test_ingest_golden_fixture(pkg):
    g = json.load(open(os.path.join(HERE, "golden", "persisted_traces.json")))
    recs = pkg.records_from_json(g["json"])
    assert recs.tobytes().hex() == g["records_hex"]


def test_ingest_edge_cases(pkg):
    assert pkg.records_from_json("[]").shape[0] == 0
    assert pkg.records_from_json("  [ ]\n").shape[0] == 0
    r = pkg.records_from_json('[{}]')                                  # a trace with nothing in it: zero record, not valid
    assert r.shape[0] == 1 and r["durClass"][0] == 0x80               # ... whose duration class says "0 ms: nothing holds"
    z = r.copy(); z["durClass"] = 0
    assert z.tobytes() == bytes(32)
    r = pkg.records_from_json('[{"summary":{"totalTokens":1e12,"totalLLMCalls":-3,"totalToolDurationMs":1234.5678,"finalReward":0,"hasErrors":true},"endTime":1}]')
    assert r["tokens"][0] == 0xFFFFFFFF and r["llmCalls"][0] == 0 and r["toolDurMs"][0] == np.float32(1234.5678)
    assert r["flags"][0] == 0x01 | 0x02 | 0x08
    assert r["durClass"][0] == 0x80 | 0x04                            # duration > 0, but no tool calls: no average to classify
    r = pkg.records_from_json('[{"summary":{"totalToolCalls":1,"totalToolDurationMs":3000.0000001}},{"summary":{"totalToolCalls":2,"totalToolDurationMs":30000.5}}]')
    assert r["toolDurMs"][0] == np.float32(3000.0) and r["durClass"][0] == 0x80 | 0x04 | 2      # the double is > 3000, its float32 copy is not
    assert r["durClass"][1] == 0x80 | 0x04 | 0x08 | 3
    # 300 spans: counts come from the spans actually stored (the reference caps at 200 before persisting, TCS:274-280)
    spans = [{"type": "user_message", "data": {}}] * 70000
    r = pkg.records_from_json(json.dumps([{"spans": spans}]))
    assert r["userMsgs"][0] == 65535
    for bad in ["", "{", "[{]", '[{"a":}]', '[{"summary":{"x":1}} {"y":2}]', '[{"spans":[{"type":"user_message"]}]', "[1]", '[{"a":"unterminated]',
                "[{}] trailing", '[{"a":tru}]', "[" * 400]:
        with pytest.raises(ValueError):
            pkg.records_from_json(bad)


def test_ingest_throughput_is_reported(pkg):
    import time
    traces, persisted = corpus(7, 300)
    text = json.dumps(persisted * 20, separators=(",", ":")).encode()
    t0 = time.perf_counter()
    recs = pkg.records_from_json(text)
    dt = time.perf_counter() - t0
    assert recs.shape[0] == 6000
    print(f"ingest: {len(text) / 1e6:.1f} MB in {dt * 1e3:.1f} ms = {len(text) / dt / 1e6:.0f} MB/s, {recs.shape[0] / dt / 1e3:.0f} k traces/s")


def test_ingest_never_crashes_on_mutated_input(pkg):
    """Byte-level mutations of a valid document: the parser returns a count or rejects with a position, nothing else."""
    traces, persisted = corpus(5, 12)
    good = json.dumps(persisted, separators=(",", ":")).encode()
    rng = random.Random(1234)
    rejected = 0
    for _ in range(1500):
        b = bytearray(good)
        for _ in range(rng.randint(1, 4)):
            op, pos = rng.randint(0, 3), rng.randrange(len(b))
            if op == 0:
                b[pos] = rng.randrange(256)
            elif op == 1:
                del b[pos:pos + rng.randint(1, 40)]
            elif op == 2:
                b[pos:pos] = bytes(rng.choice(b'{}[]",:\\0-9etn ') for _ in range(rng.randint(1, 6)))
            else:
                b = b[:pos]
            if not b:
                b = bytearray(b" ")
        try:
            recs = pkg.records_from_json(bytes(b))
            assert recs.shape[0] <= 12 + 6
        except ValueError as e:
            rejected += 1
            assert "malformed trace JSON at byte" in str(e)
    assert rejected > 500


def test_ingest_agrees_with_a_general_json_parser_on_random_documents(pkg):
    """Random nested JSON around the fields that matter: whatever json.loads makes of the document, encode_trace of
    that equals the native record (member order, nesting depth, odd value types in ignored positions)."""
    enc = import_module("senweaver-ide_b200.trace_collector").encode_trace
    rng = random.Random(77)

    def junk(depth=0):
        k = rng.randint(0, 7 if depth < 4 else 4)
        if k == 0:
            return None
        if k == 1:
            return rng.choice([True, False])
        if k == 2:
            return rng.choice([0, -1, 3.5, 1e-7, 1.5e300, 12345678901234567890, -0.0])
        if k in (3, 4):
            return rng.choice(ts.TRICKY)
        if k == 5:
            return [junk(depth + 1) for _ in range(rng.randint(0, 4))]
        return {rng.choice(["type", "data", "toolSuccess", "summary", "spans", "x", "endTime", 'q"k']): junk(depth + 1) for _ in range(rng.randint(0, 4))}

    docs = []
    for i in range(300):
        t = ts.make_trace(*random_tuple(rng))
        if rng.random() < 0.7:
            ts.compute_reward_signals(t)
        p = ts.persisted_form(t, rng, i)
        p["extra"] = junk()
        for sp in p["spans"]:
            if rng.random() < 0.3:
                sp["data"]["payload"] = junk()
        p["summary"]["toolCallsByName"] = {"a": junk(), "type": "user_message"}
        docs.append(p)
    text = json.dumps(docs)
    recs = pkg.records_from_json(text)
    want = b"".join(enc(t).tobytes() for t in json.loads(text))
    assert recs.tobytes() == want
