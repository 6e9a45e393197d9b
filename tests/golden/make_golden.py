"""Generates tests/golden/*.json.

These four fixtures (reward_cases, report_case, generator_case, persisted_traces) are produced by the independent
pure-Python transcription of the reference source (oracle/ts_transcription.py) — KATs K1..K7 of SURVEY.md 8c plus a seeded
random set — and by the C oracle's generator.  Floats are stored as hex strings (bit-exact).  The fixtures that PIN parity
are the ref_*.json files next to them: outputs of the reference's own method texts (oracle/ts_harness/run_reference.py),
which reproduce every value in reward_cases.json (tests/test_reference_pin.py).

    python tests/golden/make_golden.py
"""
import json
import math
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ts_transcription as ts  # noqa: E402
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

KATS = {
    "K1": ("good", False, True, 0, 0, 0, 0, 1, 0, 1, 1, "normal"),
    "K2": ("bad", True, True, 4, 1, 3, 50000, 4, 12000, 1, 4, "normal"),
    "K3": ("bad", True, True, 4, 1, 3, 50000, 4, 12000, 1, 4, "agent"),
    "K4": (None, False, False, 0, 0, 0, 0, 0, 0, 0, 0, "normal"),
    "K5": (None, False, True, 9, 9, 0, 4500, 3, 6000, 1, 3, "agent"),
    "K6": (None, True, True, 0, 0, 0, 0, 7, 100, 7, 7, "normal"),
    "K7": ("bad", False, True, 2, 1, 1, 0, 2, 2001, 3, 3, "normal"),
}


def hexf(x):
    return "nan" if (x is None or (isinstance(x, float) and math.isnan(x))) else float(x).hex()


def random_tuple(rng):
    tc = rng.choice([0, 0, 1, 2, 3, 4, 6, 7, 9, 11, 16, 26, 40])
    fail = rng.randint(0, tc) if tc else 0
    return (rng.choice(["good", "bad", None]), rng.random() < 0.2, rng.random() < 0.9, tc, tc - fail, fail,
            rng.choice([0, 0, 500, 999, 1000, 1001, 2999, 3001, 9000, 10001, 14999, 15001, 60000]) * (1 if tc else 0),
            rng.choice([0, 1, 2, 3, 4, 5, 6, 9]), rng.choice([0, 1, 1999, 2000, 2001, 5001, 10000, 10001, 15001, 30001]),
            rng.choice([0, 1, 2, 3, 4, 5, 6, 7, 9, 10]), rng.choice([0, 1, 2, 3, 4, 6, 7, 10]),
            rng.choice(["normal", "agent", "gather", "designer", None]))


def case(tup):
    t = ts.make_trace(*tup)
    ts.compute_reward_signals(t)
    return {"input": list(tup), "record_hex": ts.encode_record(t).hex(),
            "dims": [hexf(v) for v in ts.dims_vector(t)], "finalReward": hexf(t["summary"]["finalReward"])}


def main():
    rng = random.Random(0x5EED)
    out = {"kats": {k: case(v) for k, v in KATS.items()}, "random": [case(random_tuple(rng)) for _ in range(256)]}
    with open(os.path.join(HERE, "reward_cases.json"), "w") as f:
        json.dump(out, f, indent=0)
    # corpus report on the random traces (numeric content of APO._buildReport)
    traces = []
    for c in out["random"]:
        t = ts.make_trace(*[None if x is None else x for x in c["input"]])
        ts.compute_reward_signals(t)
        traces.append(t)
    rep = ts.build_report(traces)
    rep["avgReward"] = hexf(rep["avgReward"])
    rep["goodRate"] = hexf(rep["goodRate"])
    for v in rep["rewardByDimension"].values():
        v["sum"], v["avg"] = hexf(v["sum"]), hexf(v["avg"])
    for v in rep["byMode"].values():
        v["goodRate"] = hexf(v["goodRate"])
    with open(os.path.join(HERE, "report_case.json"), "w") as f:
        json.dump(rep, f, indent=0)
    # generator fixture: first records of the corpus and rollout streams (C oracle)
    g = {"seed": 0x5EED0003, "agent_permille": 300,
         "corpus": oracle.gen_records(0x5EED0003, oracle.STREAM_CORPUS, 0, 1, 0, 64, 300, 1).tobytes().hex(),
         "rollout_c5_t1000": oracle.gen_records(0x5EED0003, oracle.STREAM_ROLLOUT, 5, 1, 1000, 64, 300, 1).tobytes().hex(),
         "dims_c5_t1000": oracle.gen_dims(0x5EED0003, 5, 1, 1000, 64, 300, 1).tobytes().hex()}
    with open(os.path.join(HERE, "generator_case.json"), "w") as f:
        json.dump(g, f, indent=0)
    # ingest fixture: a persisted trace array (shape of TCS:296-359's JSON) and the Form R bytes it must encode to
    rng = random.Random(0x1265)
    persisted = []
    for i in range(24):
        t = ts.make_trace(*random_tuple(rng))
        if i % 5:
            ts.compute_reward_signals(t)
        persisted.append(ts.persisted_form(t, rng, i))
    fixture = {"json": json.dumps(persisted, separators=(",", ":"), ensure_ascii=False),
               "records_hex": b"".join(ts.encode_record(t) for t in persisted).hex()}
    with open(os.path.join(HERE, "persisted_traces.json"), "w") as f:
        json.dump(fixture, f)


if __name__ == "__main__":
    main()
