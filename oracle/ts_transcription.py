"""Pure-Python transcription of the reference's TypeScript hot path.

TEST INFRASTRUCTURE ONLY (see oracle/apo_oracle.h).  Pinned by tests/golden/ref_*.json: the outputs of the
reference's own method texts executed by oracle/ts_harness (tests/test_reference_pin.py compares this file with them
bit for bit).  This file restates the source text statement by statement on the reference's own object shapes
(ConversationTrace with `summary`, `spans`, `metadata`).  Python floats are IEEE-754
binary64, the same arithmetic as a JS `number`, and Python never fuses multiply-add.

It is the *second*, independent restatement used to cross-check oracle/apo_oracle.c
(which works on packed Form-R records) and to generate tests/golden/*.json.

TCS = src/vs/workbench/contrib/senweaver/common/traceCollectorService.ts
APO = src/vs/workbench/contrib/senweaver/common/apoService.ts
"""
from __future__ import annotations

import math
import struct
from typing import Any

# TCS:766-776 (a JS object literal: lookup by name, `?? 0.05` for unknown names)
WEIGHTS = {
    "user_feedback": 0.25,
    "task_completion": 0.18,
    "tool_success_rate": 0.12,
    "tool_call_reliability": 0.08,
    "tool_call_efficiency": 0.05,
    "tool_duration_efficiency": 0.05,
    "response_efficiency": 0.08,
    "token_efficiency": 0.08,
    "conversation_efficiency": 0.11,
}
# push order of TCS:679..762
DIM_ORDER = [
    "user_feedback", "task_completion", "tool_success_rate", "tool_call_reliability",
    "tool_call_efficiency", "tool_duration_efficiency", "response_efficiency",
    "token_efficiency", "conversation_efficiency",
]


def new_trace(thread_id: str = "t", metadata: dict | None = None, start_time: float = 0.0) -> dict:
    """TCS:380-401 zero state."""
    return {
        "id": "trace", "threadId": thread_id, "startTime": start_time, "spans": [],
        "summary": {
            "totalLLMCalls": 0, "totalToolCalls": 0, "totalTokens": 0, "userFeedback": None,
            "hasErrors": False, "toolCallsSucceeded": 0, "toolCallsFailed": 0, "toolCallsByName": {},
            "totalToolDurationMs": 0, "finalReward": None, "rewardDimensions": [],
        },
        "metadata": metadata,
    }


def compute_reward_signals(trace: dict) -> None:
    """TCS:668-788, statement by statement."""
    dims: list[dict[str, Any]] = []
    s = trace["summary"]

    md = trace.get("metadata") or {}
    chat_mode = md.get("chatMode") or "normal"          # TCS:673
    is_agent = chat_mode == "agent"                      # TCS:674

    fb = s["userFeedback"]
    feedback_score = 1.0 if fb == "good" else (-1.0 if fb == "bad" else 0.0)   # TCS:677-678
    dims.append({"name": "user_feedback", "value": feedback_score})

    completion = 0.5                                      # TCS:682
    if trace.get("endTime") and not s["hasErrors"]:       # TCS:683
        completion = 0.8
    if s["hasErrors"]:                                    # TCS:686
        completion = -0.5
    if fb == "good":                                      # TCS:689
        completion = 1.0
    dims.append({"name": "task_completion", "value": completion})

    if s["totalToolCalls"] > 0:                           # TCS:695
        rate = s["toolCallsSucceeded"] / s["totalToolCalls"]
        dims.append({"name": "tool_success_rate", "value": rate * 2 - 1})
        pen = 1.0
        th = {"severe": 5, "moderate": 3, "minor": 2} if is_agent else {"severe": 3, "moderate": 2, "minor": 1}
        if s["toolCallsFailed"] >= th["severe"]:
            pen = -1.0
        elif s["toolCallsFailed"] >= th["moderate"]:
            pen = -0.5
        elif s["toolCallsFailed"] >= th["minor"]:
            pen = -0.2
        dims.append({"name": "tool_call_reliability", "value": pen})
        ct = {"excellent": 8, "good": 15, "fair": 25} if is_agent else {"excellent": 3, "good": 6, "fair": 10}
        cnt = 1.0
        if s["totalToolCalls"] > ct["fair"]:
            cnt = -0.8
        elif s["totalToolCalls"] > ct["good"]:
            cnt = -0.3
        elif s["totalToolCalls"] > ct["excellent"]:
            cnt = 0.3
        dims.append({"name": "tool_call_efficiency", "value": cnt})
        if s["totalToolDurationMs"] > 0:                  # TCS:721
            avg = s["totalToolDurationMs"] / s["totalToolCalls"]
            ds = 1.0
            if avg > 10000:
                ds = -0.5
            elif avg > 3000:
                ds = 0.0
            elif avg > 1000:
                ds = 0.5
            dims.append({"name": "tool_duration_efficiency", "value": ds})

    if s["totalLLMCalls"] > 0:                            # TCS:733
        thr = 3 if is_agent else 1
        eff = max(-1, 1 - max(0, s["totalLLMCalls"] - thr) * 0.4)
        dims.append({"name": "response_efficiency", "value": eff})

    if s["totalTokens"] > 0:                              # TCS:740
        tt = ({"excellent": 5000, "good": 15000, "fair": 30000} if is_agent
              else {"excellent": 2000, "good": 5000, "fair": 10000})
        ts = 1.0
        if s["totalTokens"] > tt["fair"]:
            ts = -0.5
        elif s["totalTokens"] > tt["good"]:
            ts = 0.0
        elif s["totalTokens"] > tt["excellent"]:
            ts = 0.5
        dims.append({"name": "token_efficiency", "value": ts})

    user_msgs = len([sp for sp in trace["spans"] if sp["type"] == "user_message"])        # TCS:752
    asst_msgs = len([sp for sp in trace["spans"] if sp["type"] == "assistant_message"])   # TCS:753
    turns = min(user_msgs, asst_msgs)
    if turns > 0:
        tthr = 3 if is_agent else 2
        tsc = 1.0
        if turns > tthr * 3:
            tsc = -0.8
        elif turns > tthr * 2:
            tsc = -0.3
        elif turns > tthr:
            tsc = 0.3
        dims.append({"name": "conversation_efficiency", "value": tsc})

    weighted_sum = 0.0
    total_weight = 0.0
    for d in dims:                                        # TCS:779-783
        w = WEIGHTS.get(d["name"], 0.05)
        weighted_sum += d["value"] * w
        total_weight += w
    final = weighted_sum / total_weight if total_weight > 0 else None

    s["rewardDimensions"] = dims
    s["finalReward"] = final


def _extract_mode(trace: dict) -> str:
    md = trace.get("metadata") or {}
    return md["chatMode"] if md.get("chatMode") else "unknown"      # APO:627-633


def analyze_patterns(bad_examples: list, traces: list[dict]) -> list[dict]:
    """APO:635-773 (numeric content; example payloads reduced to the trace index)."""
    patterns: list[dict] = []
    if len(bad_examples) == 0:
        return patterns
    idx = {id(t): i for i, t in enumerate(traces)}

    def emit(pid, matched, minc, severity, category):
        if len(matched) >= minc:
            patterns.append({"pid": pid, "frequency": len(matched), "severity": severity,
                             "relatedCategory": category, "examples": [idx[id(t)] for t in matched[:3]]})

    bad = lambda t: t["summary"]["userFeedback"] == "bad"
    m = [t for t in traces if t["summary"]["hasErrors"] and bad(t)]
    emit(1, m, 2, "high" if len(m) >= 5 else "medium", "core_behavior")
    m = [t for t in traces
         if len([s for s in t["spans"] if s["type"] == "tool_call" and s["data"].get("toolSuccess") is False]) > 0 and bad(t)]
    emit(2, m, 2, "high" if len(m) >= 5 else "medium", "tool_usage")
    m = [t for t in traces if t["summary"]["totalTokens"] > 10000 and bad(t)]
    emit(3, m, 3, "medium", "context_management")
    m = [t for t in traces if t["summary"]["totalLLMCalls"] > 2 and bad(t)]
    emit(4, m, 2, "high", "core_behavior")
    m = [t for t in traces if len([s for s in t["spans"] if s["type"] == "user_message"]) >= 4 and bad(t)]
    emit(5, m, 2, "high" if len(m) >= 4 else "medium", "core_behavior")
    m = [t for t in traces if t["summary"]["totalToolDurationMs"] > 15000 and bad(t)]
    emit(6, m, 2, "medium", "tool_usage")
    return patterns


def build_report(traces: list[dict]) -> dict:
    """APO:498-625, numeric content."""
    good = bad = none = 0
    by_mode: dict[str, dict] = {}
    bad_examples = []
    for t in traces:
        fb = t["summary"]["userFeedback"]
        if fb == "good":
            good += 1
        elif fb == "bad":
            bad += 1
        else:
            none += 1
        mk = _extract_mode(t)
        bm = by_mode.setdefault(mk, {"total": 0, "good": 0, "bad": 0, "goodRate": 0})
        bm["total"] += 1
        if fb == "good":
            bm["good"] += 1
        if fb == "bad":
            bm["bad"] += 1
        if fb == "bad":
            bad_examples.append(t)
    for bm in by_mode.values():
        tot = bm["good"] + bm["bad"]
        bm["goodRate"] = bm["good"] / tot if tot > 0 else 0
    twf = good + bad
    good_rate = good / twf if twf > 0 else 0

    with_reward = [t for t in traces if t["summary"]["finalReward"] is not None]
    avg_reward = None
    if len(with_reward) > 0:
        acc = 0
        for t in with_reward:
            acc = acc + (t["summary"]["finalReward"] or 0)
        avg_reward = acc / len(with_reward)

    rbd: dict[str, dict] = {}
    for t in with_reward:
        for d in t["summary"]["rewardDimensions"]:
            e = rbd.setdefault(d["name"], {"sum": 0, "count": 0, "avg": 0})
            e["sum"] += d["value"]
            e["count"] += 1
    for e in rbd.values():
        e["avg"] = e["sum"] / e["count"] if e["count"] > 0 else 0

    patterns = analyze_patterns(bad_examples, traces)
    dim_patterns = []
    for name, st in rbd.items():
        if st["avg"] < -0.3 and st["count"] >= 5:                    # APO:575
            dim_patterns.append({"dim": name, "frequency": st["count"],
                                 "severity": "high" if st["avg"] < -0.5 else "medium"})
    dim_suggestions = []
    for name, st in rbd.items():
        if st["avg"] < 0 and st["count"] >= 3:                       # APO:802
            dim_suggestions.append({"dim": name, "priority": "high" if st["avg"] < -0.5 else "medium"})

    return {"totalConversations": len(traces), "goodFeedbackCount": good, "badFeedbackCount": bad,
            "noFeedbackCount": none, "goodRate": good_rate, "byMode": by_mode, "avgReward": avg_reward,
            "rewardByDimension": rbd, "patterns": patterns, "dimPatterns": dim_patterns,
            "dimSuggestions": dim_suggestions}


# ------------------------------------------------------------------ Form R encoder (test helper)
MODE_CODE = {None: 0, "normal": 1, "agent": 2, "gather": 3, "designer": 4}
FB_CODE = {None: 0, "good": 1, "bad": 2}
RECORD_STRUCT = struct.Struct("<BBBBHHIIIIIf")
assert RECORD_STRUCT.size == 32


def duration_class(dur, total_tool_calls) -> int:
    """apo_record.durClass: the binary64 comparisons TCS:721-728 / APO:754 make on totalToolDurationMs, decided here because
    the record only carries a binary32 copy of the value."""
    dur = float(dur)
    dc = 0x80
    if dur > 0:
        dc |= 0x04
    if dur > 15000:
        dc |= 0x08
    if total_tool_calls > 0 and dur > 0:
        avg = dur / total_tool_calls
        dc |= (avg > 1000) + (avg > 3000) + (avg > 10000)
    return dc


def encode_record(trace: dict) -> bytes:
    """ConversationTrace -> 32-byte Form R (layout: oracle/apo_oracle.h orc_record)."""
    s = trace["summary"]
    md = trace.get("metadata") or {}
    mode = MODE_CODE.get(md.get("chatMode") or None, 0)
    user = len([sp for sp in trace["spans"] if sp["type"] == "user_message"])
    asst = len([sp for sp in trace["spans"] if sp["type"] == "assistant_message"])
    failspan = any(sp["type"] == "tool_call" and sp["data"].get("toolSuccess") is False for sp in trace["spans"])
    flags = ((1 if s["hasErrors"] else 0) | (2 if trace.get("endTime") else 0) |
             (8 if s["finalReward"] is not None else 0) | (16 if failspan else 0))
    sat32 = lambda v: int(min(max(v, 0), 0xFFFFFFFF))
    return RECORD_STRUCT.pack(FB_CODE[s["userFeedback"]], flags, mode, duration_class(s["totalToolDurationMs"], s["totalToolCalls"]),
                              min(user, 65535), min(asst, 65535),
                              sat32(s["totalToolCalls"]), sat32(s["toolCallsSucceeded"]), sat32(s["toolCallsFailed"]),
                              sat32(s["totalLLMCalls"]), sat32(s["totalTokens"]), float(s["totalToolDurationMs"]))


def make_trace(feedback, has_errors, ended, tool_calls, succ, fail, tool_dur_ms, llm_calls, tokens,
               user_msgs, asst_msgs, mode) -> dict:
    """Build a trace object from the 12-tuple used by the KATs in SURVEY.md 8c."""
    md = None if mode is None else {"chatMode": mode}
    t = new_trace(metadata=md)
    if ended:
        t["endTime"] = 1.0
    s = t["summary"]
    s.update(userFeedback=feedback, hasErrors=has_errors, totalToolCalls=tool_calls, toolCallsSucceeded=succ,
             toolCallsFailed=fail, totalToolDurationMs=tool_dur_ms, totalLLMCalls=llm_calls, totalTokens=tokens)
    t["spans"] = ([{"type": "user_message", "data": {}} for _ in range(user_msgs)] +
                  [{"type": "assistant_message", "data": {}} for _ in range(asst_msgs)] +
                  [{"type": "tool_call", "data": {"toolSuccess": i >= fail}} for i in range(tool_calls)])
    return t


def dims_vector(trace: dict) -> list[float]:
    """rewardDimensions -> 9-vector in push order, NaN = absent."""
    by = {d["name"]: d["value"] for d in trace["summary"]["rewardDimensions"]}
    return [by.get(n, math.nan) for n in DIM_ORDER]


TRICKY = ['plain', 'He said "type":"user_message" twice', 'back\\slash at end\\', 'quote at end\\"', '{"toolSuccess":false}',
          'unicode é中\U0001F600 and \t tab \n newline', '', '[]]]{{{', '\\\\"', 'a' * 700]


def persisted_form(trace: dict, rng, idx: int = 0) -> dict:
    """The object JSON.stringify would see for this trace (TCS:83-109, spans TCS:30-81): ids, timestamps, message
    previews and tool payloads added around the fields the scoring path reads.  Member order is shuffled (JSON objects
    are unordered; a parser must not depend on it) and the free-text fields carry text that looks like structure."""
    def text():
        return rng.choice(TRICKY)
    spans = []
    for k, sp in enumerate(trace["spans"]):
        data = dict(sp["data"])
        if sp["type"] == "tool_call":
            data.update(toolName=rng.choice(["read_file", "run_command", 'we"ird']), toolParams={"uri": text(), "nested": {"type": "user_message", "a": [1, {"toolSuccess": False}]}},
                        toolResult=text())
        else:
            data.update(content=text(), role="user" if sp["type"] == "user_message" else "assistant")
        full = {"id": f"s{idx}-{k}", "traceId": f"trace-{idx}", "threadId": trace["threadId"], "messageIdx": k, "type": sp["type"],
                "timestamp": 1.7e12 + k + 0.25, "data": data}
        if sp["type"] == "tool_call" and rng.random() < 0.7:
            full["duration"] = rng.choice([0, 12.5, 1e3, 2.5e-3])
        items = list(full.items())
        rng.shuffle(items)
        spans.append(dict(items))
    for extra in range(rng.choice([0, 0, 1, 3])):                       # span types the scoring path ignores
        spans.insert(rng.randint(0, len(spans)), {"id": f"x{extra}", "type": rng.choice(["llm_call", "error", "feedback", "tool_result"]), "traceId": f"trace-{idx}",
                                                   "threadId": trace["threadId"], "messageIdx": 0, "timestamp": 1.0, "data": {"errorMessage": text(), "toolSuccess": False}})
    out = {"id": f"trace-{idx}", "threadId": trace["threadId"], "startTime": 1.7e12 + idx, "spans": spans,
           "summary": dict(trace["summary"], toolCallsByName={"read_file": 2, text(): 1})}
    if trace.get("endTime"):
        out["endTime"] = 1.7e12 + idx + 5.5
    elif rng.random() < 0.3:
        out["endTime"] = rng.choice([0, None])                            # falsy forms of `trace.endTime` (TCS:686)
    if trace.get("metadata") is not None:
        out["metadata"] = dict(trace["metadata"], modelName=text(), providerName="p")
    elif rng.random() < 0.5:
        out["metadata"] = rng.choice([None, {}, {"modelName": "m"}])
    items = list(out.items())
    rng.shuffle(items)
    return dict(items)


def upload_summaries(new_traces: list[dict]) -> dict:
    """TCS:805-845 + 860-873: rewardSummary and toolCallSummary of uploadToServer, statement by statement."""
    with_reward = [t for t in new_traces if t["summary"]["finalReward"] is not None]
    avg = None
    if with_reward:
        acc = 0
        for t in with_reward:
            acc = acc + (t["summary"]["finalReward"] or 0)
        avg = acc / len(with_reward)
    succ = fail = dur = 0
    by_name: dict = {}
    for t in new_traces:
        succ += t["summary"]["toolCallsSucceeded"]
        fail += t["summary"]["toolCallsFailed"]
        dur += t["summary"]["totalToolDurationMs"]
        for name, st in t["summary"]["toolCallsByName"].items():
            g = by_name.setdefault(name, {"total": 0, "succeeded": 0, "failed": 0})
            g["total"] += st["total"]; g["succeeded"] += st["succeeded"]; g["failed"] += st["failed"]
    agg: dict = {}
    for t in with_reward:
        for d in t["summary"]["rewardDimensions"]:
            a = agg.setdefault(d["name"], {"sum": 0, "count": 0})
            a["sum"] += d["value"]; a["count"] += 1
    return {"rewardSummary": {"totalTracesWithReward": len(with_reward), "avgFinalReward": avg,
                              "rewardDimensionAvg": {k: (v["sum"] / v["count"] if v["count"] > 0 else 0) for k, v in agg.items()}},
            "toolCallSummary": {"totalToolCalls": succ + fail, "totalSucceeded": succ, "totalFailed": fail,
                                "successRate": succ / (succ + fail) if succ + fail > 0 else None,
                                "totalDurationMs": dur, "byToolName": by_name}}
