"""TraceCollectorService / APOService mirrors driven the way the reference's callers drive them
(CTS:1120-1738, 2745-2746; SidebarChat.tsx:4378), checked against the pure-Python transcription
of the reference source (oracle/ts_transcription.py).  -m gpu: every reduction runs on the B200."""
import json
import math
import random
from importlib import import_module

import numpy as np
import pytest

from oracle import ts_transcription as ts

pytestmark = pytest.mark.gpu


@pytest.fixture()
def services(engine):
    pkg = import_module("senweaver-ide_b200")
    tc = import_module("senweaver-ide_b200.trace_collector").TraceCollectorService(engine, storageService={})
    apo = import_module("senweaver-ide_b200.apo_service").APOService(engine, tc, storageService={})
    return pkg, tc, apo


def drive(tc, rng, n_threads=60):
    """Replays random chat turns through the recorder API; returns nothing (state lives in tc)."""
    for i in range(n_threads):
        th = f"thread-{i}"
        mode = rng.choice(["agent", "normal", "gather", "designer", None])
        tc.startTrace(th, {"chatMode": mode} if mode else None)
        for m in range(rng.choice([1, 1, 2, 4, 5, 8])):
            tc.recordUserMessage(th, m, "please fix the bug " * rng.randint(1, 40))
        for k in range(rng.choice([0, 1, 2, 3, 5])):
            tc.recordLLMCall(th, k, {"model": "m", "inputTokens": rng.choice([0, 500, 3000, 9000]), "outputTokens": rng.choice([0, 200, 4000])})
            tc.recordAssistantMessage(th, k, "done", "m", "p")
        for k in range(rng.choice([0, 0, 1, 3, 7, 12, 30])):
            tc.recordToolCall(th, k, {"toolName": rng.choice(["read_file", "run_command"]), "toolSuccess": rng.random() > 0.2,
                                      "toolResult": "x" * 10, "duration": rng.choice([None, 0, 120.5, 2500.25, 20000.0])})
        if rng.random() < 0.15:
            tc.recordError(th, 0, "boom")
        if rng.random() < 0.9:
            tc.endTraceForThread(th)
        fb = rng.choice(["good", "bad", "bad", None, "skip"])
        if fb != "skip":
            tc.recordUserFeedback(th, 0, fb)
        if rng.random() < 0.1:                            # late event after scoring: counters move, reward does not (TCS:420-425)
            tc.recordToolCall(th, 99, {"toolName": "late", "toolSuccess": False, "duration": 99999.0})


def test_reward_signals_match_transcription(services):
    _, tc, _ = services
    drive(tc, random.Random(1))
    checked = 0
    for t in tc.getAllTraces():
        if t["summary"]["finalReward"] is None:
            continue
        # the stored reward was computed from a snapshot; recompute it from the same snapshot state
        snap = tc._scored[t["id"]]
        rec_now = import_module("senweaver-ide_b200.trace_collector").encode_trace(t, valid=True)
        if snap.tobytes() != rec_now.tobytes():
            continue                                        # mutated after scoring; covered by the report test
        ref = {k: (dict(v) if isinstance(v, dict) else v) for k, v in t.items()}
        ref["summary"] = dict(t["summary"])
        ts.compute_reward_signals(ref)
        assert t["summary"]["finalReward"] == ref["summary"]["finalReward"]
        assert t["summary"]["rewardDimensions"] == ref["summary"]["rewardDimensions"]
        checked += 1
    assert checked > 30


def test_never_scored_trace_keeps_null_reward(services):
    _, tc, _ = services
    tc.startTrace("a")
    tc.recordLLMCall("a", 0, {"inputTokens": 10})
    assert tc.getAllTraces()[0]["summary"]["finalReward"] is None
    st = tc.getStats()
    assert st["avgFinalReward"] is None and st["tracesWithReward"] == 0 and st["totalTraces"] == 1


def test_build_report_matches_transcription(services):
    _, tc, apo = services
    drive(tc, random.Random(7), n_threads=120)
    traces = tc.getAllTraces()
    rep = apo.analyzePromptEffectiveness()
    ref = ts.build_report(traces)
    for k in ("totalConversations", "goodFeedbackCount", "badFeedbackCount", "noFeedbackCount", "goodRate"):
        assert rep[k] == ref[k], k
    assert {k: (v["total"], v["good"], v["bad"], v["goodRate"]) for k, v in rep["byMode"].items()} == \
           {k: (v["total"], v["good"], v["bad"], v["goodRate"]) for k, v in ref["byMode"].items()}
    assert (rep["avgReward"] is None) == (ref["avgReward"] is None)
    if ref["avgReward"] is not None:
        assert abs(rep["avgReward"] - ref["avgReward"]) <= 1e-5 * max(abs(ref["avgReward"]), 1e-6)
    assert set(rep["rewardByDimension"]) == set(ref["rewardByDimension"])
    for n, e in ref["rewardByDimension"].items():
        g = rep["rewardByDimension"][n]
        assert g["count"] == e["count"] and abs(g["avg"] - e["avg"]) <= 1e-5 * max(abs(e["avg"]), 1e-6)
    # the six patterns: emitted set, frequency, severity, first-3 examples (by thread id)
    got = [p for p in rep["patterns"] if "dimension reward signal" not in p["description"]]
    assert len(got) == len(ref["patterns"])
    for g, e in zip(got, ref["patterns"]):
        assert (g["frequency"], g["severity"], g["relatedCategory"]) == (e["frequency"], e["severity"], e["relatedCategory"])
        assert [x["threadId"] for x in g["examples"]] == [traces[i]["threadId"] for i in e["examples"]]
    dimp = [p for p in rep["patterns"] if "dimension reward signal" in p["description"]]
    assert sorted((p["description"].split()[0], p["severity"], p["frequency"]) for p in dimp) == \
           sorted((d["dim"], d["severity"], d["frequency"]) for d in ref["dimPatterns"])
    st = tc.getStats()
    assert st["totalToolCalls"] == sum(t["summary"]["totalToolCalls"] for t in traces)
    vals = [t["summary"]["finalReward"] for t in traces if t["summary"]["finalReward"] is not None]
    assert st["tracesWithReward"] == len(vals)
    assert abs(st["avgFinalReward"] - sum(vals) / len(vals)) < 1e-12


def test_beam_update_strict_greater(services):
    pkg, tc, apo = services
    rng = np.random.default_rng(0)
    C, T = 8, 512
    dims = rng.uniform(-1, 1, (C, T, 9)).astype(np.float32)
    dims[5] = dims[2]                                        # tie: candidate 2 must rank before 5
    cands = [{"version": f"v{c}", "content": f"- rule {c}\n- shared rule", "score": None, "createdAt": 0} for c in range(C)]
    res = apo.evaluateBeam(cands, dims=dims)
    st = apo.getBeamState()
    assert [b["version"] for b in st["beam"]] == [f"v{c}" for c in res.topk] and len(st["beam"]) == 4
    assert st["historyBestScore"] == res.scores[res.topk[0]]
    order = list(res.topk)
    if 2 in order and 5 in order:
        assert order.index(2) < order.index(5)
    rules = apo.getOptimizedRules()
    assert f"rule {res.topk[0]}" in rules and "shared rule" in rules
    # same scores again: equal is not greater -> the incumbent stays (APO:1159), no duplicate segments
    apo.evaluateBeam(cands, dims=dims)
    st2 = apo.getBeamState()
    assert st2["historyBestPrompt"]["version"] == st["historyBestPrompt"]["version"] and st2["currentRound"] == 2
    assert apo.getOptimizedRules() == rules


def test_engine_failure_never_raises_into_callers(services, monkeypatch):
    _, tc, apo = services
    tc.startTrace("x")

    def boom(*a, **k):
        raise RuntimeError("device lost")
    monkeypatch.setattr(tc._engine, "reward_batch", boom)
    tc.recordUserFeedback("x", 0, "good")                     # swallowed (TCS:554)
    assert tc.getAllTraces()[0]["summary"]["finalReward"] is None
    assert apo.evaluateBeam([{"version": "v0", "content": "c"}], dims=None, rollouts=None) is None


def test_payload_and_rule_budget(services):
    _, tc, apo = services
    drive(tc, random.Random(3), n_threads=40)
    pay = apo.buildOptimizePayload()
    recent = sorted([t for t in tc.getAllTraces() if t["summary"]["userFeedback"] is not None], key=lambda t: -t["startTime"])[:16]
    assert [r["traceId"] for r in pay["rolloutResults"]] == [t["id"] for t in recent]
    vals = [t["summary"]["finalReward"] for t in recent if t["summary"]["finalReward"] is not None]
    assert pay["rewardSummary"]["totalWithReward"] == len(vals)
    if vals:
        assert abs(pay["rewardSummary"]["avgFinalReward"] - sum(vals) / len(vals)) < 1e-12
    for r, t in zip(pay["rolloutResults"], recent):
        fb = t["summary"]["userFeedback"]
        assert r["status"] == ("succeeded" if fb == "good" else "failed")
    # rule budget (C2L:832-854)
    apo._segments = [{"id": str(i), "category": "core_behavior", "content": "x" * 300, "isActive": True, "isOptimized": True,
                      "version": 1, "createdAt": 0, "updatedAt": 0} for i in range(10)]
    block = apo.packOptimizedRules()
    assert "(6/10 rules, budget limited)" in block and len(block.split("\n", 3)[3]) <= 2000


def test_beam_search_round_driver(services):
    """A toy search space: candidate quality = number of 'good' tokens in its text; the driver must climb."""
    _, tc, apo = services
    rng = np.random.default_rng(4)

    def propose(parent, n):
        return [parent["content"] + (" good" if rng.random() < 0.6 else " bad") for _ in range(n)]

    def rollouts(cands):
        out = np.full((len(cands), 256, 9), np.nan, np.float32)
        for i, c in enumerate(cands):
            q = (c["content"].count("good") - c["content"].count("bad")) / 10.0
            out[i, :, 0] = np.clip(q + rng.normal(0, 0.01, 256), -1, 1)
        return out

    st = apo.runBeamSearch("- be helpful", propose, rollouts, rounds=3)
    assert st["currentRound"] == 3 and len(st["beam"]) == 4
    scores = [b["score"] for b in st["beam"]]
    assert scores == sorted(scores, reverse=True) and st["historyBestScore"] >= scores[0]
    assert st["historyBestPrompt"]["content"].count("good") >= 1


def test_storage_round_trip_and_trace_cap(engine):
    """Persisted JSON under the reference's storage keys (TCS:216-217) reloads into an equivalent service;
    more than MAX_TRACES traces are trimmed to the newest 1000 by startTime on save (TCS:337-345)."""
    tcmod = import_module("senweaver-ide_b200.trace_collector")
    store = {}
    tc = tcmod.TraceCollectorService(engine, storageService=store)
    drive(tc, random.Random(11), n_threads=30)
    tc._dirty = True
    tc._saveToStorage()
    assert set(store) >= {"senweaver.traceCollector.data", "senweaver.traceCollector.feedbacks"}
    tc2 = tcmod.TraceCollectorService(engine, storageService=store)
    a, b = tc.getStats(), tc2.getStats()
    for k in ("totalTraces", "totalSpans", "goodFeedbacks", "badFeedbacks", "totalToolCalls", "totalToolSucceeded",
              "totalToolFailed", "toolSuccessRate", "tracesWithReward"):
        assert a[k] == b[k], k
    assert abs(a["avgFinalReward"] - b["avgFinalReward"]) < 1e-12
    # cap
    for i in range(tcmod.MAX_TRACES + 25):
        tid = tc.startTrace(f"cap-{i}")
        tc._traces[tid]["startTime"] = 1e12 + i
    tc._saveToStorage()
    assert len(tc.getAllTraces()) == tcmod.MAX_TRACES
    assert min(t["startTime"] for t in tc.getAllTraces()) >= 1e12 + 25


def test_span_cap_keeps_counting(engine):
    """Past 200 spans the span object is dropped but the counters keep advancing (TCS:274-280, 505-510)."""
    tcmod = import_module("senweaver-ide_b200.trace_collector")
    tc = tcmod.TraceCollectorService(engine)
    tc.startTrace("t", {"chatMode": "agent"})
    for i in range(260):
        tc.recordToolCall("t", i, {"toolName": "x", "toolSuccess": i % 10 != 0, "duration": 5.0})
    tr = tc.getAllTraces()[0]
    assert len(tr["spans"]) == tcmod.MAX_SPANS_PER_TRACE and tr["summary"]["totalToolCalls"] == 260
    tc.endTraceForThread("t")
    names = {d["name"]: d["value"] for d in tr["summary"]["rewardDimensions"]}
    assert names["tool_call_efficiency"] == -0.8 and names["tool_call_reliability"] == -1.0      # agent thresholds: > 25 calls, >= 5 failures
    assert names["tool_success_rate"] == (234 / 260) * 2 - 1


def test_suggestion_lifecycle(services):
    _, tc, apo = services
    drive(tc, random.Random(5), n_threads=50)
    rep = apo.analyzePromptEffectiveness()
    pend = apo.getPendingSuggestions()
    assert len(pend) == len(rep["suggestions"])
    if pend:
        s0 = pend[0]
        s0["suggestedContent"], s0["type"] = "- always run the tests", "add"
        apo.applySuggestion(s0["id"])
        assert "- always run the tests" in apo.getOptimizedRules() and apo.getStats()["appliedSuggestions"] == 1
        apo.revertSuggestion(s0["id"])
        assert "- always run the tests" not in apo.getOptimizedRules()
        if len(pend) > 1:
            apo.rejectSuggestion(pend[1]["id"])
            assert apo.getStats()["rejectedSuggestions"] == 1
    st = apo.getStats()
    assert st["totalReports"] == 1 and st["currentGoodRate"] == rep["goodRate"]
    cfg = apo.getConfig()
    apo.setConfig({"beamWidth": 2})
    assert apo.getConfig()["beamWidth"] == 2 and cfg["beamWidth"] == 4


def test_persisted_json_ingest_feeds_the_corpus_scan(engine, orc):
    """Storage JSON written by the collector -> apo_corpus_upload_json -> device corpus: same records as the
    Python encoder, and the report over them equals the oracle's report of those records."""
    tcmod = import_module("senweaver-ide_b200.trace_collector")
    store = {}
    tc = tcmod.TraceCollectorService(engine, storageService=store)
    drive(tc, random.Random(21), n_threads=120)
    tc._dirty = True
    tc._saveToStorage()
    text = store["senweaver.traceCollector.data"]
    n = engine.corpus_upload_json(text)
    traces = tc.getAllTraces()
    assert n == len(traces)
    want = tc.corpus_records(traces)                        # live counters, valid = finalReward !== null
    got = engine.corpus_download(0, n)
    assert got.tobytes() == want.tobytes()
    engine.dims_upload(np.full((1, 4, 9), np.nan, np.float32))
    rep = engine.score(1, 0, corpus=True).report
    ref = orc.report(want)
    assert (rep.total, rep.good, rep.bad, rep.none, rep.withReward) == (ref.total, ref.good, ref.bad, ref.none, ref.withReward)
    assert rep.avgReward == pytest.approx(ref.avgReward, rel=1e-12)
    for p in range(6):
        assert (rep.pat[p].count, rep.pat[p].flag, list(rep.pat[p].examples)) == (ref.pat[p].count, ref.pat[p].flag, list(ref.pat[p].examples))
    with pytest.raises(import_module("senweaver-ide_b200").ApoError) as ei:
        engine.corpus_upload_json(text[:-5])
    assert "malformed trace JSON at byte" in str(ei.value)


def test_upload_payload_matches_transcription(engine):
    """uploadToServer (TCS:797-898): incremental selection, v2.0.0 body with rewardSummary / toolCallSummary from the
    engine, uploaded ids persisted under the reference's key, HTTP status handling and messages."""
    tcmod = import_module("senweaver-ide_b200.trace_collector")
    sent, store = [], {}

    def request(url, payload):
        sent.append((url, payload))
        return {"statusCode": 200}
    tc = tcmod.TraceCollectorService(engine, storageService=store, productService=None, requestService=request)
    assert tc.uploadToServer() == {"success": True, "message": "No new traces to upload", "uploadedCount": 0}
    drive(tc, random.Random(33), n_threads=80)
    res = tc.uploadToServer()
    assert res == {"success": True, "message": "Upload successful", "uploadedCount": 80}
    url, pay = sent[-1]
    assert url == "https://ide-api.senweaver.com/api/traces" and pay["version"] == "2.0.0" and len(pay["traces"]) == 80
    ref = ts.upload_summaries(pay["traces"])
    assert pay["toolCallSummary"] == ref["toolCallSummary"]
    rs, rr = pay["rewardSummary"], ref["rewardSummary"]
    assert rs["totalTracesWithReward"] == rr["totalTracesWithReward"] and set(rs["rewardDimensionAvg"]) == set(rr["rewardDimensionAvg"])
    assert rs["avgFinalReward"] == pytest.approx(rr["avgFinalReward"], rel=1e-12)
    for k, v in rr["rewardDimensionAvg"].items():
        assert rs["rewardDimensionAvg"][k] == pytest.approx(v, rel=1e-12, abs=1e-15)
    threads = {t["threadId"] for t in pay["traces"]}
    assert all(k.split(":")[0] in threads for k in pay["feedbacks"]) and len(pay["feedbacks"]) > 0
    assert sorted(json.loads(store["senweaver.traceCollector.uploadedIds"])) == sorted(t["id"] for t in pay["traces"])
    # nothing new -> nothing sent; a new trace -> only that one; a 5xx leaves it pending
    assert tc.uploadToServer()["uploadedCount"] == 0 and len(sent) == 1
    tc.startTrace("later")
    tc._request = lambda url, payload: {"statusCode": 503}
    assert tc.uploadToServer() == {"success": False, "message": "Server returned 503", "uploadedCount": 0}
    tc._request = request
    assert tc.uploadToServer()["uploadedCount"] == 1 and len(sent[-1][1]["traces"]) == 1
    tc._request = lambda url, payload: (_ for _ in ()).throw(OSError("offline"))
    tc.startTrace("later2")
    assert tc.uploadToServer() == {"success": False, "message": "Upload failed: offline", "uploadedCount": 0}
    # upload config persists under the reference's key
    tc.setAutoUploadConfig({"enabled": True, "intervalMs": 1000})
    tc2 = tcmod.TraceCollectorService(engine, storageService=store)
    assert tc2.getAutoUploadConfig() == {"enabled": True, "intervalMs": 1000, "traceApiUrl": "https://ide-api.senweaver.com/api/traces"}
    tc2.startTrace("fresh")
    assert tc2.uploadToServer()["message"].startswith("Upload failed")      # no request service: reported as a failure, trace stays pending


def test_optimize_payload_carries_the_gradient_prompt(services):
    _, tc, apo = services
    drive(tc, random.Random(8), n_threads=60)
    pay = apo.buildOptimizePayload()
    assert pay["textualGradientPrompt"].count("--- Experiment ") == min(apo.getConfig()["gradientBatchSize"], len(pay["rolloutResults"]))
    assert "## Current Prompt Rules\n(No optimized prompt rules currently active)" in pay["textualGradientPrompt"]
    live_s = sum(r["toolCallStats"]["succeeded"] for r in apo._convertTracesToRolloutResults(
        sorted([t for t in tc.getAllTraces() if t["summary"]["userFeedback"] is not None], key=lambda t: -t["startTime"])[:16]))
    assert pay["toolCallSummary"]["totalSucceeded"] == live_s


def test_apo_state_persists_and_gradient_flow(engine):
    """APO storage keys (APO:273-275,360,364), suggestion modify/revert against an existing segment (APO:1375-1458), the
    textual-gradient request (APO:1268-1343) and the auto-analyze gate (APO:453-475)."""
    pkg = import_module("senweaver-ide_b200")
    tcmod, apomod = import_module("senweaver-ide_b200.trace_collector"), import_module("senweaver-ide_b200.apo_service")
    store, sent = {}, []

    def request(url, payload):
        sent.append((url, payload))
        if url.endswith("/gradient"):
            return {"statusCode": 200, "critique": "Rules lack a stop condition. " * 8, "editedPrompt": "- stop after two failed tool calls"}
        return {"statusCode": 200}
    tc = tcmod.TraceCollectorService(engine, storageService=store)
    apo = apomod.APOService(engine, tc, storageService=store, requestService=request)
    assert apo._tryAutoAnalyze() is None                               # too few traces
    drive(tc, random.Random(77), n_threads=60)
    rep = apo._tryAutoAnalyze()
    assert rep is not None and sent[0][0].endswith("/api/apo/report") and sent[0][1]["version"] == "1.0.0"
    assert apo._tryAutoAnalyze() is None                               # interval not elapsed
    grads = apo.getTextualGradients()
    if rep["goodRate"] < 0.7:
        assert len(grads) == 1 and grads[0]["promptVersion"] == "v0" and grads[0]["rolloutSummary"].startswith("Based on 4 rollouts, avg reward: ")
        url, pay = [x for x in sent if x[0].endswith("/gradient")][0]
        assert pay["action"] == "textual_gradient" and len(pay["rolloutResults"]) == 4 and "{{critique_placeholder}}" in pay["applyEditPrompt"]
    else:
        apo.requestTextualGradient()
    sug = [s for s in apo.getPendingSuggestions() if s["description"].startswith("Textual Gradient: ")][0]
    assert sug["description"].endswith("...") and len(sug["description"]) == len("Textual Gradient: ") + 103
    # modify with no segment of that category: status only (APO:1388-1406); then against an existing segment
    apo.applySuggestion(sug["id"])
    assert apo.getOptimizedRules() == [] and apo.getStats()["appliedSuggestions"] == 1
    apo._segments.append({"id": "seg-1", "category": "core_behavior", "content": "be brief", "isActive": True, "isOptimized": False,
                          "version": 1, "createdAt": 0, "updatedAt": 0})
    s2 = dict(sug, id="s2", status="pending")
    apo._suggestions.append(s2)
    apo.applySuggestion("s2")
    assert apo.getOptimizedRules() == ["- stop after two failed tool calls"] and apo._segments[0]["originalContent"] == "be brief"
    assert apo.getOptimizedPromptForCategory("core_behavior") == "- stop after two failed tool calls"
    apo.revertSuggestion("s2")
    assert apo.getOptimizedRules() == [] and apo._segments[0]["content"] == "be brief" and apo._segments[0]["version"] == 3
    # beam state + everything else survives a restart; -Infinity is stored as null and compares like 0 afterwards
    apo._applyBeamUpdate({"beam": [{"version": "v1", "content": "- a", "score": -0.2}], "round": 1})
    apo.setConfig({"beamWidth": 8})
    apo.dispose()
    assert {"senweaver.apo.data", "senweaver.apo.config", "senweaver.apo.segments", "senweaver.apo.beamState", "senweaver.apo.gradients"} <= set(store)
    apo2 = apomod.APOService(engine, tc, storageService=store)
    assert apo2.getConfig()["beamWidth"] == 8 and apo2.getStats()["totalReports"] == apo.getStats()["totalReports"]
    assert apo2.getBeamState()["historyBestScore"] is None and apo2.getStats()["beamSearchActive"] is True
    assert len(apo2.getTextualGradients()) == len(apo.getTextualGradients())
    apo2._applyBeamUpdate({"bestPrompt": {"version": "v2", "content": "- b", "score": -0.1}, "bestScore": -0.1, "round": 2})
    assert apo2.getBeamState()["historyBestPrompt"] is None             # -0.1 > null is false in JS
    apo2._applyBeamUpdate({"bestPrompt": {"version": "v3", "content": "- c", "score": 0.1}, "bestScore": 0.1, "round": 3})
    assert apo2.getBeamState()["historyBestPrompt"]["version"] == "v3" and "c" in " ".join(apo2.getOptimizedRules())
    assert pkg is not None


def test_server_optimization_reply_is_merged(engine):
    """requestOptimizationFromServer (APO:992-1215): payload sent to {apo}/optimize, reply suggestions become pending,
    beamUpdate follows the strict-greater rule, textualGradient is stored with its edit suggestion; HTTP >= 400 -> []."""
    tcmod, apomod = import_module("senweaver-ide_b200.trace_collector"), import_module("senweaver-ide_b200.apo_service")
    sent, fired = [], []
    reply = {"statusCode": 200,
             "suggestions": [{"targetCategory": "tool_usage", "type": "add", "priority": "medium", "description": "d", "suggestedContent": "- check tool output",
                              "reasoning": "r", "estimatedImpact": "i"}],
             "beamUpdate": {"beam": [{"version": "v1", "content": "- x\n- y", "score": 0.4}], "bestPrompt": {"version": "v1", "content": "- x\n- y", "score": 0.4},
                            "bestScore": 0.4, "round": 1},
             "textualGradient": {"critique": "c" * 150, "editedPrompt": "- z"}}

    def request(url, payload):
        sent.append((url, payload))
        return reply if url.endswith("/optimize") else {"statusCode": 200}
    tc = tcmod.TraceCollectorService(engine)
    apo = apomod.APOService(engine, tc, requestService=request)
    apo.onDidGenerateSuggestions(lambda s: fired.append(s))
    drive(tc, random.Random(4), n_threads=40)
    out = apo.requestOptimizationFromServer()
    url, pay = [x for x in sent if x[0].endswith("/optimize")][0]
    assert url == "https://ide-api.senweaver.com/api/apo/optimize" and pay["version"] == "2.0.0" and "textualGradientPrompt" in pay
    assert len(out) == 2 and all(s["status"] == "pending" and s["id"] for s in out) and fired[-1] == out
    assert out[1]["description"] == "Textual Gradient optimization: " + "c" * 100 + "..." and out[1]["promptVersion"] == "v1"
    st = apo.getBeamState()
    assert st["currentRound"] == 1 and st["historyBestScore"] == 0.4 and apo.getOptimizedRules() == ["x", "y"]
    g = apo.getTextualGradients()
    n_rated = min(16, sum(1 for t in tc.getAllTraces() if t["summary"]["userFeedback"] is not None))
    assert len(g) == 1 and g[0]["rolloutSummary"] == f"Based on {n_rated} rollouts"
    # an equal score does not replace the incumbent; an HTTP error yields [] and changes nothing
    reply["beamUpdate"] = {"bestPrompt": {"version": "v2", "content": "- q", "score": 0.4}, "bestScore": 0.4, "round": 2}
    reply["suggestions"], reply["textualGradient"] = [], None
    assert apo.requestOptimizationFromServer() == [] and apo.getBeamState()["historyBestPrompt"]["version"] == "v1"
    reply["statusCode"] = 500
    before = apo.getStats()
    assert apo.requestOptimizationFromServer() == [] and apo.getStats()["totalSuggestions"] == before["totalSuggestions"]
