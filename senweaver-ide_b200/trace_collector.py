"""Host-side mirror of the reference's TraceCollectorService API (TCS:133-210).

Same method names, argument meaning and error behaviour as
src/vs/workbench/contrib/senweaver/common/traceCollectorService.ts, with the arithmetic
delegated to the B200 engine through the C ABI:

    _computeRewardSignals (TCS:668-788)      -> Engine.reward_batch   (k_reward_batch)
    getStats means / tool totals (TCS:596-626) -> Engine.score(corpus=True) (k_detect6)

Recorders never raise into the caller (TCS:438 `catch { /* silent */ }`); engine failures
leave the trace untouched.  The reference defers recorders with queueMicrotask; this mirror
runs them immediately (single-threaded host, same FIFO order).  Storage / HTTP upload are
injected collaborators exactly as in the reference constructor (storage, product, request)
and are optional here — the scoring engine itself is stateless.
"""
from __future__ import annotations

import json
import time
import uuid

import numpy as np

from .engine import DIM_NAMES, F_ENDED, F_ERRORS, F_FAILSPAN, F_VALID, RECORD_DTYPE, Engine

MAX_CONTENT_PREVIEW = 500          # TCS:218
MAX_TRACES = 1000                  # TCS:219
MAX_SPANS_PER_TRACE = 200          # TCS:220
TRACE_STORAGE_KEY = "senweaver.traceCollector.data"          # TCS:216
TRACE_FEEDBACK_KEY = "senweaver.traceCollector.feedbacks"    # TCS:217
UPLOADED_IDS_KEY = "senweaver.traceCollector.uploadedIds"     # TCS:795
UPLOAD_CONFIG_KEY = "senweaver.traceCollector.uploadConfig"   # TCS:317
TRACE_SCORED_KEY = "senweaver.traceCollector.scoredRecords"  # engine-side: hex Form R snapshot per scored trace
MODE_CODE = {"normal": 1, "agent": 2, "gather": 3, "designer": 4}
FB_CODE = {None: 0, "good": 1, "bad": 2}
U32 = 0xFFFFFFFF


def duration_class(dur: float, total_tool_calls: int) -> int:
    """apo_record.durClass (include/apo_b200.h APO_DC_*): the comparisons TCS:721-728 / APO:754 make on the binary64
    totalToolDurationMs, stored because the record keeps only a binary32 copy of the value (apo_duration_class in C)."""
    dc = 0x80
    if dur > 0:
        dc |= 0x04
    if dur > 15000:
        dc |= 0x08
    if total_tool_calls > 0 and dur > 0:
        avg = dur / total_tool_calls
        dc |= (avg > 1000) + (avg > 3000) + (avg > 10000)
    return dc


def encode_trace(trace: dict, valid: bool | None = None) -> np.ndarray:
    """ConversationTrace -> Form R (include/apo_b200.h apo_record): the fields the scoring
    path reads (TCS:94-108, :87, :91) plus the span-derived counts of TCS:752-753, APO:667-669."""
    s = trace["summary"]
    md = trace.get("metadata") or {}
    spans = trace["spans"]
    user = sum(1 for sp in spans if sp["type"] == "user_message")
    asst = sum(1 for sp in spans if sp["type"] == "assistant_message")
    failspan = any(sp["type"] == "tool_call" and sp["data"].get("toolSuccess") is False for sp in spans)
    if valid is None:
        valid = s["finalReward"] is not None
    rec = np.zeros(1, RECORD_DTYPE)
    rec["feedback"] = FB_CODE.get(s["userFeedback"], 0)
    rec["flags"] = ((F_ERRORS if s["hasErrors"] else 0) | (F_ENDED if trace.get("endTime") else 0) |
                    (F_VALID if valid else 0) | (F_FAILSPAN if failspan else 0))
    rec["mode"] = MODE_CODE.get(md.get("chatMode") or "", 0)
    rec["userMsgs"], rec["asstMsgs"] = min(user, 65535), min(asst, 65535)
    rec["toolCalls"] = min(int(s["totalToolCalls"]), U32)
    rec["toolSucc"] = min(int(s["toolCallsSucceeded"]), U32)
    rec["toolFail"] = min(int(s["toolCallsFailed"]), U32)
    rec["llmCalls"] = min(int(s["totalLLMCalls"]), U32)
    rec["tokens"] = min(int(s["totalTokens"]), U32)
    rec["toolDurMs"] = float(s["totalToolDurationMs"])
    rec["durClass"] = duration_class(float(s["totalToolDurationMs"]), int(s["totalToolCalls"]))
    return rec


class TraceCollectorService:
    def __init__(self, engine: Engine, storageService=None, productService=None, requestService=None):
        self._engine = engine
        self._storage, self._product, self._request = storageService, productService, requestService
        self._listeners = []
        self._traces: dict[str, dict] = {}
        self._active: dict[str, str] = {}
        self._feedbacks: dict[str, str | None] = {}
        self._scored: dict[str, np.ndarray] = {}       # Form R snapshot taken when the reward was computed
        self._dirty = False
        api = (getattr(productService, "senweaverApiConfig", None) or {}).get("apiBaseUrl") if productService else None
        self._traceApiUrl = f"{api or 'https://ide-api.senweaver.com'}/api/traces"     # TCS:245-246
        self._autoUploadConfig = {"enabled": False, "intervalMs": 300000}              # TCS:792
        self._uploadedIds: set[str] = set()
        self._loadFromStorage()
        self._loadUploadConfig()
        self._loadUploadedIds()

    # ---- events
    def onDidChangeState(self, listener):
        self._listeners.append(listener)

    def _fire(self):
        for fn in list(self._listeners):
            try:
                fn()
            except Exception:
                pass

    # ---- storage (TCS:296-359)
    def _loadFromStorage(self):
        if not self._storage:
            return
        try:
            for t in json.loads(self._storage.get(TRACE_STORAGE_KEY, "[]")):
                self._traces[t["id"]] = t
            self._feedbacks.update(json.loads(self._storage.get(TRACE_FEEDBACK_KEY, "{}")))
            # the Form R snapshot each stored finalReward was computed from (engine-side key, not a reference key):
            # without it a reloaded trace would be re-encoded from counters that kept moving after the reward was set
            for tid, hx in json.loads(self._storage.get(TRACE_SCORED_KEY, "{}")).items():
                if tid in self._traces:
                    self._scored[tid] = np.frombuffer(bytes.fromhex(hx), dtype=RECORD_DTYPE).copy()
        except Exception as e:                      # TCS:311 warn and continue
            print("[TraceCollector] Failed to load from storage:", e)

    def _saveToStorage(self):
        if not self._dirty:
            return
        try:
            if len(self._traces) > MAX_TRACES:      # keep the newest MAX_TRACES by startTime (TCS:337-345)
                keep = sorted(self._traces.values(), key=lambda t: -(t.get("startTime") or 0))[:MAX_TRACES]
                self._traces = {t["id"]: t for t in keep}
                self._scored = {k: v for k, v in self._scored.items() if k in self._traces}
            if self._storage is not None:
                self._storage[TRACE_STORAGE_KEY] = json.dumps(list(self._traces.values()))
                self._storage[TRACE_FEEDBACK_KEY] = json.dumps(self._feedbacks)
                self._storage[TRACE_SCORED_KEY] = json.dumps({k: v.tobytes().hex() for k, v in self._scored.items()})
            self._dirty = False
        except Exception as e:
            print("[TraceCollector] Failed to save to storage:", e)

    # ---- helpers
    @staticmethod
    def _truncate(s, n=MAX_CONTENT_PREVIEW):
        if not s:
            return ""
        return s[:n] + "..." if len(s) > n else s

    def _getOrCreateTrace(self, threadId):
        tid = self._active.get(threadId)
        if tid and tid in self._traces:
            return self._traces[tid]
        return self._traces[self.startTrace(threadId)]          # auto-created: no metadata (TCS:264-272)

    def _addSpan(self, trace, type_, messageIdx, data, duration=None):
        if len(trace["spans"]) >= MAX_SPANS_PER_TRACE:          # TCS:275-277: counters still advance
            return
        span = {"id": str(uuid.uuid4()), "traceId": trace["id"], "threadId": trace["threadId"], "messageIdx": messageIdx,
                "type": type_, "timestamp": time.time() * 1000.0, "data": data}
        if duration is not None:
            span["duration"] = duration
        trace["spans"].append(span)
        self._dirty = True

    # ---- lifecycle (TCS:380-425)
    def startTrace(self, threadId, metadata=None):
        tid = str(uuid.uuid4())
        self._traces[tid] = {
            "id": tid, "threadId": threadId, "startTime": time.time() * 1000.0, "spans": [],
            "summary": {"totalLLMCalls": 0, "totalToolCalls": 0, "totalTokens": 0, "userFeedback": None, "hasErrors": False,
                        "toolCallsSucceeded": 0, "toolCallsFailed": 0, "toolCallsByName": {}, "totalToolDurationMs": 0,
                        "finalReward": None, "rewardDimensions": []},
            "metadata": metadata}
        self._active[threadId] = tid
        self._dirty = True
        return tid

    def endTrace(self, traceId):
        t = self._traces.get(traceId)
        if t:
            t["endTime"] = time.time() * 1000.0
            self._computeRewardSignals(t)
            self._dirty = True
            self._saveToStorage()

    def endTraceForThread(self, threadId):
        tid = self._active.get(threadId)
        if tid:
            self.endTrace(tid)

    # ---- recorders (TCS:429-569): fire-and-forget, never raise
    def recordUserMessage(self, threadId, messageIdx, content):
        try:
            t = self._getOrCreateTrace(threadId)
            self._addSpan(t, "user_message", messageIdx, {"contentPreview": self._truncate(content), "contentLength": len(content)})
        except Exception:
            pass

    def recordAssistantMessage(self, threadId, messageIdx, content, model=None, provider=None):
        try:
            t = self._getOrCreateTrace(threadId)
            self._addSpan(t, "assistant_message", messageIdx, {"contentPreview": self._truncate(content),
                                                               "contentLength": len(content), "model": model, "provider": provider})
        except Exception:
            pass

    def recordLLMCall(self, threadId, messageIdx, data):
        try:
            t = self._getOrCreateTrace(threadId)
            self._addSpan(t, "llm_call", messageIdx, {k: data.get(k) for k in ("model", "provider", "inputTokens", "outputTokens", "temperature")},
                          duration=data.get("duration"))
            t["summary"]["totalLLMCalls"] += 1
            t["summary"]["totalTokens"] += (data.get("inputTokens") or 0) + (data.get("outputTokens") or 0)
        except Exception:
            pass

    def recordToolCall(self, threadId, messageIdx, data):
        try:
            t = self._getOrCreateTrace(threadId)
            self._addSpan(t, "tool_call", messageIdx, {"toolName": data["toolName"], "toolParams": self._truncate(data.get("toolParams")),
                                                       "toolResult": self._truncate(data.get("toolResult")),
                                                       "toolSuccess": bool(data["toolSuccess"])}, duration=data.get("duration"))
            s = t["summary"]
            s["totalToolCalls"] += 1
            key = "toolCallsSucceeded" if data["toolSuccess"] else "toolCallsFailed"
            s[key] += 1
            by = s["toolCallsByName"].setdefault(data["toolName"], {"total": 0, "succeeded": 0, "failed": 0})
            by["total"] += 1
            by["succeeded" if data["toolSuccess"] else "failed"] += 1
            if data.get("duration") and data["duration"] > 0:
                s["totalToolDurationMs"] += data["duration"]
            self._dirty = True
        except Exception:
            pass

    def recordUserFeedback(self, threadId, messageIdx, feedback):
        try:
            self._feedbacks[f"{threadId}:{messageIdx}"] = feedback
            t = self._getOrCreateTrace(threadId)             # the thread's *active* trace (TCS:538)
            self._addSpan(t, "user_feedback", messageIdx, {"feedback": feedback})
            t["summary"]["userFeedback"] = feedback
            self._dirty = True
            self._computeRewardSignals(t)                     # TCS:547
            self._fire()
            self._saveToStorage()
        except Exception:
            pass

    def recordError(self, threadId, messageIdx, errorMessage):
        try:
            t = self._getOrCreateTrace(threadId)
            self._addSpan(t, "error", messageIdx, {"errorMessage": self._truncate(errorMessage, 1000)})
            t["summary"]["hasErrors"] = True
        except Exception:
            pass

    # ---- the hot step: TCS:668-788 on the GPU
    def _computeRewardSignals(self, trace):
        rec = encode_trace(trace, valid=True)
        dims, masks, finals = self._engine.reward_batch(rec)
        m = int(masks[0])
        trace["summary"]["rewardDimensions"] = [{"name": DIM_NAMES[i], "value": float(dims[0, i])} for i in range(9) if m & (1 << i)]
        trace["summary"]["finalReward"] = float(finals[0])
        self._scored[trace["id"]] = rec

    # ---- queries
    def getFeedback(self, threadId, messageIdx):
        return self._feedbacks.get(f"{threadId}:{messageIdx}")

    def getAllTraces(self):
        return list(self._traces.values())

    def corpus_records(self, traces=None, scored: bool = False) -> np.ndarray:
        """Form R for a list of traces.  scored=True returns the snapshot taken when each reward was
        computed (late events mutate counters without re-scoring, TCS:420-425)."""
        traces = self.getAllTraces() if traces is None else traces
        out = np.zeros(len(traces), RECORD_DTYPE)
        for i, t in enumerate(traces):
            snap = self._scored.get(t["id"]) if scored else None
            if snap is not None and t["summary"]["finalReward"] is not None:
                out[i] = snap[0]
            else:
                out[i] = encode_trace(t)[0]
        return out

    def getStats(self):
        traces = self.getAllTraces()
        total_spans = sum(len(t["spans"]) for t in traces)
        starts = [t["startTime"] for t in traces]
        good = sum(1 for v in self._feedbacks.values() if v == "good")
        bad = sum(1 for v in self._feedbacks.values() if v == "bad")
        tool = (0, 0, 0)
        avg, with_reward, rate = None, 0, None
        if traces:
            # tool totals from the live counters, reward mean from the scored snapshots (TCS:602-610)
            self._engine.corpus_upload(self.corpus_records(traces))
            self._engine.dims_upload(np.full((1, 4, 9), np.nan, np.float32))
            rep = self._engine.score(1, 0, corpus=True).report
            tool = (int(rep.toolCalls), int(rep.toolSucc), int(rep.toolFail))
            rate = None if rep.toolCalls == 0 else float(rep.toolSuccessRate)
            vals = [t["summary"]["finalReward"] for t in traces if t["summary"]["finalReward"] is not None]
            with_reward = len(vals)
            if vals:
                self._engine.corpus_upload(self.corpus_records(traces, scored=True))
                rep2 = self._engine.score(1, 0, corpus=True).report
                avg = float(rep2.avgReward)
        return {"totalTraces": len(traces), "totalSpans": total_spans, "totalFeedbacks": good + bad, "goodFeedbacks": good,
                "badFeedbacks": bad, "storageUsedBytes": self._estimateStorageBytes(),
                "oldestTraceTime": min(starts) if starts else None, "newestTraceTime": max(starts) if starts else None,
                "totalToolCalls": tool[0], "totalToolSucceeded": tool[1], "totalToolFailed": tool[2], "toolSuccessRate": rate,
                "avgFinalReward": avg, "tracesWithReward": with_reward}

    def _estimateStorageBytes(self):
        try:
            return len(json.dumps(list(self._traces.values()))) + len(json.dumps(self._feedbacks))
        except Exception:
            return 0

    def exportData(self):
        return json.dumps({"version": "1.0.0", "exportTime": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
                           "stats": self.getStats(), "traces": self.getAllTraces(), "feedbacks": self._feedbacks}, indent=2)

    def clearAllData(self):
        self._traces.clear(); self._active.clear(); self._feedbacks.clear(); self._scored.clear()
        self._dirty = True
        self._saveToStorage()
        self._fire()

    # ---- backend upload (TCS:797-898): payload assembly; HTTP goes through the injected request service
    def uploadToServer(self):
        try:
            new = [t for t in self._traces.values() if t["id"] not in self._uploadedIds]     # incremental (TCS:800)
            if not new:
                return {"success": True, "message": "No new traces to upload", "uploadedCount": 0}
            payload = self.buildUploadPayload(new)
            if self._request is None:
                raise RuntimeError("no request service configured")
            resp = self._request(self._traceApiUrl, payload)
            status = resp.get("statusCode") if isinstance(resp, dict) else (resp if isinstance(resp, int) and not isinstance(resp, bool) else (200 if resp else 500))
            if status and status >= 400:                                                     # TCS:881-883
                return {"success": False, "message": f"Server returned {status}", "uploadedCount": 0}
            self._uploadedIds.update(t["id"] for t in new)
            self._saveUploadedIds()
            return {"success": True, "message": "Upload successful", "uploadedCount": len(new)}
        except Exception as e:                          # TCS:892-896
            print("[TraceCollector] Upload failed:", e)
            return {"success": False, "message": f"Upload failed: {e}", "uploadedCount": 0}

    def buildUploadPayload(self, new):
        """The v2.0.0 body of TCS:846-873.  rewardSummary and the tool totals are reductions of the same records the
        scoring path reads and run on the engine (reward from the records the rewards were computed from, tool totals
        from the live counters, TCS:813-818); the string-keyed byToolName table and the duration total stay on the host."""
        nan_dims = np.full((1, 4, 9), np.nan, np.float32)
        self._engine.corpus_upload(self.corpus_records(new))
        self._engine.dims_upload(nan_dims)
        live = self._engine.score(1, 0, corpus=True).report
        self._engine.corpus_upload(self.corpus_records(new, scored=True))
        rep = self._engine.score(1, 0, corpus=True).report
        by_name, duration = {}, 0
        for t in new:
            duration += t["summary"]["totalToolDurationMs"]
            for name, st in t["summary"]["toolCallsByName"].items():
                g = by_name.setdefault(name, {"total": 0, "succeeded": 0, "failed": 0})
                for k in g:
                    g[k] += st[k]
        threads = {t["threadId"] for t in new}
        succ, fail = int(live.toolSucc), int(live.toolFail)
        return {
            "version": "2.0.0",
            "uploadTime": time.strftime("%Y-%m-%dT%H:%M:%S.000Z", time.gmtime()),
            "traces": new,
            "feedbacks": {k: v for k, v in self._feedbacks.items() if k.split(":")[0] in threads},
            "rewardSummary": {
                "totalTracesWithReward": int(rep.withReward),
                "avgFinalReward": None if rep.withReward == 0 else float(rep.avgReward),
                "rewardDimensionAvg": {DIM_NAMES[i]: float(rep.dim[i].avg) for i in range(9) if rep.dim[i].count},
            },
            "toolCallSummary": {
                "totalToolCalls": succ + fail, "totalSucceeded": succ, "totalFailed": fail,
                "successRate": succ / (succ + fail) if succ + fail > 0 else None,
                "totalDurationMs": duration, "byToolName": by_name,
            },
        }

    def _loadUploadedIds(self):                          # TCS:939-947
        try:
            js = self._storage.get(UPLOADED_IDS_KEY) if self._storage else None
            if js:
                self._uploadedIds = set(json.loads(js))
        except Exception:
            pass

    def _saveUploadedIds(self):                          # TCS:949-961: only ids that still exist
        try:
            self._uploadedIds = {i for i in self._uploadedIds if i in self._traces}
            if self._storage is not None:
                self._storage[UPLOADED_IDS_KEY] = json.dumps(sorted(self._uploadedIds))
        except Exception:
            pass

    def _loadUploadConfig(self):                         # TCS:314-330
        try:
            js = self._storage.get(UPLOAD_CONFIG_KEY) if self._storage else None
            if js:
                cfg = json.loads(js)
                iv = cfg.get("intervalMs")
                self._autoUploadConfig = {"enabled": bool(cfg.get("enabled")), "intervalMs": 300000 if iv is None else iv}
        except Exception:
            pass

    def setAutoUploadConfig(self, config):               # TCS:900-933 (the interval timer itself belongs to the host application)
        iv = config.get("intervalMs")
        self._autoUploadConfig = {"enabled": bool(config.get("enabled")), "intervalMs": 300000 if iv is None else iv}
        try:
            if self._storage is not None:
                self._storage[UPLOAD_CONFIG_KEY] = json.dumps(self._autoUploadConfig)
        except Exception:
            pass

    def getAutoUploadConfig(self):
        return {**self._autoUploadConfig, "traceApiUrl": self._traceApiUrl}
