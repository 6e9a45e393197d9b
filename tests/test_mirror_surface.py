"""The host mirrors expose the reference's service surface: ITraceCollectorService — 1 event + 17 methods
(traceCollectorService.ts:133-210), IAPOService — 2 events + 16 methods (apoService.ts:203-267) — under the same
names, plus the storage keys and caps a maintainer would diff against (TCS:216-220, APO:273-277,360,364,405).
Pure introspection: no engine, no GPU."""
import inspect
from importlib import import_module

tcmod = import_module("senweaver-ide_b200.trace_collector")
apomod = import_module("senweaver-ide_b200.apo_service")

TRACE_COLLECTOR = {
    "onDidChangeState": 1, "startTrace": 2, "endTrace": 1, "endTraceForThread": 1, "recordUserMessage": 3, "recordAssistantMessage": 5,
    "recordLLMCall": 3, "recordToolCall": 3, "recordUserFeedback": 3, "recordError": 3, "getFeedback": 2, "getStats": 0, "getAllTraces": 0,
    "exportData": 0, "clearAllData": 0, "uploadToServer": 0, "setAutoUploadConfig": 1, "getAutoUploadConfig": 0,
}
APO = {
    "onDidChangeState": 1, "onDidGenerateSuggestions": 1, "analyzePromptEffectiveness": 0, "requestOptimizationFromServer": 0,
    "requestTextualGradient": 0, "getActiveSegments": 0, "getOptimizedPromptForCategory": 1, "getOptimizedRules": 0, "applySuggestion": 1,
    "rejectSuggestion": 1, "revertSuggestion": 1, "getLatestReport": 0, "getPendingSuggestions": 0, "getStats": 0, "getConfig": 0,
    "setConfig": 1, "getBeamState": 0, "getTextualGradients": 1,
}


def check(cls, surface):
    for name, nargs in surface.items():
        fn = getattr(cls, name, None)
        assert callable(fn), f"{cls.__name__}.{name} missing"
        params = [p for p in list(inspect.signature(fn).parameters.values())[1:] if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        assert len(params) >= nargs and sum(1 for p in params if p.default is p.empty) <= nargs, (name, params)


def test_trace_collector_surface():
    assert len(TRACE_COLLECTOR) == 18                     # 1 event + 17 methods
    check(tcmod.TraceCollectorService, TRACE_COLLECTOR)
    assert list(inspect.signature(tcmod.TraceCollectorService.__init__).parameters)[1:] == ["engine", "storageService", "productService", "requestService"]


def test_apo_surface():
    assert len(APO) == 18                                 # 2 events + 16 methods
    check(apomod.APOService, APO)
    assert list(inspect.signature(apomod.APOService.__init__).parameters)[1:] == ["engine", "traceCollectorService", "storageService", "productService", "requestService"]


def test_storage_keys_and_caps():
    assert (tcmod.TRACE_STORAGE_KEY, tcmod.TRACE_FEEDBACK_KEY) == ("senweaver.traceCollector.data", "senweaver.traceCollector.feedbacks")
    assert (tcmod.UPLOADED_IDS_KEY, tcmod.UPLOAD_CONFIG_KEY) == ("senweaver.traceCollector.uploadedIds", "senweaver.traceCollector.uploadConfig")
    assert (tcmod.MAX_CONTENT_PREVIEW, tcmod.MAX_TRACES, tcmod.MAX_SPANS_PER_TRACE) == (500, 1000, 200)
    assert (apomod.APO_STORAGE_KEY, apomod.APO_CONFIG_KEY, apomod.APO_SEGMENTS_KEY) == ("senweaver.apo.data", "senweaver.apo.config", "senweaver.apo.segments")
    assert (apomod.APO_BEAM_KEY, apomod.APO_GRADIENTS_KEY) == ("senweaver.apo.beamState", "senweaver.apo.gradients")
    assert (apomod.MAX_REPORTS, apomod.MAX_SUGGESTIONS, apomod.MAX_GRADIENTS) == (50, 200, 50)
    d = apomod.DEFAULT_APO_CONFIG
    assert (d["beamWidth"], d["branchFactor"], d["beamRounds"], d["gradientBatchSize"]) == (4, 4, 3, 4)      # APO:288-291
    assert (d["autoAnalyzeIntervalMs"], d["minTracesForAnalysis"], d["minFeedbacksForAnalysis"]) == (3600000, 20, 10)
