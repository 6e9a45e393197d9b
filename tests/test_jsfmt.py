"""JS number / string formatting used by the wire formats (toFixed ties, UTF-16 lengths) and the experiment block of
the textual-gradient prompt (APO:926-941).  Known answers are what V8 prints for the same expressions."""
from importlib import import_module

import pytest

jsfmt = import_module("senweaver-ide_b200.jsfmt")
apo = import_module("senweaver-ide_b200.apo_service")


@pytest.mark.parametrize("x,d,want", [
    (0.125, 2, "0.13"), (0.375, 2, "0.38"), (2.5, 0, "3"), (-2.5, 0, "-3"), (1.005, 2, "1.00"), (1.45, 1, "1.4"), (8.345, 2, "8.35"),
    (-0.0004, 3, "-0.000"), (-0.0, 2, "0.00"), (0.0, 0, "0"), (0.639, 3, "0.639"), (-0.43600000000000005, 3, "-0.436"),
    (0.20930232558139533, 3, "0.209"), (-0.20000000000000018, 2, "-0.20"), (25.0, 0, "25"), (99.5, 0, "100"), (1e20, 0, "100000000000000000000"),
    (float("nan"), 2, "NaN"), (float("inf"), 1, "Infinity"), (123456789.987654321, 3, "123456789.988")])
def test_to_fixed_known_answers(x, d, want):
    assert jsfmt.js_to_fixed(x, d) == want


def test_utf16_units():
    assert jsfmt.js_length("a\U0001F600") == 3
    assert jsfmt.js_substring("\U0001F600" * 150, 0, 200) == "\U0001F600" * 100
    assert jsfmt.js_substring("abc", 0, 200) == "abc" and jsfmt.js_substring("héllo", 1, 3) == "él"


def test_experiment_block_known_answer():
    r = {"status": "failed", "finalReward": -0.43600000000000005, "chatMode": "normal",
         "messages": [{"role": "user", "content": "x" * 300}, {"role": "tool", "content": "ok"}],
         "toolCallStats": {"totalCalls": 4, "succeeded": 1, "failed": 3, "successRate": 0.25, "totalDurationMs": 50000.5, "byToolName": {}},
         "rewardDimensions": [{"name": "user_feedback", "value": -1}, {"name": "response_efficiency", "value": -0.20000000000000018}],
         "llmStats": {"totalCalls": 4, "totalTokens": 12000}}
    want = ("--- Experiment 3 ---\nStatus: ❌ Failed\nFinal Reward: -0.436\nChat Mode: normal\n"
            "Tool Calls: 4 (1 succeeded, 3 failed, rate: 25%, duration: 50001ms)\nLLM Calls: 4, Tokens: 12000\n"
            "Reward Dims: user_feedback=-1.00, response_efficiency=-0.20\nMessages:\n    [user] " + "x" * 200 + "\n    [tool] ok")
    assert apo.APOService._experiment_block(2, r) == want
    r2 = dict(r, status="unknown", finalReward=None, rewardDimensions=[], toolCallStats=dict(r["toolCallStats"], totalCalls=0), messages=[])
    assert apo.APOService._experiment_block(0, r2) == ("--- Experiment 1 ---\nStatus: ❓ Unknown\nFinal Reward: N/A\nChat Mode: normal\n"
                                                       "Tool Calls: none\nLLM Calls: 4, Tokens: 12000\n\nMessages:\n    ")
