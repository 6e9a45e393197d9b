#!/usr/bin/env python
"""Generates ts/patches/*.patch — the edits a senweaver-ide maintainer applies to bind the B200 scoring engine.

    python ts/patches/make_patches.py [--reference /root/reference]

The committed patches are ed scripts (`diff -e`: line-number commands plus the NEW text only — no line of the reference is
copied into this repository); apply them with `patch --ed FILE SCRIPT` or `make_patches.py --apply SCRIPT FILE`, or regenerate
reviewable unified diffs with `--unified DIR`.  Each script names the sha256 of the file it was generated against.  Targets:

    src/vs/workbench/contrib/senweaver/common/traceCollectorService.ts   _computeRewardSignals, getStats delegate
    src/vs/workbench/contrib/senweaver/common/apoService.ts              _buildReport, _analyzePatterns delegate; evaluateBeam feeds APO:1153-1165
    src/vs/code/electron-main/app.ts                                     main-process service + IPC channel registration

They never touch the service interfaces (TCS:133-210, APO:203-267), the decorator ids (TCS:212, APO:269), the exported
types, the storage keys or the `registerSingleton(..., InstantiationType.Delayed)` lines: tests/test_ts_patches.py checks the
hunk ranges (always) and, when the reference checkout is present, applies the patches and compares those regions byte for
byte.  The new dependency is appended to each constructor, so the existing injection order (TCS:238-242, APO:321-326) stays.

Together with the three new files in ts/ (apoScoringService.ts, apoScoringMainService.ts, traceRecordCodec.ts) this is the
whole reference-side binding.  Nothing here can be compiled in the engine's build image (no tsc / node).
"""
from __future__ import annotations

import argparse
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TCS = "src/vs/workbench/contrib/senweaver/common/traceCollectorService.ts"
APO = "src/vs/workbench/contrib/senweaver/common/apoService.ts"
APP = "src/vs/code/electron-main/app.ts"


def replace_lines(lines, first, last, new_text):
    """Replace 1-based inclusive [first, last] by new_text (a string ending in a newline, or '' to delete)."""
    return lines[:first - 1] + new_text.splitlines(keepends=True) + lines[last:]


def insert_after(lines, lineno, new_text):
    return lines[:lineno] + new_text.splitlines(keepends=True) + lines[lineno:]


def find(lines, needle, start=1):
    for i in range(start - 1, len(lines)):
        if needle in lines[i]:
            return i + 1
    raise KeyError(needle)


def method_end(lines, first):
    """Line of the closing brace of the tab-indented member that starts at `first`."""
    for i in range(first, len(lines)):
        if lines[i].rstrip("\n") == "\t}":
            return i + 1
    raise KeyError("method end")


# ------------------------------------------------------------------------------------------------ traceCollectorService.ts
TCS_IMPORTS = """import { VSBuffer } from '../../../../base/common/buffer.js';
import { IApoScoringService } from './apoScoringService.js';
import { APO_RECORD_BYTES, CorpusReportNumbers, EMPTY_DIMS, decodeCorpusReport, decodeReward, encodeTraceRecord, encodeTraceRecords } from './traceRecordCodec.js';
"""

TCS_FIELDS = """
	// Corpus aggregates of the B200 scoring engine (tool totals, mean finalReward), refreshed after every scored trace
	private _engineStats: CorpusReportNumbers | null = null;
"""

TCS_CTOR_PARAM = """		@IApoScoringService private readonly _scoring: IApoScoringService,
"""

TCS_CTOR_TAIL = """		this._refreshEngineStats();
"""

TCS_GETSTATS_TAIL = """		// Tool-call and reward aggregates: one corpus pass of the scoring engine (apo_corpus_report: TCS.getStats and
		// APO._buildReport take the same sums), cached by _refreshEngineStats — getStats stays synchronous.
		const es = this._engineStats;
		const totalToolCalls = es ? es.toolCalls : 0;
		const totalToolSucceeded = es ? es.toolSucc : 0;
		const totalToolFailed = es ? es.toolFail : 0;
		const tracesWithReward = es ? es.withReward : 0;
"""

TCS_GETSTATS_RETURN_OLD = ("toolSuccessRate: totalToolCalls > 0 ? totalToolSucceeded / totalToolCalls : null,",
                           "avgFinalReward: tracesWithReward > 0 ? rewardSum / tracesWithReward : null,")
TCS_GETSTATS_RETURN_NEW = ("toolSuccessRate: es ? es.toolSuccessRate : null,",
                           "avgFinalReward: es ? es.avgReward : null,")

TCS_REWARD = """	private _computeRewardSignals(trace: ConversationTrace): void {
		// Delegated to the B200 scoring engine: apo_reward_batch computes the nine dimensions and the weighted mean of this
		// function bit for bit (tests/test_reference_pin.py).  The record is a snapshot of the summary taken now, as the
		// synchronous code did; the channel is async, so the two summary fields are assigned when the reply arrives.  On
		// any failure the trace is left untouched and a warning is logged (never throw into callers).
		const bytes = new Uint8Array(APO_RECORD_BYTES);
		encodeTraceRecord(trace, new DataView(bytes.buffer), 0, true);
		this._scoring.rewardBatch(VSBuffer.wrap(bytes)).then(r => {
			const { rewardDimensions, finalReward } = decodeReward(r.dims, r.masks, r.finals, 0);
			trace.summary.rewardDimensions = rewardDimensions;
			trace.summary.finalReward = finalReward;
			this._dirty = true;
			this._refreshEngineStats();
			queueMicrotask(() => this._saveToStorage());
			this._onDidChangeState.fire();
		}).catch(e => {
			console.warn('[TraceCollector] reward scoring failed:', e instanceof Error ? e.message : String(e));
		});
	}

	/** One engine pass over all traces (report-only call: a 1 x 4 all-absent dims block) -> cached aggregates for getStats. */
	private _refreshEngineStats(): void {
		const traces = Array.from(this._traces.values());
		if (traces.length === 0) {
			this._engineStats = null;
			return;
		}
		this._scoring.score(EMPTY_DIMS, 1, 4, encodeTraceRecords(traces), 0).then(blocks => {
			this._engineStats = decodeCorpusReport(blocks.report);
			this._onDidChangeState.fire();
		}).catch(e => {
			console.warn('[TraceCollector] stats refresh failed:', e instanceof Error ? e.message : String(e));
		});
	}
"""


def patch_tcs(lines):
    # bottom-up so that the line numbers of the reference stay valid
    a = find(lines, "private _computeRewardSignals(trace: ConversationTrace): void {")
    lines = replace_lines(lines, a, method_end(lines, a), TCS_REWARD)
    g = find(lines, "\tgetStats(): TraceCollectorStats {")
    b = find(lines, "// Aggregate tool call and reward statistics", g)
    r = find(lines, "\t\treturn {", b)
    lines = replace_lines(lines, b, r - 1, TCS_GETSTATS_TAIL + "\n")
    for old, new in zip(TCS_GETSTATS_RETURN_OLD, TCS_GETSTATS_RETURN_NEW):
        k = find(lines, old, g)
        lines[k - 1] = lines[k - 1].replace(old, new)
    c = find(lines, "\t\tthis._startAutoFlush();", find(lines, "\tconstructor("))
    lines = insert_after(lines, c, TCS_CTOR_TAIL)
    p = find(lines, "@IRequestService private readonly _requestService: IRequestService,")
    lines = insert_after(lines, p, TCS_CTOR_PARAM)
    f = find(lines, "\tprivate readonly _traceApiUrl: string;")
    lines = insert_after(lines, f, TCS_FIELDS)
    i = find(lines, "import { CancellationToken } from")
    lines = insert_after(lines, i, TCS_IMPORTS)
    return lines


# ------------------------------------------------------------------------------------------------ apoService.ts
APO_IMPORTS = """import { VSBuffer } from '../../../../base/common/buffer.js';
import { IApoScoringService } from './apoScoringService.js';
import { CorpusReportNumbers, EMPTY_DIMS, decodeCorpusReport, encodeTraceRecords } from './traceRecordCodec.js';
"""

APO_CTOR_PARAM = """		@IApoScoringService private readonly _scoring: IApoScoringService,
"""

APO_BUILD_REPORT = """	private async _buildReport(traces: ConversationTrace[]): Promise<PromptEffectivenessReport> {
		const now = Date.now();

		// One call of the B200 scoring engine replaces the trace loops of this method and of _analyzePatterns: feedback tallies,
		// per-mode tallies and goodRate, mean finalReward, per-dimension sum / count / avg with the dimension rules, and the six
		// 'bad'-gated pattern scans with their first three matches (struct apo_corpus_report, decoded by traceRecordCodec).
		const blocks = await this._scoring.score(EMPTY_DIMS, 1, 4, encodeTraceRecords(traces), 0);
		const R = decodeCorpusReport(blocks.report);

		let oldestTime = Infinity;
		let newestTime = 0;
		// Key orders of the reference objects (first appearance) — they decide the order of the generated suggestions
		const modeOrder: string[] = [];
		const dimOrder: string[] = [];
		for (const trace of traces) {
			if (trace.startTime < oldestTime) oldestTime = trace.startTime;
			if (trace.startTime > newestTime) newestTime = trace.startTime;
			const modeKey = this._extractMode(trace);
			if (!modeOrder.includes(modeKey)) modeOrder.push(modeKey);
			if (trace.summary.finalReward !== null && dimOrder.length < 9) {
				for (const dim of trace.summary.rewardDimensions) {
					if (!dimOrder.includes(dim.name)) dimOrder.push(dim.name);
				}
			}
		}

		const goodCount = R.good;
		const badCount = R.bad;
		const noFeedbackCount = R.none;
		const goodRate = R.goodRate;
		const avgReward = R.avgReward;
		const byMode: Record<string, { total: number; good: number; bad: number; goodRate: number }> = {};
		for (const modeKey of modeOrder) {
			// modes outside the four known chat modes are tallied under 'unknown' by the engine (one record code)
			const m = R.byMode[modeKey] ?? R.byMode['unknown'];
			if (m) byMode[modeKey] = { total: m.total, good: m.good, bad: m.bad, goodRate: m.goodRate };
		}
		const rewardByDimension: Record<string, { sum: number; count: number; avg: number }> = {};
		for (const name of dimOrder) {
			const d = R.rewardByDimension[name];
			if (d) rewardByDimension[name] = { sum: d.sum, count: d.count, avg: d.avg };
		}

		// Analyze problem patterns
		const patterns = this._analyzePatterns(R, traces);

		// Add additional problem patterns based on reward dimension analysis
		for (const [dimName, dimStats] of Object.entries(rewardByDimension)) {
			const low = R.rewardByDimension[dimName].low;            // avg < -0.3 && count >= 5, severity high when avg < -0.5
			if (low !== null) {
				const categoryMap: Record<string, PromptSegmentCategory> = {
					'tool_success_rate': 'tool_usage',
					'tool_call_reliability': 'tool_usage',
					'tool_call_efficiency': 'tool_usage',
					'tool_duration_efficiency': 'tool_usage',
					'token_efficiency': 'context_management',
					'response_efficiency': 'core_behavior',
					'conversation_efficiency': 'core_behavior',
					'task_completion': 'core_behavior',
					'user_feedback': 'core_behavior',
				};
				patterns.push({
					id: generateUuid(),
					description: `${dimName} dimension reward signal consistently low (avg: ${dimStats.avg.toFixed(3)})`,
					frequency: dimStats.count,
					severity: low,
					relatedCategory: categoryMap[dimName] || 'core_behavior',
					examples: [],
				});
			}
		}

		// Generate local suggestions (enhanced: pass in reward data)
		const suggestions = this._generateLocalSuggestions(goodRate, patterns, byMode, avgReward, rewardByDimension);

		const report: PromptEffectivenessReport = {
			id: generateUuid(),
			generatedAt: now,
			period: { from: oldestTime === Infinity ? now : oldestTime, to: newestTime || now },
			totalConversations: traces.length,
			goodFeedbackCount: goodCount,
			badFeedbackCount: badCount,
			noFeedbackCount: noFeedbackCount,
			goodRate,
			byMode,
			patterns,
			suggestions,
		};

		// Add suggestions to global list
		for (const s of suggestions) {
			this._suggestions.push(s);
		}

		if (suggestions.length > 0) {
			this._onDidGenerateSuggestions.fire(suggestions);
		}

		return report;
	}
"""

APO_ANALYZE = """	private _analyzePatterns(R: CorpusReportNumbers, traces: ConversationTrace[]): PromptIssuePattern[] {
		// The six scans ran on the engine (counts, emitted-flags with the minimum counts, severities, first three matching
		// trace indices in corpus order); only the presentation of the examples is assembled here.
		const patterns: PromptIssuePattern[] = [];
		if (R.bad === 0) return patterns;

		const firstUser = (t: ConversationTrace) => t.spans.find(s => s.type === 'user_message')?.data.contentPreview || '';
		const spec: Array<{ description: string; category: PromptSegmentCategory; assistant: (t: ConversationTrace) => string }> = [
			{
				description: 'Users give negative feedback after errors occur in conversations', category: 'core_behavior',
				assistant: t => t.spans.find(s => s.type === 'assistant_message')?.data.contentPreview || '',
			},
			{
				description: 'Tool call failures lead to user dissatisfaction', category: 'tool_usage',
				assistant: t => {
					const failedTool = t.spans.find(s => s.type === 'tool_call' && s.data.toolSuccess === false);
					return `Tool ${failedTool?.data.toolName} failed: ${failedTool?.data.toolResult?.substring(0, 100) || ''}`;
				},
			},
			{
				description: 'User feedback is poor in conversations with high token consumption', category: 'context_management',
				assistant: t => `Total tokens: ${t.summary.totalTokens}`,
			},
			{
				description: 'Users still dissatisfied after multiple LLM calls (possible retries)', category: 'core_behavior',
				assistant: t => `LLM calls: ${t.summary.totalLLMCalls}`,
			},
			{
				description: 'Long conversations with many turns still result in user dissatisfaction', category: 'core_behavior',
				assistant: t => `Conversation turns: ${t.spans.filter(sp => sp.type === 'user_message').length}`,
			},
			{
				description: 'Slow tool execution (>15s total) correlates with user dissatisfaction', category: 'tool_usage',
				assistant: t => `Tool duration: ${(t.summary.totalToolDurationMs / 1000).toFixed(1)}s`,
			},
		];
		R.patterns.forEach((p, k) => {
			if (!p.emitted) return;
			patterns.push({
				id: generateUuid(),
				description: spec[k].description,
				frequency: p.frequency,
				severity: p.severity,
				relatedCategory: spec[k].category,
				examples: p.examples.map(i => traces[i]).map(t => ({
					threadId: t.threadId,
					userMessagePreview: firstUser(t),
					assistantMessagePreview: spec[k].assistant(t),
					feedback: t.summary.userFeedback,
				})),
			});
		});
		return patterns;
	}
"""

APO_BEAM_CALL = """			// Update Beam Search state (ref: agent-lightning APO._update_best_prompt)
			if (serverResponse?.beamUpdate) {
				this._applyBeamUpdate(serverResponse.beamUpdate);
			}

"""

APO_BEAM_METHODS = """	/** The beam bookkeeping of requestOptimizationFromServer, shared with the engine-side evaluation below. */
	private _applyBeamUpdate(bu: { beam?: VersionedPromptTemplate[]; round?: number; bestPrompt?: VersionedPromptTemplate; bestScore?: number }): void {
		if (!this._beamState) {
			this._beamState = {
				currentRound: 0,
				totalRounds: this._config.beamRounds,
				beam: [],
				historyBestPrompt: null,
				historyBestScore: -Infinity,
				versionCounter: 0,
				startedAt: Date.now(),
				lastUpdatedAt: Date.now(),
			};
		}
		if (bu.beam) {
			this._beamState.beam = bu.beam;
		}
		if (bu.round !== undefined) {
			this._beamState.currentRound = bu.round;
		}
		if (bu.bestPrompt && bu.bestScore !== undefined && bu.bestScore > this._beamState.historyBestScore) {
			this._beamState.historyBestPrompt = bu.bestPrompt;
			this._beamState.historyBestScore = bu.bestScore;
			// Auto-apply best prompt as optimized rules
			this._applyBeamBestPrompt(bu.bestPrompt);
		}
		this._beamState.lastUpdatedAt = Date.now();
	}

	/**
	 * Client-side beam selection on the B200 scoring engine (the step the closed backend performs behind /api/apo/optimize):
	 * `evaluations` holds one 9-dimension reward row per (candidate, trace) — float32[candidates.length][T][9], NaN = dimension
	 * absent — the engine returns score[c] = mean finalReward and the top-K (K = beamWidth; score descending, ties -> lower
	 * index), and the result goes through the same strict-'>' adoption as a server beamUpdate.
	 */
	private async _evaluateBeam(candidates: VersionedPromptTemplate[], evaluations: VSBuffer, T: number): Promise<void> {
		const C = candidates.length;
		const K = Math.min(this._config.beamWidth, C);
		if (C === 0 || K === 0) return;
		try {
			const blocks = await this._scoring.score(evaluations, C, T, undefined, K);
			const scores = new Float64Array(blocks.scores.buffer.slice().buffer);
			const topk = new Int32Array(blocks.topk.buffer.slice().buffer);
			const beam = Array.from(topk).map(c => ({ ...candidates[c], score: scores[c] }));
			this._applyBeamUpdate({
				beam,
				round: (this._beamState?.currentRound ?? 0) + 1,
				bestPrompt: beam[0],
				bestScore: beam[0].score,
			});
			this._dirty = true;
			this._onDidChangeState.fire();
		} catch (e) {
			console.warn('[APO] beam evaluation failed:', e instanceof Error ? e.message : String(e));
		}
	}

"""


def patch_apo(lines):
    # bottom-up
    b = find(lines, "\tprivate _applyBeamBestPrompt(bestPrompt: VersionedPromptTemplate): void {")
    lines = insert_after(lines, b - 1, APO_BEAM_METHODS)
    u = find(lines, "// Update Beam Search state (ref: agent-lightning APO._update_best_prompt)")
    s = find(lines, "// Save Textual Gradient (ref: agent-lightning compute_textual_gradient)", u)
    lines = replace_lines(lines, u, s - 1, APO_BEAM_CALL)
    a = find(lines, "\tprivate _analyzePatterns(")
    lines = replace_lines(lines, a, method_end(lines, a), APO_ANALYZE)
    r = find(lines, "\tprivate _buildReport(traces: ConversationTrace[]): PromptEffectivenessReport {")
    lines = replace_lines(lines, r, method_end(lines, r), APO_BUILD_REPORT)
    c = find(lines, "\t\tconst report = this._buildReport(traces);")
    lines[c - 1] = lines[c - 1].replace("this._buildReport(traces)", "await this._buildReport(traces)")
    p = find(lines, "@ITraceCollectorService private readonly _traceCollectorService: ITraceCollectorService,")
    lines = insert_after(lines, p, APO_CTOR_PARAM)
    i = find(lines, "import { ITraceCollectorService, ConversationTrace, UserFeedbackType } from './traceCollectorService.js';")
    lines = insert_after(lines, i, APO_IMPORTS)
    return lines


# ------------------------------------------------------------------------------------------------ app.ts
APP_IMPORTS = """import { IApoScoringService, APO_SCORING_CHANNEL } from '../../workbench/contrib/senweaver/common/apoScoringService.js';
import { ApoScoringMainService } from '../../workbench/contrib/senweaver/electron-main/apoScoringMainService.js';
"""
APP_SERVICE = """		services.set(IApoScoringService, new SyncDescriptor(ApoScoringMainService, undefined, false));
"""
APP_CHANNEL = """
		// B200 scoring engine (native addon lives in the main process; the renderer services reach it through this channel)
		const apoScoringChannel = ProxyChannel.fromService(accessor.get(IApoScoringService), disposables);
		mainProcessElectronServer.registerChannel(APO_SCORING_CHANNEL, apoScoringChannel);
"""


def patch_app(lines):
    c = find(lines, "mainProcessElectronServer.registerChannel('senweaver-channel-scm', senweaverSCMChannel);")
    lines = insert_after(lines, c, APP_CHANNEL)
    s = find(lines, "services.set(ISenweaverSCMService, new SyncDescriptor(SenweaverSCMService, undefined, false));")
    lines = insert_after(lines, s, APP_SERVICE)
    i = find(lines, "import { MetricsMainService } from '../../workbench/contrib/senweaver/electron-main/metricsMainService.js';")
    lines = insert_after(lines, i, APP_IMPORTS)
    return lines


def apply_ed(text: str, script: str) -> str:
    """Applies a `diff -e` script (commands in descending line order: NNa / NN,MMc / NN,MMd, text blocks end with '.')."""
    import re
    lines = text.splitlines(keepends=True)
    it = iter(script.split("\n"))
    for cmd in it:
        if not cmd:
            continue
        m = re.fullmatch(r"(\d+)(?:,(\d+))?([acd])", cmd)
        if not m:
            raise ValueError(f"not an ed command: {cmd!r}")
        a, b, op = int(m.group(1)), int(m.group(2) or m.group(1)), m.group(3)
        new = []
        if op in "ac":
            for ln in it:
                if ln == ".":
                    break
                new.append(ln + "\n")
        if op == "a":
            lines[a:a] = new
        elif op == "c":
            lines[a - 1:b] = new
        else:
            lines[a - 1:b] = []
    return "".join(lines)


TARGETS = ((TCS, "patch_tcs", "traceCollectorService.ts.ed"), (APO, "patch_apo", "apoService.ts.ed"), (APP, "patch_app", "app.ts.ed"))


def patched_text(reference: str, rel: str) -> str:
    fn = {TCS: patch_tcs, APO: patch_apo, APP: patch_app}[rel]
    lines = open(os.path.join(reference, rel), encoding="utf-8").read().splitlines(keepends=True)
    return "".join(fn(list(lines)))


def main():
    import hashlib
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--apply", nargs=2, metavar=("ED_SCRIPT", "FILE"), help="apply one committed ed script to FILE in place (no `ed` binary needed)")
    ap.add_argument("--unified", default=None, help="also write reviewable unified diffs (they quote reference lines: not committed) into this directory")
    args = ap.parse_args()
    if args.apply:
        script, target = args.apply
        patched = apply_ed(open(target, encoding="utf-8").read(), open(script, encoding="utf-8").read())
        open(target, "w", encoding="utf-8").write(patched)
        print(f"patched {target}")
        return
    for rel, _, out in TARGETS:
        src = os.path.join(args.reference, rel)
        new = patched_text(args.reference, rel)
        with tempfile.NamedTemporaryFile("w", suffix=".ts", delete=False, encoding="utf-8") as f:
            f.write(new)
        r = subprocess.run(["diff", "-e", src, f.name], capture_output=True, text=True)
        assert r.returncode == 1, (rel, r.returncode, r.stderr)
        sha = hashlib.sha256(open(src, "rb").read()).hexdigest()
        with open(os.path.join(HERE, out), "w", encoding="utf-8") as o:
            o.write(r.stdout)
        with open(os.path.join(HERE, out + ".base"), "w", encoding="utf-8") as o:
            o.write(f"{rel}\nsha256 {sha}\n")
        ncmd = sum(1 for ln in r.stdout.splitlines() if ln and ln[0].isdigit() and ln[-1] in "acd")
        print(f"wrote ts/patches/{out}: {ncmd} edit commands against {rel} (sha256 {sha[:16]}...)")
        if args.unified:
            os.makedirs(args.unified, exist_ok=True)
            u = subprocess.run(["diff", "-U2", "--label", "a/" + rel, "--label", "b/" + rel, src, f.name], capture_output=True, text=True)
            open(os.path.join(args.unified, out.replace(".ed", ".patch")), "w", encoding="utf-8").write(u.stdout)
        os.unlink(f.name)


if __name__ == "__main__":
    main()
