/*---------------------------------------------------------------------------------------------
 *  NEW FILE for src/vs/workbench/contrib/senweaver/common/traceRecordCodec.ts
 *
 *  ConversationTrace  ->  32-byte Form R record (include/apo_b200.h `apo_record`), and
 *  `apo_corpus_report` (784 bytes)  ->  the numbers APOService._buildReport needs.
 *  The Python mirror senweaver-ide_b200/trace_collector.py::encode_trace is the executable
 *  twin of encodeTraceRecord and is what the engine's tests exercise.
 *--------------------------------------------------------------------------------------------*/
import { VSBuffer } from '../../../../base/common/buffer.js';
import type { ConversationTrace } from './traceCollectorService.js';

export const APO_RECORD_BYTES = 32;
/** A 1 x 4 block of all-absent evaluations (36 floats of NaN): the dims argument of a report-only scoring call. */
export const EMPTY_DIMS: VSBuffer = VSBuffer.wrap(new Uint8Array(new Float32Array(36).fill(NaN).buffer));
const F_ERRORS = 0x01, F_ENDED = 0x02, F_VALID = 0x08, F_FAILSPAN = 0x10;
const MODE_CODE: Record<string, number> = { normal: 1, agent: 2, gather: 3, designer: 4 };   // anything else -> 0 ('unknown')
const U32_MAX = 0xFFFFFFFF;
export const DIM_NAMES = [
	'user_feedback', 'task_completion', 'tool_success_rate', 'tool_call_reliability', 'tool_call_efficiency',
	'tool_duration_efficiency', 'response_efficiency', 'token_efficiency', 'conversation_efficiency',
] as const;

/** valid = "finalReward is (being) computed": pass true from _computeRewardSignals, else summary.finalReward !== null. */
export function encodeTraceRecord(trace: ConversationTrace, view: DataView, offset: number, valid = trace.summary.finalReward !== null): void {
	const s = trace.summary;
	let user = 0, asst = 0, failSpan = false;
	for (const sp of trace.spans) {
		if (sp.type === 'user_message') { user++; }
		else if (sp.type === 'assistant_message') { asst++; }
		else if (sp.type === 'tool_call' && sp.data.toolSuccess === false) { failSpan = true; }
	}
	const mode = MODE_CODE[(trace.metadata?.chatMode as string) ?? ''] ?? 0;
	view.setUint8(offset + 0, s.userFeedback === 'good' ? 1 : s.userFeedback === 'bad' ? 2 : 0);
	view.setUint8(offset + 1, (s.hasErrors ? F_ERRORS : 0) | (trace.endTime ? F_ENDED : 0) | (valid ? F_VALID : 0) | (failSpan ? F_FAILSPAN : 0));
	view.setUint8(offset + 2, mode);
	// durClass (APO_DC_*): the comparisons TCS:721-728 / APO:754 make on the double, decided here — the record keeps a float32 copy
	const dur = s.totalToolDurationMs;
	let dc = 0x80 | (dur > 0 ? 0x04 : 0) | (dur > 15000 ? 0x08 : 0);
	if (s.totalToolCalls > 0 && dur > 0) {
		const avg = dur / s.totalToolCalls;
		dc |= (avg > 1000 ? 1 : 0) + (avg > 3000 ? 1 : 0) + (avg > 10000 ? 1 : 0);
	}
	view.setUint8(offset + 3, dc);
	view.setUint16(offset + 4, Math.min(user, 0xFFFF), true);
	view.setUint16(offset + 6, Math.min(asst, 0xFFFF), true);
	view.setUint32(offset + 8, Math.min(s.totalToolCalls, U32_MAX), true);
	view.setUint32(offset + 12, Math.min(s.toolCallsSucceeded, U32_MAX), true);
	view.setUint32(offset + 16, Math.min(s.toolCallsFailed, U32_MAX), true);
	view.setUint32(offset + 20, Math.min(s.totalLLMCalls, U32_MAX), true);
	view.setUint32(offset + 24, Math.min(s.totalTokens, U32_MAX), true);
	view.setFloat32(offset + 28, s.totalToolDurationMs, true);
}

export function encodeTraceRecords(traces: readonly ConversationTrace[]): VSBuffer {
	const bytes = new Uint8Array(traces.length * APO_RECORD_BYTES);
	const view = new DataView(bytes.buffer);
	traces.forEach((t, i) => encodeTraceRecord(t, view, i * APO_RECORD_BYTES));
	return VSBuffer.wrap(bytes);
}

/** apo_reward_batch output -> the two summary fields _computeRewardSignals assigns (TCS:786-787). */
export function decodeReward(dims: VSBuffer, masks: VSBuffer, finals: VSBuffer, i = 0) {
	const d = new DataView(dims.buffer.buffer, dims.buffer.byteOffset), m = new DataView(masks.buffer.buffer, masks.buffer.byteOffset);
	const f = new DataView(finals.buffer.buffer, finals.buffer.byteOffset);
	const mask = m.getUint32(4 * i, true);
	const rewardDimensions = DIM_NAMES.map((name, k) => ({ name, value: d.getFloat64(8 * (9 * i + k), true) })).filter((_, k) => mask & (1 << k));
	const fr = f.getFloat64(8 * i, true);
	return { rewardDimensions, finalReward: Number.isNaN(fr) ? null : fr };
}

export interface CorpusReportNumbers {
	total: number; good: number; bad: number; none: number; goodRate: number;
	byMode: Record<string, { total: number; good: number; bad: number; goodRate: number }>;
	withReward: number; avgReward: number | null;
	rewardByDimension: Record<string, { sum: number; count: number; avg: number; low: 'medium' | 'high' | null; suggest: 'medium' | 'high' | null }>;
	patterns: Array<{ emitted: boolean; frequency: number; severity: 'low' | 'medium' | 'high'; examples: number[] }>;
	toolCalls: number; toolSucc: number; toolFail: number; toolSuccessRate: number | null;
}

const SEV = ['low', 'medium', 'high'] as const;
const MODES = ['unknown', 'normal', 'agent', 'gather', 'designer'];
const u64 = (v: DataView, o: number) => Number(v.getBigUint64(o, true));

/** struct apo_corpus_report, offsets of include/apo_b200.h (sizeof == 784). */
export function decodeCorpusReport(buf: VSBuffer): CorpusReportNumbers {
	const v = new DataView(buf.buffer.buffer, buf.buffer.byteOffset, 784);
	const byMode: CorpusReportNumbers['byMode'] = {};
	MODES.forEach((name, m) => {
		const total = u64(v, 40 + 24 * m);
		if (total) { byMode[name] = { total, good: u64(v, 48 + 24 * m), bad: u64(v, 56 + 24 * m), goodRate: v.getFloat64(160 + 8 * m, true) }; }
	});
	const rewardByDimension: CorpusReportNumbers['rewardByDimension'] = {};
	DIM_NAMES.forEach((name, i) => {
		const o = 224 + 32 * i, count = u64(v, o + 8);
		if (count) {
			rewardByDimension[name] = {
				sum: v.getFloat64(o, true), count, avg: v.getFloat64(o + 16, true),
				low: v.getUint8(o + 24) ? (v.getUint8(o + 25) === 2 ? 'high' : 'medium') : null,          // APO:575, 591
				suggest: v.getUint8(o + 26) ? (v.getUint8(o + 27) === 2 ? 'high' : 'medium') : null,      // APO:802, 819
			};
		}
	});
	const patterns = [0, 1, 2, 3, 4, 5].map(p => {
		const o = 512 + 40 * p;
		const examples = [0, 1, 2].map(k => Number(v.getBigInt64(o + 16 + 8 * k, true))).filter(x => x >= 0);
		return { emitted: v.getUint8(o + 8) !== 0, frequency: u64(v, o), severity: SEV[v.getUint8(o + 9)], examples };
	});
	const withReward = u64(v, 200), toolCalls = u64(v, 752);
	return {
		total: u64(v, 0), good: u64(v, 8), bad: u64(v, 16), none: u64(v, 24), goodRate: v.getFloat64(32, true), byMode,
		withReward, avgReward: withReward ? v.getFloat64(216, true) : null, rewardByDimension, patterns,
		toolCalls, toolSucc: u64(v, 760), toolFail: u64(v, 768), toolSuccessRate: toolCalls ? v.getFloat64(776, true) : null,
	};
}
