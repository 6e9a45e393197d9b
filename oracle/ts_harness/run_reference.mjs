#!/usr/bin/env node
// run_reference.mjs — pins the oracle with a STOCK JavaScript engine (Node >= 18).  Same job as run_reference.py, which
// executes the same method texts with the in-repo minijs interpreter because the engine's build image has no JS runtime:
//
//     node oracle/ts_harness/run_reference.mjs /path/to/senweaver-ide [--check]
//
// Extracts the UNMODIFIED method texts of the reference by name —
//     traceCollectorService.ts : _computeRewardSignals, getStats
//     apoService.ts            : _buildReport, _extractMode, _analyzePatterns, _generateLocalSuggestions, getStats
// — strips the TypeScript types (typescript.transpileModule when the package is installed, else module.stripTypeScriptTypes
// of Node >= 22.13, else the small regex stripper below, which covers exactly the syntax these seven methods use), evaluates
// them as methods of two plain objects holding the instance state they touch, runs them over tests/golden/ref_inputs.json
// and writes tests/golden/ref_reward_cases.json / ref_report_cases.json in the format run_reference.py writes (binary64
// values as C99 hex strings, engine "node ...").  With --check it compares against the committed fixtures instead:
// "identical" from this script on any machine with Node is the independent confirmation of the pin.
import fs from 'node:fs';
import path from 'node:path';
import crypto from 'node:crypto';
import { fileURLToPath } from 'node:url';
import { createRequire } from 'node:module';

const HERE = path.dirname(fileURLToPath(import.meta.url));
const ROOT = path.resolve(HERE, '..', '..');
const GOLDEN = path.join(ROOT, 'tests', 'golden');
const TCS_REL = 'src/vs/workbench/contrib/senweaver/common/traceCollectorService.ts';
const APO_REL = 'src/vs/workbench/contrib/senweaver/common/apoService.ts';
const refRoot = process.argv[2] && !process.argv[2].startsWith('--') ? process.argv[2] : '/root/reference';
const CHECK = process.argv.includes('--check');

// ---- method extraction: declaration at one tab of indentation that has a body (the interface holds signatures only)
function extractMethod(text, name) {
	const re = new RegExp('^\\t(?:(?:private|public|protected)\\s+)?' + name + '\\s*\\(', 'gm');
	let m;
	while ((m = re.exec(text)) !== null) {
		let i = text.indexOf('(', m.index), depth = 0;
		for (; ; i++) { if (text[i] === '(') { depth++; } else if (text[i] === ')') { if (--depth === 0) { break; } } }
		let j = i + 1, ad = 0, body = false;
		for (; j < text.length; j++) {
			const c = text[j];
			if ('<[('.includes(c)) { ad++; } else if ('>])'.includes(c)) { ad--; }
			else if (c === ';' && ad === 0) { break; }
			else if (c === '{' && ad === 0) { body = true; break; }
		}
		if (!body) { continue; }
		let k = j, d = 0, str = null;
		for (; ; k++) {
			const c = text[k];
			if (str) {
				if (c === '\\') { k++; }
				else if (c === str) { str = null; }
				else if (str === '`' && text.startsWith('${', k)) { let d2 = 1; k += 2; while (d2) { if (text[k] === '{') { d2++; } else if (text[k] === '}') { d2--; } k++; } k--; }
			} else if (c === '\'' || c === '"' || c === '`') { str = c; }
			else if (text.startsWith('//', k)) { k = text.indexOf('\n', k) - 1; }
			else if (text.startsWith('/*', k)) { k = text.indexOf('*/', k) + 1; }
			else if (c === '{') { d++; }
			else if (c === '}') { if (--d === 0) { break; } }
		}
		const src = text.slice(m.index, k + 1);
		const first = text.slice(0, m.index).split('\n').length;
		return { src, lines: [first, first + src.split('\n').length - 1] };
	}
	throw new Error('method not found: ' + name);
}

// ---- type stripping
const require = createRequire(import.meta.url);
let stripper = 'regex';
function stripTypes(methodSrc) {
	// as a class member so that a real TypeScript front end accepts it
	const wrapped = 'class X {\n' + methodSrc + '\n}';
	try {
		const ts = require('typescript');
		stripper = 'typescript ' + ts.version;
		const out = ts.transpileModule(wrapped, { compilerOptions: { target: ts.ScriptTarget.ES2022, module: ts.ModuleKind.ESNext } }).outputText;
		return out;
	} catch { /* not installed */ }
	try {
		const mod = require('node:module');
		if (typeof mod.stripTypeScriptTypes === 'function') { stripper = 'module.stripTypeScriptTypes'; return mod.stripTypeScriptTypes(wrapped); }
	} catch { /* older Node */ }
	// regex fallback: exactly the annotations these methods carry
	let s = wrapped;
	s = s.replace(/^\t(?:private|public|protected)\s+/m, '\t');                                   // member modifier
	s = s.replace(/^(\t[A-Za-z_$][\w$]*\s*\()([^)]*)\)\s*:\s*[^{]+\{/m, (all, head, params) =>       // header: param and return types
		head + params.split(',').map(p => p.replace(/\??\s*:\s*[\s\S]*$/, '').trim()).filter(Boolean).join(', ') + ') {');
	s = s.replace(/\b(const|let)\s+([A-Za-z_$][\w$]*)\s*:\s*[^=;]+?(=(?![=>])|;)/g, (all, kw, id, end) => kw + ' ' + id + (end === ';' ? ';' : ' ='));
	s = s.replace(/\s+as\s+[A-Za-z_$][\w$.]*(\[\])?/g, '');                                        // `x as string`
	return s;
}

function makeService(text, names, state) {
	const lines = {};
	let cls = 'class X {\n';
	for (const n of names) {
		const { src, lines: ln } = extractMethod(text, n);
		lines[n] = ln;
		const js = stripTypes(src);
		cls += js.replace(/^\s*class X \{\n?/, '').replace(/\}\s*$/, '') + '\n';
	}
	cls += '}\nreturn X;';
	let uuid = 0;
	const X = new Function('generateUuid', 'DateNow', cls.replace(/Date\.now\(\)/g, 'DateNow()'))(() => 'uuid-' + (++uuid), () => 1700000000000);
	const svc = Object.assign(Object.create(X.prototype), state);
	return { svc, lines };
}

// ---- inputs in the reference's object shape (oracle/ts_transcription.py make_trace / run_reference.py corpus_traces)
function makeTrace(t) {
	const [feedback, hasErrors, ended, toolCalls, succ, fail, toolDurMs, llmCalls, tokens, userMsgs, asstMsgs, mode] = t;
	const trace = {
		id: 'trace', threadId: 't', startTime: 0, spans: [], metadata: mode === null ? null : { chatMode: mode },
		summary: { totalLLMCalls: llmCalls, totalToolCalls: toolCalls, totalTokens: tokens, userFeedback: feedback, hasErrors, toolCallsSucceeded: succ,
			toolCallsFailed: fail, toolCallsByName: {}, totalToolDurationMs: toolDurMs, finalReward: null, rewardDimensions: [] },
	};
	if (ended) { trace.endTime = 1.0; }
	for (let i = 0; i < userMsgs; i++) { trace.spans.push({ type: 'user_message', data: {} }); }
	for (let i = 0; i < asstMsgs; i++) { trace.spans.push({ type: 'assistant_message', data: {} }); }
	for (let i = 0; i < toolCalls; i++) { trace.spans.push({ type: 'tool_call', data: { toolSuccess: i >= fail } }); }
	return trace;
}

function hexf(x) {                                    // C99 / Python float.hex() spelling of a binary64
	if (x === null || x === undefined) { return null; }
	if (typeof x !== 'number') { return x; }
	if (Number.isNaN(x)) { return 'nan'; }
	if (!Number.isFinite(x)) { return x > 0 ? 'inf' : '-inf'; }
	if (Number.isInteger(x) && !Object.is(x, -0) && Math.abs(x) < 2 ** 53 && HEX_INT_AS_INT.v) { return x; }
	const buf = new DataView(new ArrayBuffer(8)); buf.setFloat64(0, x);
	const hi = buf.getUint32(0), lo = buf.getUint32(4);
	const sign = hi >>> 31 ? '-' : '', exp = (hi >>> 20) & 0x7ff;
	const mant = (BigInt(hi & 0xfffff) << 32n) | BigInt(lo);
	if (exp === 0 && mant === 0n) { return sign + '0x0.0p+0'; }
	const m = mant.toString(16).padStart(13, '0');
	if (exp === 0) { return sign + '0x0.' + m + 'p-1022'; }
	const e = exp - 1023;
	return sign + '0x1.' + m + 'p' + (e >= 0 ? '+' : '') + e;
}
const HEX_INT_AS_INT = { v: false };
function hexify(o, floatKeys) {
	if (Array.isArray(o)) { return o.map(v => hexify(v, floatKeys)); }
	if (o && typeof o === 'object') { const r = {}; for (const [k, v] of Object.entries(o)) { r[k] = hexify(v, floatKeys); } return r; }
	if (typeof o === 'number') { return Number.isInteger(o) && !Object.is(o, -0) ? o : hexf(o); }
	return o === undefined ? null : o;
}
// NOTE on integers: JSON cannot tell 1 from 1.0.  run_reference.py writes a value as a hex string when the interpreter held it
// as a float (it came out of a division or a decimal literal) and as an integer when it was an integer literal / counter.  This
// script cannot know that history, so --check compares NUMERICALLY: both spellings are decoded to binary64 and compared bit for bit.
function decode(v) { return typeof v === 'string' && /^-?(0x|nan|inf)/.test(v) ? (v === 'nan' ? NaN : parseHexFloat(v)) : v; }
function parseHexFloat(s) {
	const m = /^(-?)0x([01])\.([0-9a-f]+)p([+-]\d+)$/.exec(s);
	if (!m) { return s; }
	let mant = BigInt('0x' + m[3].padEnd(13, '0')), v = Number(m[2]) + Number(mant) / 2 ** 52;
	v = v * 2 ** Number(m[4]);
	return m[1] ? -v : v;
}
function sameDeep(a, b) {
	a = decode(a); b = decode(b);
	if (typeof a === 'number' && typeof b === 'number') { return Object.is(a, b) || (Number.isNaN(a) && Number.isNaN(b)); }
	if (Array.isArray(a)) { return Array.isArray(b) && a.length === b.length && a.every((v, i) => sameDeep(v, b[i])); }
	if (a && typeof a === 'object') {
		if (!b || typeof b !== 'object') { return false; }
		const ka = Object.keys(a), kb = Object.keys(b);
		return ka.length === kb.length && ka.every((k, i) => k === kb[i] && sameDeep(a[k], b[k]));
	}
	return a === b;
}

// ---- run
const tcsText = fs.readFileSync(path.join(refRoot, TCS_REL), 'utf8');
const apoText = fs.readFileSync(path.join(refRoot, APO_REL), 'utf8');
const inputs = JSON.parse(fs.readFileSync(path.join(GOLDEN, 'ref_inputs.json'), 'utf8'));
const tcs = makeService(tcsText, ['_computeRewardSignals', 'getStats'], { _traces: new Map(), _feedbacks: new Map(), _estimateStorageBytes: () => 0 });
const spy = { last: {} };
const apo = makeService(apoText, ['_buildReport', '_extractMode', '_analyzePatterns', '_generateLocalSuggestions', 'getStats'], {
	_suggestions: [], _segments: [], _reports: [], _beamState: null, _textualGradients: [], _onDidGenerateSuggestions: { fire() { } },
	_traceCollectorService: { getAllTraces: () => Array.from(tcs.svc._traces.values()) },
});
const realSuggest = apo.svc._generateLocalSuggestions;
apo.svc._generateLocalSuggestions = function (goodRate, patterns, byMode, avgReward, rewardByDimension) {
	spy.last = { goodRate, avgReward, rewardByDimension };
	return realSuggest.call(this, goodRate, patterns, byMode, avgReward, rewardByDimension);
};

const methodLines = {};
for (const [k, v] of Object.entries(tcs.lines)) { methodLines['TCS.' + k] = v; }
for (const [k, v] of Object.entries(apo.lines)) { methodLines['APO.' + k] = v; }
const sha = t => crypto.createHash('sha256').update(t, 'utf8').digest('hex');
const provenance = {
	engine: `node ${process.version} (${stripper}) executing the unmodified reference method text`,
	reference: { [TCS_REL]: sha(tcsText), [APO_REL]: sha(apoText) }, method_lines: methodLines,
};

const cases = inputs.tuples.map(([name, tup]) => {
	const t = makeTrace(tup);
	tcs.svc._computeRewardSignals(t);
	return { name, input: tup, dims: t.summary.rewardDimensions.map(d => ({ name: d.name, value: hexf(d.value) })), finalReward: hexf(t.summary.finalReward) };
});

const corpora = {};
for (const [cname, spec] of Object.entries(inputs.corpora)) {
	const traces = spec.indices.map((i, n) => {
		const t = makeTrace(inputs.tuples[i][1]);
		t.id = `trace-${i}`; t.threadId = `thread-${i}`; t.startTime = 1000 + (i * 7919) % 1013;
		t.spans.forEach((sp, k) => {
			sp.data.contentPreview = `${sp.type} ${i}.${k}`;
			if (sp.type === 'tool_call') { sp.data.toolName = `tool${k % 3}`; sp.data.toolResult = 'x'.repeat(90 + 7 * (k % 4)); }
		});
		if (!spec.unscored.includes(n)) { tcs.svc._computeRewardSignals(t); }
		return t;
	});
	apo.svc._suggestions = []; spy.last = {};
	const report = apo.svc._buildReport(traces);
	tcs.svc._traces = new Map(traces.map(t => [t.id, t]));
	tcs.svc._feedbacks = new Map(traces.filter(t => t.summary.userFeedback).map(t => [t.id, t.summary.userFeedback]));
	const stats = { traceCollector: tcs.svc.getStats(), apo: apo.svc.getStats() };
	corpora[cname] = hexify({ report, locals: { goodRate: spy.last.goodRate, avgReward: spy.last.avgReward ?? null, rewardByDimension: spy.last.rewardByDimension ?? null }, stats, indices: spec.indices, unscored: spec.unscored });
}

const out = { 'ref_reward_cases.json': { provenance, cases }, 'ref_report_cases.json': { provenance, corpora } };
let rc = 0;
for (const [fname, obj] of Object.entries(out)) {
	const p = path.join(GOLDEN, fname);
	if (CHECK) {
		const committed = JSON.parse(fs.readFileSync(p, 'utf8'));
		// ids ("uuid-N") depend on how many the run has handed out before; everything else must agree value for value
		const scrub = o => JSON.parse(JSON.stringify(o, (k, v) => (k === 'id' || k === 'provenance' ? undefined : v)));
		const same = sameDeep(scrub(committed), scrub(JSON.parse(JSON.stringify(obj))));
		console.log(`${fname}: ${same ? 'identical to what the reference text produces under ' + provenance.engine : 'DIFFERS'}`);
		if (!same) { rc = 1; }
	} else {
		fs.writeFileSync(p, JSON.stringify(obj, null, 0));
		console.log('wrote ' + p);
	}
}
process.exit(rc);
