// Ingest: the reference's persisted trace array -> Form R records.
//
// Input is what TraceCollectorService._saveToStorage writes under the storage key
// 'senweaver.traceCollector.data' (TCS:296-359): JSON.stringify(ConversationTrace[]), each trace
// {id, threadId, startTime, endTime?, spans:[TraceSpan], metadata?:{chatMode?..}, summary:{..}}
// (TCS:83-109, span TCS:30-81).  Output is one apo_record per trace, in array order (= corpus order,
// the order the pattern examples of APO:635-773 are reported in).
//
// Host-side format code only: one pass, no allocation per token, no DOM.  The scoring itself never
// runs here — the records go to apo_corpus_upload / apo_rollouts_upload.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/apo_b200.h"

namespace {

struct Cur {
	const char *p, *end, *base;
	bool ok = true;
	void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
	bool eat(char c) { ws(); if (p < end && *p == c) { ++p; return true; } return false; }
	bool peek(char c) { ws(); return p < end && *p == c; }
	void bad() { ok = false; }
};

// A JSON string; returns a view of the raw (still escaped) bytes.  Keys and the enum-like values this
// parser compares against contain no escapes in what JSON.stringify emits, so raw comparison is exact.
struct Str { const char *s; size_t n; bool is(const char *lit) const { return n == strlen(lit) && memcmp(s, lit, n) == 0; } };

bool parse_string(Cur &c, Str &out) {
	c.ws();
	if (c.p >= c.end || *c.p != '"') { c.bad(); return false; }
	const char *s = ++c.p;
	while (c.p < c.end) {
		const char *q = static_cast<const char *>(memchr(c.p, '"', size_t(c.end - c.p)));
		if (!q) break;
		// count the backslashes directly before the quote: an odd run escapes it
		size_t bs = 0;
		for (const char *b = q; b > s && b[-1] == '\\'; --b) ++bs;
		c.p = q + 1;
		if ((bs & 1) == 0) { out = {s, size_t(q - s)}; return true; }
	}
	c.bad();
	return false;
}

enum Kind { K_NULL, K_TRUE, K_FALSE, K_NUM, K_STR, K_OBJ, K_ARR };
struct Val { Kind k; double num; Str str; };

bool skip_value(Cur &c, int depth);

// Parses a scalar fully; for containers stops just before '{' / '['.
bool parse_head(Cur &c, Val &v) {
	c.ws();
	if (c.p >= c.end) { c.bad(); return false; }
	char ch = *c.p;
	if (ch == '{') { v.k = K_OBJ; return true; }
	if (ch == '[') { v.k = K_ARR; return true; }
	if (ch == '"') { v.k = K_STR; return parse_string(c, v.str); }
	auto lit = [&](const char *w, Kind k) {
		size_t n = strlen(w);
		if (size_t(c.end - c.p) >= n && memcmp(c.p, w, n) == 0) { c.p += n; v.k = k; return true; }
		c.bad();
		return false;
	};
	if (ch == 't') return lit("true", K_TRUE);
	if (ch == 'f') return lit("false", K_FALSE);
	if (ch == 'n') return lit("null", K_NULL);
	if (ch == '-' || (ch >= '0' && ch <= '9')) {
		char buf[64];
		size_t n = 0;
		const char *q = c.p;
		while (q < c.end && n < sizeof buf - 1 && (*q == '-' || *q == '+' || *q == '.' || *q == 'e' || *q == 'E' || (*q >= '0' && *q <= '9'))) buf[n++] = *q++;
		buf[n] = 0;
		char *e = nullptr;
		v.num = strtod(buf, &e);
		if (e == buf) { c.bad(); return false; }
		c.p += (e - buf);
		v.k = K_NUM;
		return true;
	}
	c.bad();
	return false;
}

bool skip_container(Cur &c, int depth) {
	if (depth > 256) { c.bad(); return false; }
	if (c.eat('{')) {
		if (c.eat('}')) return true;
		do {
			Str k;
			if (!parse_string(c, k) || !c.eat(':') || !skip_value(c, depth + 1)) { c.bad(); return false; }
		} while (c.eat(','));
		if (!c.eat('}')) { c.bad(); return false; }
		return true;
	}
	if (c.eat('[')) {
		if (c.eat(']')) return true;
		do { if (!skip_value(c, depth + 1)) return false; } while (c.eat(','));
		if (!c.eat(']')) { c.bad(); return false; }
		return true;
	}
	c.bad();
	return false;
}

bool skip_value(Cur &c, int depth) {
	c.ws();
	if (c.p < c.end) {
		const char ch = *c.p;
		if (ch == '"') { Str s; return parse_string(c, s); }
		if (ch == '{' || ch == '[') return skip_container(c, depth);
		if (ch == '-' || (ch >= '0' && ch <= '9')) {          // a skipped number needs its extent, not its value
			const char *q = c.p + 1;
			while (q < c.end && ((*q >= '0' && *q <= '9') || *q == '.' || *q == 'e' || *q == 'E' || *q == '+' || *q == '-')) ++q;
			c.p = q;
			return true;
		}
	}
	Val v;
	return parse_head(c, v);                                  // true / false / null, or an error
}

// Calls fn(key, cursor-at-value) for every member of the object at the cursor; fn must consume the value.
template <class F> bool each_member(Cur &c, F fn) {
	if (!c.eat('{')) { c.bad(); return false; }
	if (c.eat('}')) return true;
	do {
		Str k;
		if (!parse_string(c, k) || !c.eat(':')) { c.bad(); return false; }
		if (!fn(k) || !c.ok) return false;
	} while (c.eat(','));
	if (!c.eat('}')) { c.bad(); return false; }
	return true;
}

bool truthy(const Val &v) {          // JS truthiness of a parsed scalar (containers are always truthy)
	switch (v.k) {
	case K_NULL: case K_FALSE: return false;
	case K_NUM: return v.num != 0.0 && v.num == v.num;
	case K_STR: return v.str.n != 0;
	default: return true;
	}
}

uint32_t to_u32(const Val &v) {      // counters are non-negative integers held in JS doubles
	if (v.k != K_NUM || !(v.num > 0.0)) return 0;
	return v.num >= 4294967295.0 ? 0xFFFFFFFFu : uint32_t(v.num);
}

struct SpanInfo { int type = 0; bool tool_failed = false; };   // type: 1 user_message, 2 assistant_message, 3 tool_call

bool parse_span(Cur &c, SpanInfo &si) {
	return each_member(c, [&](const Str &k) {
		Val v;
		if (k.is("type")) {
			if (!parse_head(c, v)) return false;
			if (v.k == K_STR) si.type = v.str.is("user_message") ? 1 : v.str.is("assistant_message") ? 2 : v.str.is("tool_call") ? 3 : 0;
			else if (v.k == K_OBJ || v.k == K_ARR) return skip_container(c, 2);
			return true;
		}
		if (k.is("data") && c.peek('{')) {
			return each_member(c, [&](const Str &dk) {
				Val dv;
				if (!parse_head(c, dv)) return false;
				if (dv.k == K_OBJ || dv.k == K_ARR) return skip_container(c, 3);
				if (dk.is("toolSuccess") && dv.k == K_FALSE) si.tool_failed = true;   // === false, APO:668
				return true;
			});
		}
		return skip_value(c, 2);
	});
}

bool parse_trace(Cur &c, apo_record &r) {
	memset(&r, 0, sizeof r);
	uint32_t user = 0, asst = 0;
	double dur64 = 0.0;
	bool failspan = false, ended = false, valid = false, errors = false;
	bool ok = each_member(c, [&](const Str &k) {
		Val v;
		if (k.is("endTime")) {
			if (!parse_head(c, v)) return false;
			if (v.k == K_OBJ || v.k == K_ARR) { ended = true; return skip_container(c, 1); }
			ended = truthy(v);                                                   // `trace.endTime && ...` TCS:686
			return true;
		}
		if (k.is("spans") && c.peek('[')) {
			c.eat('[');
			if (c.eat(']')) return true;
			do {
				SpanInfo si;
				if (c.peek('{')) { if (!parse_span(c, si)) return false; }
				else if (!skip_value(c, 1)) return false;
				if (si.type == 1) ++user;
				else if (si.type == 2) ++asst;
				else if (si.type == 3 && si.tool_failed) failspan = true;
			} while (c.eat(','));
			if (!c.eat(']')) { c.bad(); return false; }
			return true;
		}
		if (k.is("metadata") && c.peek('{')) {
			return each_member(c, [&](const Str &mk) {
				Val mv;
				if (!parse_head(c, mv)) return false;
				if (mv.k == K_OBJ || mv.k == K_ARR) return skip_container(c, 2);
				if (mk.is("chatMode") && mv.k == K_STR)                           // TCS:91, APO:627-633
					r.mode = mv.str.is("normal") ? 1 : mv.str.is("agent") ? 2 : mv.str.is("gather") ? 3 : mv.str.is("designer") ? 4 : 0;
				return true;
			});
		}
		if (k.is("summary") && c.peek('{')) {
			return each_member(c, [&](const Str &sk) {
				Val sv;
				if (!parse_head(c, sv)) return false;
				if (sv.k == K_OBJ || sv.k == K_ARR) return skip_container(c, 2);   // toolCallsByName, rewardDimensions
				if (sk.is("totalLLMCalls")) r.llmCalls = to_u32(sv);
				else if (sk.is("totalToolCalls")) r.toolCalls = to_u32(sv);
				else if (sk.is("totalTokens")) r.tokens = to_u32(sv);
				else if (sk.is("toolCallsSucceeded")) r.toolSucc = to_u32(sv);
				else if (sk.is("toolCallsFailed")) r.toolFail = to_u32(sv);
				else if (sk.is("totalToolDurationMs")) { dur64 = sv.k == K_NUM ? sv.num : 0.0; r.toolDurMs = float(dur64); }
				else if (sk.is("userFeedback")) r.feedback = sv.k == K_STR ? (sv.str.is("good") ? 1 : sv.str.is("bad") ? 2 : 0) : 0;
				else if (sk.is("hasErrors")) errors = truthy(sv);
				else if (sk.is("finalReward")) valid = sv.k == K_NUM;               // !== null, TCS:606 / APO:550
				return true;
			});
		}
		return skip_value(c, 1);
	});
	if (!ok) return false;
	r.durClass = apo_duration_class(dur64, r.toolCalls);     // the comparisons of TCS:721-728 / APO:754 on the stored double
	r.userMsgs = uint16_t(user > 65535u ? 65535u : user);
	r.asstMsgs = uint16_t(asst > 65535u ? 65535u : asst);
	r.flags = uint8_t((errors ? APO_F_ERRORS : 0) | (ended ? APO_F_ENDED : 0) | (valid ? APO_F_VALID : 0) | (failspan ? APO_F_FAILSPAN : 0));
	return true;
}

}  // namespace

extern "C" uint8_t apo_duration_class(double dur, uint32_t totalToolCalls) {
	uint8_t dc = APO_DC_SET;
	if (dur > 0) dc |= APO_DC_POS;                           // TCS:721 (NaN compares false, as in JS)
	if (dur > 15000) dc |= APO_DC_SLOW;                      // APO:754
	if (totalToolCalls > 0 && dur > 0) {
		const double avg = dur / (double)totalToolCalls;     // TCS:722
		dc |= uint8_t((avg > 1000) + (avg > 3000) + (avg > 10000));
	}
	return dc;
}

extern "C" int64_t apo_records_from_json(const char *json, uint64_t len, apo_record *out, uint64_t cap, uint64_t *err_pos) {
	if (err_pos) *err_pos = 0;
	if (!json || (cap && !out)) return APO_E_ARG;
	Cur c{json, json + len, json};
	int64_t n = 0;
	auto fail = [&]() -> int64_t { if (err_pos) *err_pos = uint64_t(c.p - c.base); return APO_E_ARG; };
	if (!c.eat('[')) return fail();
	if (!c.eat(']')) {
		do {
			apo_record r;
			if (!c.peek('{')) { c.bad(); return fail(); }
			if (!parse_trace(c, r) || !c.ok) return fail();
			if (uint64_t(n) < cap) out[n] = r;
			++n;
		} while (c.eat(','));
		if (!c.eat(']')) return fail();
	}
	c.ws();
	if (c.p != c.end) return fail();
	return n;
}

// =================================================================== Form D -> Form Q on the host (wire format)
// The CPU twin of k_transcode + k_recode (apo_compact.cu): per coded dimension (0,1,3,4,5,6,7,8) the distinct fp32 bit
// patterns, ordered by value (-0.0 before +0.0), become dense one-byte codes (255 = absent); d2 keeps its fp32 value
// (+0.0 when absent); li = the presence mask rotated like apo::lut_index.  Same planes, same codebook, bit for bit, as the
// device transcoder produces for the same tensor (tests/test_gpu_compact.py).  Two passes on `nthreads` host threads.
#include <algorithm>
#include <thread>
#include <vector>

namespace {
inline uint32_t fbits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }
inline int dim_of(int j) { return j < 2 ? j : j + 1; }

// Distinct fp32 bit patterns of one coded dimension (<= 255) with a value: open addressing over 1024 slots, so that the hit every
// evaluation after the first few takes is one predictable probe (a linear scan of the list ended at a random position: one branch
// mispredict per dimension and evaluation).  0xFFFFFFFF = free slot: a NaN pattern, and NaNs (absent) never get here.
struct DistinctSet {
	static constexpr uint32_t FREE = 0xFFFFFFFFu;
	uint32_t v[256]; int n = 0; bool overflow = false;
	uint32_t key[1024]; uint8_t val[1024];
	DistinctSet() { memset(key, 0xFF, sizeof key); }
	static inline uint32_t slot(uint32_t b) { return (b * 2654435761u) >> 22; }
	inline void add(uint32_t b) {
		uint32_t i = slot(b);
		while (key[i] != b) {
			if (key[i] == FREE) {
				if (n == 255) { overflow = true; return; }
				key[i] = b; val[i] = (uint8_t)n; v[n++] = b;
				return;
			}
			i = (i + 1) & 1023u;
		}
	}
	// value -> code map of a finished codebook column (codes = positions in `dense`)
	void assign(const std::vector<uint32_t> &dense) {
		memset(key, 0xFF, sizeof key); n = 0; overflow = false;
		for (size_t c = 0; c < dense.size(); c++) { add(dense[c]); }
	}
	inline uint32_t code_of(uint32_t b, uint32_t absent) const {
		uint32_t i = slot(b);
		while (key[i] != b) {
			if (key[i] == FREE) return absent;
			i = (i + 1) & 1023u;
		}
		return val[i];
	}
};

template <class F>
void run_threads(int nthreads, F fn) {
	if (nthreads <= 1) { fn(0, 1); return; }
	std::vector<std::thread> th;
	for (int k = 1; k < nthreads; k++) th.emplace_back(fn, k, nthreads);
	fn(0, nthreads);
	for (auto &t : th) t.join();
}
}  // namespace

extern "C" int apo_compact_encode_host(const float *dims, uint32_t C, uint64_t T, uint64_t *q8, float *d2, uint16_t *li,
                                       uint32_t *codebook, int nthreads) {
	if ((C && T && (!dims || !q8 || !d2 || !li)) || !codebook) return APO_E_ARG;
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	const uint64_t N = (uint64_t)C * T;
	// ---- pass 1: distinct values per coded dimension
	std::vector<DistinctSet> sets((size_t)nthreads * 8);
	run_threads(nthreads, [&](int k, int n) {
		DistinctSet *mine = &sets[(size_t)k * 8];
		const uint64_t a = N * (uint64_t)k / (uint64_t)n, b = N * (uint64_t)(k + 1) / (uint64_t)n;
		for (uint64_t i = a; i < b; i++) {
			const float *row = dims + i * APO_NDIM;
			for (int j = 0; j < 8; j++) { const float f = row[dim_of(j)]; if (f == f) mine[j].add(fbits(f)); }
		}
	});
	std::vector<uint32_t> dense[8];
	for (int j = 0; j < 8; j++) {
		DistinctSet all;
		for (int k = 0; k < nthreads; k++) {
			const DistinctSet &s = sets[(size_t)k * 8 + j];
			if (s.overflow) all.overflow = true;
			for (int i = 0; i < s.n; i++) all.add(s.v[i]);
		}
		if (all.overflow) return APO_E_STATE;              // not categorical: more than 255 distinct values (Form D must be kept)
		dense[j].assign(all.v, all.v + all.n);
		std::sort(dense[j].begin(), dense[j].end(), [](uint32_t x, uint32_t y) {
			float fx, fy; memcpy(&fx, &x, 4); memcpy(&fy, &y, 4);
			if (fx != fy) return fx < fy;
			return x > y;                                    // -0.0 (sign bit set) before +0.0
		});
		for (int c = 0; c < 256; c++) codebook[256 * j + c] = c < (int)dense[j].size() ? dense[j][c] : 0xFFFFFFFFu;
	}
	// ---- pass 2: encode
	std::vector<DistinctSet> codes(8);
	for (int j = 0; j < 8; j++) codes[(size_t)j].assign(dense[j]);
	run_threads(nthreads, [&](int k, int n) {
		const uint64_t a = N * (uint64_t)k / (uint64_t)n, b = N * (uint64_t)(k + 1) / (uint64_t)n;
		for (uint64_t i = a; i < b; i++) {
			const float *row = dims + i * APO_NDIM;
			uint64_t q = 0; uint32_t mask = 0;
			for (int j = 0; j < 8; j++) {
				const float f = row[dim_of(j)];
				uint32_t code = 255u;
				if (f == f) {
					code = codes[(size_t)j].code_of(fbits(f), 255u);
					mask |= 1u << dim_of(j);
				}
				q |= (uint64_t)code << (8 * j);
			}
			const float f2 = row[2];
			const bool p2 = (f2 == f2);
			mask |= (p2 ? 1u : 0u) << 2;
			q8[i] = q;
			d2[i] = p2 ? f2 : 0.0f;
			li[i] = (uint16_t)(((mask >> 5) | (mask << 4)) & 511u);      // apo::lut_index: dims 5..8 in the low bits
		}
	});
	return APO_OK;
}

// Form D -> Form P on the host (6 B / evaluation: a uint32 of eight 4-bit codes + a 12-bit tool_success_rate index, stored as
// uint16): the PCIe wire format of apo_score_host_packed.  codebook as apo_compact_encode_host (<= 15 codes used per
// dimension), d2book[4096] = the distinct fp32 values of dimension 2 ordered by value (unused = 0xFFFFFFFF).
// APO_E_STATE when a coded dimension has more than 15 distinct values or dimension 2 more than 4095.
extern "C" int apo_packed_encode_host(const float *dims, uint32_t C, uint64_t T, uint32_t *pc, uint16_t *pd,
                                      uint32_t *codebook, uint32_t *d2book, int nthreads) {
	if ((C && T && (!dims || !pc || !pd)) || !codebook || !d2book) return APO_E_ARG;
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	const uint64_t N = (uint64_t)C * T;
	// ---- pass 1: distinct values (coded dimensions: tiny sets; dimension 2: a sorted vector per thread)
	std::vector<DistinctSet> sets((size_t)nthreads * 8);
	std::vector<std::vector<uint32_t>> d2sets((size_t)nthreads);
	std::vector<int> too_many((size_t)nthreads, 0);
	run_threads(nthreads, [&](int k, int n) {
		DistinctSet *mine = &sets[(size_t)k * 8];
		std::vector<uint32_t> &seen = d2sets[(size_t)k];
		// open addressing over the bit patterns: dimension 2 has a few hundred distinct ratios at most
		std::vector<uint32_t> table(16384, 0xFFFFFFFFu);
		const uint64_t a = N * (uint64_t)k / (uint64_t)n, b = N * (uint64_t)(k + 1) / (uint64_t)n;
		for (uint64_t i = a; i < b; i++) {
			const float *row = dims + i * APO_NDIM;
			for (int j = 0; j < 8; j++) { const float f = row[dim_of(j)]; if (f == f) mine[j].add(fbits(f)); }
			const float f2 = row[2];
			if (f2 == f2 && !too_many[(size_t)k]) {
				const uint32_t bits = fbits(f2);
				uint32_t h = (bits * 2654435761u) >> 18;
				for (;;) {
					if (table[h] == bits) break;
					if (table[h] == 0xFFFFFFFFu) { table[h] = bits; seen.push_back(bits); if (seen.size() > 4095) too_many[(size_t)k] = 1; break; }
					h = (h + 1) & 16383u;
				}
			}
		}
	});
	std::vector<uint32_t> dense[8], d2all;
	auto by_value = [](uint32_t x, uint32_t y) {
		float fx, fy; memcpy(&fx, &x, 4); memcpy(&fy, &y, 4);
		if (fx != fy) return fx < fy;
		return x > y;
	};
	for (int j = 0; j < 8; j++) {
		DistinctSet all;
		for (int k = 0; k < nthreads; k++) {
			const DistinctSet &st = sets[(size_t)k * 8 + j];
			if (st.overflow) all.overflow = true;
			for (int i = 0; i < st.n; i++) all.add(st.v[i]);
		}
		if (all.overflow || all.n > 15) return APO_E_STATE;
		dense[j].assign(all.v, all.v + all.n);
		std::sort(dense[j].begin(), dense[j].end(), by_value);
		for (int c = 0; c < 256; c++) codebook[256 * j + c] = c < (int)dense[j].size() ? dense[j][c] : 0xFFFFFFFFu;
	}
	for (int k = 0; k < nthreads; k++) {
		if (too_many[(size_t)k]) return APO_E_STATE;
		d2all.insert(d2all.end(), d2sets[(size_t)k].begin(), d2sets[(size_t)k].end());
	}
	std::sort(d2all.begin(), d2all.end(), by_value);
	d2all.erase(std::unique(d2all.begin(), d2all.end()), d2all.end());
	if (d2all.size() > 4095) return APO_E_STATE;
	for (size_t c = 0; c < 4096; c++) d2book[c] = c < d2all.size() ? d2all[c] : 0xFFFFFFFFu;
	// ---- pass 2: encode (value -> code through hash maps: the binary search of the sorted ratios mispredicted ~11 branches per
	// evaluation)
	std::vector<DistinctSet> codes(8);
	for (int j = 0; j < 8; j++) codes[(size_t)j].assign(dense[j]);
	std::vector<uint32_t> d2key(16384, 0xFFFFFFFFu);
	std::vector<uint16_t> d2val(16384, 4095);
	for (size_t c = 0; c < d2all.size(); c++) {
		uint32_t h = (d2all[c] * 2654435761u) >> 18;
		while (d2key[h] != 0xFFFFFFFFu) h = (h + 1) & 16383u;
		d2key[h] = d2all[c]; d2val[h] = (uint16_t)c;
	}
	run_threads(nthreads, [&](int k, int n) {
		const uint64_t a = N * (uint64_t)k / (uint64_t)n, b = N * (uint64_t)(k + 1) / (uint64_t)n;
		for (uint64_t i = a; i < b; i++) {
			const float *row = dims + i * APO_NDIM;
			uint32_t w = 0;
			for (int j = 0; j < 8; j++) {
				const float f = row[dim_of(j)];
				const uint32_t code = f == f ? codes[(size_t)j].code_of(fbits(f), 15u) : 15u;
				w |= code << (4 * j);
			}
			const float f2 = row[2];
			uint16_t k2 = 4095;
			if (f2 == f2) {
				const uint32_t bits = fbits(f2);
				uint32_t h = (bits * 2654435761u) >> 18;
				while (d2key[h] != bits) h = (h + 1) & 16383u;                // present by construction (pass 1 saw every value)
				k2 = d2val[h];
			}
			pc[i] = w;
			pd[i] = k2;
		}
	});
	return APO_OK;
}

// ---- Form T: Form P planes -> 24-bit dictionary indices (csrc/apo_tuple.cu) -------------------------------------------
namespace {
// open addressing over 44-bit keys (code word | tool_success_rate index << 32); EMPTY is not a key (bits 44.. set)
struct TupleMap {
	static constexpr uint64_t EMPTY = ~0ull;
	std::vector<uint64_t> key, val;
	uint64_t mask = 0, used = 0;
	explicit TupleMap(uint64_t slots = 1u << 16) { reset(slots); }
	void reset(uint64_t slots) { key.assign(slots, EMPTY); val.assign(slots, 0); mask = slots - 1; used = 0; }
	static uint64_t hash(uint64_t k) { k ^= k >> 29; k *= 0xBF58476D1CE4E5B9ull; k ^= k >> 32; return k; }
	void grow() {
		std::vector<uint64_t> ok, ov;
		ok.swap(key); ov.swap(val);
		reset((mask + 1) * 4);
		for (size_t i = 0; i < ok.size(); i++) if (ok[i] != EMPTY) *slot(ok[i]) = ov[i];
	}
	// value cell of k, inserted with 0 when new
	uint64_t *slot(uint64_t k) {
		if ((used + 1) * 2 > mask + 1) grow();
		uint64_t h = hash(k) & mask;
		for (;;) {
			if (key[h] == k) return &val[h];
			if (key[h] == EMPTY) { key[h] = k; used++; return &val[h]; }
			h = (h + 1) & mask;
		}
	}
	const uint64_t *find(uint64_t k) const {
		uint64_t h = hash(k) & mask;
		for (;;) {
			if (key[h] == k) return &val[h];
			if (key[h] == EMPTY) return nullptr;
			h = (h + 1) & mask;
		}
	}
};
inline uint64_t tuple_key(uint32_t pc, uint16_t pd) { return (uint64_t)pc | ((uint64_t)(pd & 4095u) << 32); }
}  // namespace

extern "C" int apo_tuple_encode_host(const uint32_t *pc, const uint16_t *pd, uint32_t C, uint64_t T, uint16_t *tl, uint8_t *th,
                                     uint32_t *tbook_pc, uint16_t *tbook_pd, uint32_t cap, uint32_t *n_tuples, int nthreads) {
	if (!n_tuples || (cap && (!tbook_pc || !tbook_pd)) || (C && T && (!pc || !pd || !tl || !th))) return APO_E_ARG;
	*n_tuples = 0;
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	if (cap > 0xFFFFFFu) cap = 0xFFFFFFu;
	const uint64_t N = (uint64_t)C * T;
	// ---- pass 1: distinct evaluations and how often each occurs, per thread
	std::vector<TupleMap> maps((size_t)nthreads);
	std::vector<uint64_t> over((size_t)nthreads, 0);
	run_threads(nthreads, [&](int k, int n) {
		TupleMap &m = maps[(size_t)k];
		const uint64_t a = N * (uint64_t)k / (uint64_t)n, b = N * (uint64_t)(k + 1) / (uint64_t)n;
		for (uint64_t i = a; i < b; i++) {
			++*m.slot(tuple_key(pc[i], pd[i]));
			// not categorical (free-form data): stop before the maps grow with the tensor; the count reported is a lower bound
			if (m.used > cap) { over[(size_t)k] = m.used; return; }
		}
	});
	for (int k = 0; k < nthreads; k++)
		if (over[(size_t)k]) { *n_tuples = over[(size_t)k] > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)over[(size_t)k]; return APO_E_STATE; }
	TupleMap all(1u << 18);
	for (int k = 0; k < nthreads; k++) {
		const TupleMap &m = maps[(size_t)k];
		for (size_t i = 0; i < m.key.size(); i++) if (m.key[i] != TupleMap::EMPTY) *all.slot(m.key[i]) += m.val[i];
		maps[(size_t)k].reset(2);
		if (all.used > cap) break;
	}
	if (all.used > cap) { *n_tuples = all.used > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)all.used; return APO_E_STATE; }
	// ---- dictionary order: most frequent first (the head of the table lives in shared memory), ties by key: deterministic
	std::vector<std::pair<uint64_t, uint64_t>> order;              // (count, key)
	order.reserve(all.used);
	for (size_t i = 0; i < all.key.size(); i++) if (all.key[i] != TupleMap::EMPTY) order.push_back({all.val[i], all.key[i]});
	std::sort(order.begin(), order.end(), [](const std::pair<uint64_t, uint64_t> &x, const std::pair<uint64_t, uint64_t> &y) {
		if (x.first != y.first) return x.first > y.first;
		return x.second < y.second;
	});
	for (size_t r = 0; r < order.size(); r++) {
		tbook_pc[r] = (uint32_t)order[r].second;
		tbook_pd[r] = (uint16_t)(order[r].second >> 32);
		*all.slot(order[r].second) = r;                               // count -> index
	}
	*n_tuples = (uint32_t)order.size();
	// ---- pass 2: encode
	const TupleMap &index = all;
	run_threads(nthreads, [&](int k, int n) {
		const uint64_t a = N * (uint64_t)k / (uint64_t)n, b = N * (uint64_t)(k + 1) / (uint64_t)n;
		for (uint64_t i = a; i < b; i++) {
			const uint64_t r = *index.find(tuple_key(pc[i], pd[i]));
			tl[i] = (uint16_t)r;
			th[i] = (uint8_t)(r >> 16);
		}
	});
	return APO_OK;
}
