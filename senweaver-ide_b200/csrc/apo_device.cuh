// apo_device.cuh — device-side building blocks shared by the kernels of libapo_b200.
//
// Reward arithmetic mirrors the reference statement by statement with explicit
// round-to-nearest binary64 intrinsics (no FMA contraction), so per-evaluation results
// are bit-identical to a JS `number` evaluation of
//   TCS = src/vs/workbench/contrib/senweaver/common/traceCollectorService.ts :668-788.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/apo_b200.h"
#include "apo_kernels.h"

namespace apo {

// ---------------------------------------------------------------- accumulator vector
// One int64 vector holds every partial of a scoring call so that a single
// ncclAllReduce(sum,int64) joins record-axis shards.  Fixed-point sums use 2^-52 units
// kept as three limbs (value = l2*2^64 + l1*2^32 + l0): integer addition is associative,
// so results do not depend on grid size, scheduling, shard or rank count.
constexpr int ACC_PER_CAND = 4;            // l0,l1,l2,count
constexpr int CORP_REWARD  = 0;            // l0,l1,l2,count
constexpr int CORP_DIM     = 4;            // 9 x (l0,l1,l2,count)
constexpr int CORP_TALLY   = 40;           // good,bad,none
constexpr int CORP_MODE    = 43;           // [5][3] total,good,bad
constexpr int CORP_PAT     = 58;           // [6]
constexpr int CORP_TOOL    = 64;           // calls,succ,fail
constexpr int CORP_NREC    = 67;
constexpr int CORP_EX      = 68;           // [nranks][6][3] (index+1, 0 = none); rank r fills its own slot
constexpr int CORP_FIXED   = 68;
__host__ __device__ inline uint64_t acc_words(uint32_t C, int nranks) { return (uint64_t)ACC_PER_CAND * C + CORP_FIXED + 18ull * nranks; }

constexpr double FX_SCALE = 4503599627370496.0;        // 2^52
constexpr double FX_INV   = 1.0 / 4503599627370496.0;  // 2^-52

// ---------------------------------------------------------------- 128-bit lane accumulator
struct Acc128 {
	unsigned long long lo; long long hi;
	__device__ __forceinline__ void zero() { lo = 0; hi = 0; }
	__device__ __forceinline__ void add(long long x) {
		const long long sx = x >> 63;
		asm("add.cc.u64 %0, %0, %2;\n\taddc.s64 %1, %1, %3;" : "+l"(lo), "+l"(hi) : "l"(x), "l"(sx));
	}
};

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }

// Warp-reduce a lane accumulator into three limb sums and add them to acc[0..2] (global atomics).
__device__ __forceinline__ void flush_acc128(const Acc128 &a, long long *dst, int lane) {
	const unsigned long long s0 = warp_sum_u64(a.lo & 0xffffffffull);
	const unsigned long long s1 = warp_sum_u64(a.lo >> 32);
	const unsigned long long s2 = warp_sum_u64((unsigned long long)a.hi);
	if (lane == 0) {
		if (s0) atomicAdd((unsigned long long *)dst + 0, s0);
		if (s1) atomicAdd((unsigned long long *)dst + 1, s1);
		if (s2) atomicAdd((unsigned long long *)dst + 2, s2);
	}
}

// limbs -> binary64 (one rounding of the 128-bit magnitude, then exact scaling by 2^-52)
__device__ inline double limbs_to_double(const long long *l) {
	__int128 v = ((__int128)l[2] << 64) + ((__int128)l[1] << 32) + (__int128)l[0];
	const bool neg = v < 0;
	unsigned __int128 m = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
	const unsigned long long hi = (unsigned long long)(m >> 64), lo = (unsigned long long)m;
	double d;
	if (hi == 0) d = (double)lo;
	else {
		// keep 64 significant bits + sticky, then a single correctly rounded conversion
		const int sh = 64 - __clzll((long long)hi);
		unsigned long long top = (unsigned long long)(m >> sh);
		const bool sticky = (m & ((((unsigned __int128)1) << sh) - 1)) != 0;
		top |= sticky ? 1ull : 0ull;
		d = ldexp((double)top, sh);
	}
	d *= FX_INV;
	return neg ? -d : d;
}

// a / b for integer-valued binary64 operands, 0 <= a, 1 <= b <= 2^32, correctly rounded and
// branch-free (the generic IEEE division carries a special-case branch that serialises
// independent evaluations).  y0 = rcp.approx (rel. error <= 2^-23), two Newton steps in FMA
// arithmetic leave y within 2^-104 of 1/b before the final rounding; 1/b for an integer b with
// <= 33 significant bits is never closer than 2^-86 (relative) to a rounding boundary unless it
// is exact, hence y == RN(1/b).  Then q = RN(a*y), r = a - b*q (exact), q' = RN(q + r*y) is the
// correctly rounded quotient (Markstein; the significand of b is never all ones).
__device__ __forceinline__ double div_small_int(double a, double b) {
	double y;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
	double e = __fma_rn(-b, y, 1.0);
	e = __fma_rn(e, e, e);
	y = __fma_rn(y, e, y);
	e = __fma_rn(-b, y, 1.0);
	y = __fma_rn(y, e, y);
	const double q = __dmul_rn(a, y);
	const double r = __fma_rn(-b, q, a);
	return __fma_rn(r, y, q);
}

// ---------------------------------------------------------------- TCS:668-763 on device
// p ? x : +0.0 as a bit mask, so the compiler keeps x unconditional instead of branching around it
__device__ __forceinline__ double keep_if(bool p, double x) {
	return __longlong_as_double(__double_as_longlong(x) & (p ? -1ll : 0ll));
}

// What the path asks of totalToolDurationMs (TCS:721-728, APO:754): taken from the record's durClass when the encoder decided
// it in binary64 (APO_DC_SET), else derived from the binary32 copy.  `avg > thr` is evaluated as `dur > thr * total` (exact
// product).
// TRUST: the record is resident — every load path (uploads, streaming windows, generators) has run fill_dur_class over it,
// so the class is always present and the streaming kernels (K1r, the corpus scan) spend no float compares on the duration.
struct DurClass { uint32_t level; bool pos, slow; };
__device__ __forceinline__ uint32_t dur_class_from_float(double dur, double total, bool tool) {
	const uint32_t lv = (dur > __dmul_rn(1000.0, total)) + (dur > __dmul_rn(3000.0, total)) + (dur > __dmul_rn(10000.0, total));
	return APO_DC_SET | ((tool && dur > 0.0) ? lv : 0u) | (dur > 0.0 ? APO_DC_POS : 0u) | (dur > 15000.0 ? APO_DC_SLOW : 0u);
}
template <bool TRUST = false>
__device__ __forceinline__ DurClass dur_class(const apo_record &r, double dur, double total) {
	uint32_t dcb = r.durClass;
	if (!TRUST && !(dcb & APO_DC_SET)) dcb = dur_class_from_float(dur, total, r.toolCalls > 0);
	DurClass d;
	d.level = dcb & APO_DC_LEVEL;
	d.pos = (dcb & APO_DC_POS) != 0;
	d.slow = (dcb & APO_DC_SLOW) != 0;
	return d;
}

// dims[i] is the pushed value where bit i of the returned mask is set and +0.0 elsewhere.
// Written select-style (no data-dependent branches) so that independent evaluations of one
// thread interleave and warps do not diverge; every conditional push of the reference becomes a
// mask bit.
__device__ __forceinline__ uint32_t reward_dims(const apo_record &r, double dims[APO_NDIM]) {
	const bool agent = r.mode == 2;                                   // TCS:673-674
	const bool err = (r.flags & APO_F_ERRORS) != 0;
	const bool ended = (r.flags & APO_F_ENDED) != 0;
	dims[0] = r.feedback == 1 ? 1.0 : (r.feedback == 2 ? -1.0 : 0.0); // TCS:677-678
	double comp = (ended && !err) ? 0.8 : 0.5;                        // TCS:682-685
	comp = err ? -0.5 : comp;                                         // TCS:686-688
	comp = r.feedback == 1 ? 1.0 : comp;                              // TCS:689-691
	dims[1] = comp;

	const bool tool = r.toolCalls > 0;                                // TCS:695
	const double total = (double)(tool ? r.toolCalls : 1u);
	const double rate = div_small_int((double)r.toolSucc, total);     // TCS:697 (correctly rounded quotient)
	dims[2] = keep_if(tool, __dadd_rn(__dmul_rn(rate, 2.0), -1.0));   // TCS:698
	const uint32_t sev = agent ? 5u : 3u, mod = agent ? 3u : 2u, mnr = agent ? 2u : 1u;          // TCS:702-704
	const double rel = r.toolFail >= sev ? -1.0 : (r.toolFail >= mod ? -0.5 : (r.toolFail >= mnr ? -0.2 : 1.0));
	dims[3] = keep_if(tool, rel);
	const uint32_t cexc = agent ? 8u : 3u, cgood = agent ? 15u : 6u, cfair = agent ? 25u : 10u;  // TCS:711-713
	const double eff = r.toolCalls > cfair ? -0.8 : (r.toolCalls > cgood ? -0.3 : (r.toolCalls > cexc ? 0.3 : 1.0));
	dims[4] = keep_if(tool, eff);
	// TCS:721-728.  avg = dur / total; `avg > thr`  <=>  dur > thr * total (exact product): a quotient of
	// an fp32 value by an integer <= 2^32 cannot lie in (thr, thr + ulp/2].
	const double dur = (double)r.toolDurMs;
	const DurClass dc = dur_class(r, dur, total);
	const bool hasdur = tool && dc.pos;
	const double ds = dc.level == 3u ? -0.5 : (dc.level == 2u ? 0.0 : (dc.level == 1u ? 0.5 : 1.0));
	dims[5] = keep_if(hasdur, ds);

	const bool llm = r.llmCalls > 0;                                  // TCS:733-736
	double over = __dadd_rn((double)r.llmCalls, agent ? -3.0 : -1.0);
	over = over > 0.0 ? over : 0.0;
	double ef = __dadd_rn(1.0, -__dmul_rn(over, 0.4));
	ef = ef < -1.0 ? -1.0 : ef;
	dims[6] = keep_if(llm, ef);

	const bool tok = r.tokens > 0;                                    // TCS:740-748
	const uint32_t texc = agent ? 5000u : 2000u, tgood = agent ? 15000u : 5000u, tfair = agent ? 30000u : 10000u;
	const double tks = r.tokens > tfair ? -0.5 : (r.tokens > tgood ? 0.0 : (r.tokens > texc ? 0.5 : 1.0));
	dims[7] = keep_if(tok, tks);

	const uint32_t turns = r.userMsgs < r.asstMsgs ? r.userMsgs : r.asstMsgs;   // TCS:752-754
	const bool conv = turns > 0;                                      // TCS:755-762
	const uint32_t thr = agent ? 3u : 2u;
	const double cs = turns > thr * 3 ? -0.8 : (turns > thr * 2 ? -0.3 : (turns > thr ? 0.3 : 1.0));
	dims[8] = keep_if(conv, cs);

	return 3u | (tool ? 0x1cu : 0u) | (hasdur ? 0x20u : 0u) | (llm ? 0x40u : 0u) | (tok ? 0x80u : 0u) | (conv ? 0x100u : 0u);
}

// ---------------------------------------------------------------- categorical product table
// Seven of the nine dimensions take one of <= 4 values (TCS:677-761), so value*weight is one of
// a handful of binary64 products.  The host tabulates them with the same IEEE multiplications
// (apo_abi.cu build_luts); K1r then adds table entries in push order instead of running select
// chains + DMULs.  The last slot of every group is +0.0 = "dimension not pushed".
// (offsets CAT_* / DIR_* / CAT_WORDS: apo_kernels.h, shared with the host table builder)

struct CatIdx { uint32_t i01, i3, i4, i5, i6, i7, i8; double d2; };   // effective slots (last slot of a group = absent)

// TCS:668-783 for one record through the product table: weighted sum in push order + mask.
template <bool WANT_IDX, bool TRUST = false>
__device__ __forceinline__ uint32_t record_ws_table_t(const apo_record &r, double w2, const double *cat, double &ws_out, CatIdx &ix) {
	const bool agent = r.mode == 2;
	const uint32_t err = (r.flags & APO_F_ERRORS) ? 1u : 0u, ended = (r.flags & APO_F_ENDED) ? 1u : 0u;
	const uint32_t i01 = (r.feedback < 3 ? r.feedback : 0u) + 3u * err + 6u * ended;
	double ws = cat[CAT_D01 + i01];
	const bool tool = r.toolCalls > 0;
	const double total = (double)(tool ? r.toolCalls : 1u);
	const double rate = div_small_int((double)r.toolSucc, total);
	const double d2 = keep_if(tool, __dadd_rn(__dmul_rn(rate, 2.0), -1.0));
	ws = __dadd_rn(ws, __dmul_rn(d2, w2));
	const uint32_t sev = agent ? 5u : 3u, mod = agent ? 3u : 2u, mnr = agent ? 2u : 1u;
	const uint32_t i3 = (r.toolFail >= mnr) + (r.toolFail >= mod) + (r.toolFail >= sev);
	ws = __dadd_rn(ws, cat[CAT_D3 + (tool ? i3 : 4u)]);
	const uint32_t cexc = agent ? 8u : 3u, cgood = agent ? 15u : 6u, cfair = agent ? 25u : 10u;
	const uint32_t i4 = (r.toolCalls > cexc) + (r.toolCalls > cgood) + (r.toolCalls > cfair);
	ws = __dadd_rn(ws, cat[CAT_D4 + (tool ? i4 : 4u)]);
	const double dur = (double)r.toolDurMs;
	const DurClass dc = dur_class<TRUST>(r, dur, total);
	const bool hasdur = tool && dc.pos;
	const uint32_t i5 = dc.level;
	ws = __dadd_rn(ws, cat[CAT_D5 + (hasdur ? i5 : 4u)]);
	const bool llm = r.llmCalls > 0;
	const uint32_t thr6 = agent ? 3u : 1u;
	const uint32_t over = r.llmCalls > thr6 ? r.llmCalls - thr6 : 0u;
	ws = __dadd_rn(ws, cat[CAT_D6 + (llm ? (over < 5u ? over : 5u) : 6u)]);
	const bool tok = r.tokens > 0;
	const uint32_t texc = agent ? 5000u : 2000u, tgood = agent ? 15000u : 5000u, tfair = agent ? 30000u : 10000u;
	const uint32_t i7 = (r.tokens > texc) + (r.tokens > tgood) + (r.tokens > tfair);
	ws = __dadd_rn(ws, cat[CAT_D7 + (tok ? i7 : 4u)]);
	const uint32_t turns = r.userMsgs < r.asstMsgs ? r.userMsgs : r.asstMsgs;
	const bool conv = turns > 0;
	const uint32_t thr8 = agent ? 3u : 2u;
	const uint32_t i8 = (turns > thr8) + (turns > thr8 * 2) + (turns > thr8 * 3);
	ws = __dadd_rn(ws, cat[CAT_D8 + (conv ? i8 : 4u)]);
	ws_out = ws;
	if (WANT_IDX) {
		ix.i01 = i01; ix.i3 = tool ? i3 : 4u; ix.i4 = tool ? i4 : 4u; ix.i5 = hasdur ? i5 : 4u;
		ix.i6 = llm ? (over < 5u ? over : 5u) : 6u; ix.i7 = tok ? i7 : 4u; ix.i8 = conv ? i8 : 4u; ix.d2 = d2;
	}
	return 3u | (tool ? 0x1cu : 0u) | (hasdur ? 0x20u : 0u) | (llm ? 0x40u : 0u) | (tok ? 0x80u : 0u) | (conv ? 0x100u : 0u);
}
__device__ __forceinline__ uint32_t record_ws_table(const apo_record &r, double w2, const double *cat, double &ws_out) {
	CatIdx ix;
	return record_ws_table_t<false>(r, w2, cat, ws_out, ix);
}

// Same weighted sum through the direct tables (K1r's inner loop): identical table values added in the
// same push order, about a quarter fewer instructions than the threshold form above.
__device__ __forceinline__ uint32_t record_ws_direct(const apo_record &r, double w2, const double *cat, double &ws_out) {
	const uint32_t ag = r.mode == 2 ? 1u : 0u;
	const uint32_t err = (r.flags & APO_F_ERRORS) ? 1u : 0u, ended = (r.flags & APO_F_ENDED) ? 1u : 0u;
	const uint32_t i01 = (r.feedback < 3 ? r.feedback : 0u) + 3u * err + 6u * ended;
	double ws = cat[CAT_D01 + i01];
	const bool tool = r.toolCalls > 0;
	const double total = (double)(tool ? r.toolCalls : 1u);
	const double rate = div_small_int((double)r.toolSucc, total);
	const double d2 = keep_if(tool, __dadd_rn(__dmul_rn(rate, 2.0), -1.0));
	ws = __dadd_rn(ws, __dmul_rn(d2, w2));
	ws = __dadd_rn(ws, cat[DIR_D3 + ag * 7u + (tool ? min(r.toolFail, 5u) : 6u)]);
	ws = __dadd_rn(ws, cat[DIR_D4 + ag * 27u + min(r.toolCalls, 26u)]);
	const double dur = (double)r.toolDurMs;
	const DurClass dc = dur_class<true>(r, dur, total);        // resident records always carry the class (fill_dur_class)
	const bool hasdur = tool && dc.pos;
	const uint32_t i5 = dc.level;
	ws = __dadd_rn(ws, cat[CAT_D5 + (hasdur ? i5 : 4u)]);
	ws = __dadd_rn(ws, cat[DIR_D6 + ag * 10u + min(r.llmCalls, 9u)]);
	const bool tok = r.tokens > 0;
	const uint32_t texc = ag ? 5000u : 2000u, tgood = ag ? 15000u : 5000u, tfair = ag ? 30000u : 10000u;
	const uint32_t i7 = (r.tokens > texc) + (r.tokens > tgood) + (r.tokens > tfair);
	ws = __dadd_rn(ws, cat[CAT_D7 + (tok ? i7 : 4u)]);
	const uint32_t turns = r.userMsgs < r.asstMsgs ? r.userMsgs : r.asstMsgs;
	ws = __dadd_rn(ws, cat[DIR_D8 + ag * 11u + min(turns, 10u)]);
	ws_out = ws;
	return 3u | (tool ? 0x1cu : 0u) | (hasdur ? 0x20u : 0u) | (r.llmCalls ? 0x40u : 0u) | (tok ? 0x80u : 0u) | (turns ? 0x100u : 0u);
}

// TCS:777-784.  lut[mask] = sum of the present weights added in push order (host-built,
// same sequential binary64 adds); entries whose total weight is 0 hold -1 (finalReward null).
// dims of absent dimensions must be +0.0 (adding +0.0*w leaves the running sum unchanged).
template <bool RECIP>
__device__ __forceinline__ double final_reward(const double dims[APO_NDIM], uint32_t mask, const Weights &W,
                                               const double *lut) {
	double ws = 0.0;
#pragma unroll
	for (int i = 0; i < APO_NDIM; i++) ws = __dadd_rn(ws, __dmul_rn(dims[i], W.w[i]));
	const double tw = lut[mask];
	return RECIP ? __dmul_rn(ws, tw) : __ddiv_rn(ws, tw);
}

__device__ __forceinline__ long long to_fx(double v) { return __double2ll_rn(v * FX_SCALE); }

// ws / tw, correctly rounded, without the generic division's special-case branch (which
// would serialise the four evaluations a thread interleaves): y = RN(1/tw) comes from the
// LUT, q = RN(ws*y), r = ws - tw*q (exact, one FMA), result = RN(q + r*y).  By Markstein's
// theorem this is the correctly rounded quotient when y is the correctly rounded reciprocal,
// the significand of tw is not all ones, and nothing underflows — tw is one of <= 511 sums of
// validated weights (apo_set_weights: 0 or within [1e-100,1e100]) and |ws| is 0 or >= 2^-160
// (fp32 inputs x weights).  Masks whose total weight has an all-ones significand (e.g. weights
// that sum to 0.9999999999999999) are flagged by the host with a NEGATIVE reciprocal and take
// the generic IEEE division instead.
template <bool RECIP>
__device__ __forceinline__ double div_fast(double ws, double2 t) {
	const double y = fabs(t.y);
	const double q = __dmul_rn(ws, y);
	if (RECIP) return q;
	const double r = __fma_rn(-t.x, q, ws);
	return __fma_rn(r, y, q);
}
template <bool RECIP>
__device__ __forceinline__ double div_lut(double ws, double2 t) {
	if (!RECIP && t.y < 0.0) return __ddiv_rn(ws, t.x);
	return div_fast<RECIP>(ws, t);
}

// The shared-memory LUT is indexed by a bit-rotated presence mask: dims 5..8 (the ones that
// vary independently in real data; d0,d1 are always present and d2..d4 come together) land
// in the low bits, so lanes with different masks mostly hit different banks.
__host__ __device__ constexpr int lut_bit(int dim) { return (dim + 4) % APO_NDIM; }
__device__ __forceinline__ uint32_t lut_index(uint32_t natural_mask) {
	return ((natural_mask >> 5) | (natural_mask << 4)) & 511u;
}


// ---------------------------------------------------------------- synthetic generator
// Build-defined; integer-only so host restatements agree bit for bit.
// Spec: DESIGN.md "Generator".
constexpr unsigned long long GOLD = 0x9E3779B97F4A7C15ull;
constexpr uint32_t STREAM_CORPUS = 1, STREAM_ROLLOUT = 2;

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
	z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
	z ^= z >> 27; z *= 0x94D049BB133111EBull;
	z ^= z >> 31;
	return z;
}
__device__ __forceinline__ uint32_t ctz64(unsigned long long x) { return (uint32_t)(__ffsll((long long)x) - 1); }

// key0 = mix64(mix64(seed ^ stream*GOLD) ^ mix64((c+1)*GOLD)) is hoisted by callers.
__device__ __forceinline__ unsigned long long gen_key(unsigned long long seed, uint32_t stream, uint32_t c) {
	return mix64(mix64(seed ^ ((unsigned long long)stream * GOLD)) ^ mix64(((unsigned long long)c + 1) * GOLD));
}
__device__ __forceinline__ uint32_t gen_quality(unsigned long long seed, uint32_t stream, uint32_t c) {
	return stream == STREAM_CORPUS ? 512u : (uint32_t)(mix64(seed ^ 0xC0FFEEull ^ ((unsigned long long)c * GOLD)) & 1023);
}

__device__ __forceinline__ apo_record gen_record(unsigned long long key, uint32_t qc, unsigned long long t,
                                                 uint32_t agent_permille) {
	const unsigned long long base = key + t * GOLD;
	const unsigned long long h1 = mix64(base + GOLD), h2 = mix64(base + 2 * GOLD);
	const unsigned long long h3 = mix64(base + 3 * GOLD), h4 = mix64(base + 4 * GOLD);
	const uint32_t r = (uint32_t)(h1 & 1023);
	const uint32_t pgood = 192 + (qc >> 2), pbad = 256 - (qc >> 3);
	const uint32_t feedback = r < pgood ? 1u : (r < pgood + pbad ? 2u : 0u);
	const bool err = ((h1 >> 10) & 1023) < 102;
	const bool ended = ((h1 >> 20) & 1023) < 973;
	uint32_t mode;
	if (((h1 >> 30) & 1023) < agent_permille) mode = 2;
	else {
		const uint32_t k = (uint32_t)((h1 >> 40) & 15);
		mode = k < 10 ? 1u : (k < 12 ? 3u : (k < 14 ? 4u : 0u));
	}
	const bool agent = mode == 2;
	uint32_t tool = 0, fail = 0;
	float dur = 0.0f;
	if (!(((h1 >> 44) & 1023) < 307)) {
		const uint32_t g = ctz64(h2 | (1ull << 20));
		tool = agent ? 1 + (uint32_t)((h2 >> 21) & 15) + g : 1 + g;
		const uint32_t n = tool < 16 ? tool : 16;
		for (uint32_t i = 0; i < n; i++) fail += (((h3 >> (4 * i)) & 15) == 0) ? 1u : 0u;
		if (((h2 >> 25) & 15) != 0) {
			const uint32_t e = (uint32_t)((h2 >> 29) & 15) % 10;
			const uint32_t basems = 50u << e;
			const uint32_t frac = (uint32_t)((h2 >> 33) & 4095);
			const uint32_t avg = basems + ((basems * frac) >> 12);
			dur = (float)(avg * tool);
		}
	}
	uint32_t llm = 0;
	if (((h2 >> 45) & 31) != 0) llm = 1 + ctz64((h2 >> 50) | (1ull << 13)) + (agent ? (uint32_t)((h2 >> 48) & 3) : 0u);
	uint32_t tokens = 0;
	if (!((h4 & 1023) < 205)) {
		const uint32_t e = (uint32_t)((h4 >> 10) & 15) % 9;
		const uint32_t b = 200u << e;
		const uint32_t f = (uint32_t)((h4 >> 14) & 4095);
		tokens = b + ((b * f) >> 12);
	}
	uint32_t user = 0;
	if (((h4 >> 40) & 63) != 0) user = 1 + ctz64((h4 >> 26) | (1ull << 12));
	apo_record o;
	o.feedback = (uint8_t)feedback;
	o.flags = (uint8_t)((err ? APO_F_ERRORS : 0u) | (ended ? APO_F_ENDED : 0u) |
	                    ((ended || feedback) ? APO_F_VALID : 0u) | (fail > 0 ? APO_F_FAILSPAN : 0u));
	o.mode = (uint8_t)mode;
	// generator durations are integers < 2^24 (exact in binary32, dur = avg * tool): the class follows from the integers
	{
		const uint32_t dms = (uint32_t)dur, avg = tool ? dms / tool : 0u;
		o.durClass = (uint8_t)(APO_DC_SET | (dms > 0u ? APO_DC_POS : 0u) | (dms > 15000u ? APO_DC_SLOW : 0u) |
		                       ((avg > 1000u) + (avg > 3000u) + (avg > 10000u)));
	}
	o.userMsgs = (uint16_t)user;
	const uint32_t asst = llm + ((((h4 >> 46) & 15) == 0) ? 1u : 0u);
	o.asstMsgs = (uint16_t)(asst < 65535u ? asst : 65535u);
	o.toolCalls = tool; o.toolSucc = tool - fail; o.toolFail = fail;
	o.llmCalls = llm; o.tokens = tokens; o.toolDurMs = dur;
	return o;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize once per (kernel, device): the driver call costs a few microseconds,
// which a 4 x 1000 call would pay on every launch.  Keyed by the kernel's address (instantiations share pointer types).
inline cudaError_t allow_big_smem_impl(const void *kernel, int bytes) {
	struct Entry { const void *k; int dev; };
	static Entry table[256];
	static int n = 0;
	static volatile int lock = 0;
	int dev = 0;
	cudaGetDevice(&dev);
	while (__sync_lock_test_and_set(&lock, 1)) {}
	bool known = false;
	for (int i = 0; i < n; i++) if (table[i].k == kernel && table[i].dev == dev) { known = true; break; }
	__sync_lock_release(&lock);
	if (known) return cudaSuccess;
	const cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	if (err == cudaSuccess) {
		while (__sync_lock_test_and_set(&lock, 1)) {}
		if (n < 256) { table[n].k = kernel; table[n].dev = dev; n++; }
		__sync_lock_release(&lock);
	}
	return err;
}
template <class K>
inline cudaError_t allow_big_smem(K kernel, int bytes) { return allow_big_smem_impl(reinterpret_cast<const void *>(kernel), bytes); }

// ---------------------------------------------------------------- mbarrier / bulk-copy PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "WAIT_%=:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	    "@p bra DONE_%=;\n\t"
	    "bra WAIT_%=;\n\t"
	    "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
	asm volatile(
	    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
	        smem_u32(dst)),
	    "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
	    : "memory");
}

__device__ __forceinline__ void bulk_g2s_nohint(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

}  // namespace apo
