/*
 * apo_oracle.h — CPU ORACLE for the SenWeaver APO scoring hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (senweaver-ide_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * PARITY PIN: the reference (TypeScript, VS Code fork) ships no tests, golden vectors or fixtures
 * for this path and no JS runtime exists in this environment, so its OWN method texts (`_computeRewardSignals`,
 * `getStats`, `_buildReport`, `_analyzePatterns`, `_generateLocalSuggestions`) are executed unmodified by the in-repo
 * interpreter oracle/ts_harness/minijs.py (run_reference.py; run_reference.mjs does the same under Node): the outputs
 * are the fixtures tests/golden/ref_*.json, and tests/test_reference_pin.py holds this file to them bit for bit
 * (409 traces, 11 corpora).  The beam top-K alone stays defined by SURVEY 8c: the reference performs it on a closed
 * backend.  This file is a plain-C binary64 restatement of the reference *source text*,
 * also cross-checked against an independent pure-Python transcription
 * (oracle/ts_transcription.py) and the hand-derived known-answer vectors K1..K7.
 *
 * Citations: TCS = src/vs/workbench/contrib/senweaver/common/traceCollectorService.ts
 *            APO = src/vs/workbench/contrib/senweaver/common/apoService.ts
 * (paths relative to the reference checkout).
 */
#ifndef APO_ORACLE_H
#define APO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NDIM 9
#define ORC_NPAT 6
#define ORC_NMODE 5

/* Form R: one trace-summary record (the fields of ConversationTrace.summary that the
 * hot path reads, TCS:94-108, plus endTime TCS:87, metadata.chatMode TCS:91 and the two
 * span-type counts used at TCS:752-753 / APO:733).  32 bytes, little endian. */
typedef struct {
	uint8_t  feedback;   /* 0 null, 1 'good', 2 'bad'                      (TCS:98)  */
	uint8_t  flags;      /* bit0 hasErrors (TCS:99), bit1 endTime set (TCS:87),
	                        bit3 finalReward!==null (TCS:106, "valid"),
	                        bit4 any tool_call span with toolSuccess===false (APO:667-669) */
	uint8_t  mode;       /* 0 no metadata.chatMode, 1 'normal', 2 'agent', 3 'gather', 4 'designer' */
	uint8_t  durClass;   /* 0x80 set | 0x03 level | 0x04 duration > 0 | 0x08 duration > 15000: the binary64 comparisons of
	                        TCS:721-728 / APO:754 made by the encoder; 0 = derive them from toolDurMs */
	uint16_t userMsgs;   /* # spans of type user_message      (TCS:752, APO:733) */
	uint16_t asstMsgs;   /* # spans of type assistant_message (TCS:753)          */
	uint32_t toolCalls;  /* totalToolCalls      (TCS:96)  */
	uint32_t toolSucc;   /* toolCallsSucceeded  (TCS:101) */
	uint32_t toolFail;   /* toolCallsFailed     (TCS:102) */
	uint32_t llmCalls;   /* totalLLMCalls       (TCS:95)  */
	uint32_t tokens;     /* totalTokens         (TCS:97)  */
	float    toolDurMs;  /* totalToolDurationMs (TCS:104) */
} orc_record;

#define ORC_F_ERRORS   0x01u
#define ORC_F_ENDED    0x02u
#define ORC_F_VALID    0x08u
#define ORC_F_FAILSPAN 0x10u

/* TCS:766-776 weights in push order d0..d8. */
extern const double orc_weights[ORC_NDIM];

/* TCS:668-763: raw record -> up to 9 dims.  dims[i] is NaN when dim i is not pushed.
 * Returns the 9-bit presence mask (bit i = dim i pushed). */
uint32_t orc_reward_dims(const orc_record *r, double dims[ORC_NDIM]);

/* TCS:777-784: weighted mean over the present dims in push order.
 * Returns 1 and *out when totalWeight>0, 0 (finalReward === null) otherwise. */
int orc_final_reward(const double dims[ORC_NDIM], uint32_t mask, const double w[ORC_NDIM], double *out);

/* Form D row (9 x fp32, NaN = absent): widen, derive mask, TCS:777-784. */
int orc_final_reward_f32(const float row[ORC_NDIM], const double w[ORC_NDIM], double *out);

/* APO:550-553 applied per candidate: score[c] = (sum_t finalReward[c,t]) / n over
 * non-null, sequential in t.  dims is [C][pitch_evals][9] fp32, first T evals of each
 * candidate row used.  score = -inf when n == 0. */
void orc_score_dims(const float *dims, uint32_t C, uint64_t T, uint64_t pitch_evals,
                    const double w[ORC_NDIM], double *scores, uint64_t *counts);

/* Same but per-(c,t) Form R input: dims derived by TCS:668-763 first.  Records whose
 * VALID flag is clear have finalReward === null and are skipped (TCS:606, APO:550). */
void orc_score_records(const orc_record *recs, uint32_t C, uint64_t T, uint64_t pitch,
                       const double w[ORC_NDIM], double *scores, uint64_t *counts);

/* Exact integer restatement of the engine's accumulator: per candidate
 * sum_t rint(finalReward[c,t] * 2^52) as a 128-bit integer (lo, hi) + the non-null count.
 * rint = round-half-even of the binary64 product (exact scaling by a power of two).
 * Lets tests compare the GPU partial sums with NO tolerance. */
void orc_score_dims_fx(const float *dims, uint32_t C, uint64_t T, uint64_t pitch_evals,
                       const double w[ORC_NDIM], uint64_t *lo, int64_t *hi, uint64_t *counts);
void orc_score_records_fx(const orc_record *recs, uint32_t C, uint64_t T, uint64_t pitch,
                          const double w[ORC_NDIM], uint64_t *lo, int64_t *hi, uint64_t *counts);

/* Multi-threaded variants for the CPU baseline: contiguous T-slices per thread,
 * per-slice sequential partials merged in slice order. */
void orc_score_dims_mt(const float *dims, uint32_t C, uint64_t T, uint64_t pitch_evals,
                       const double w[ORC_NDIM], double *scores, uint64_t *counts, int nthreads);
void orc_score_records_mt(const orc_record *recs, uint32_t C, uint64_t T, uint64_t pitch,
                          const double w[ORC_NDIM], double *scores, uint64_t *counts, int nthreads);

/* SURVEY 8c top-K spec: order by score descending, ties -> lower index (stable sort,
 * cf. APO:1006, 1482), first K.  best adoption rule APO:1159 is strict '>'. */
void orc_topk(const double *scores, uint32_t C, uint32_t K, int32_t *out_idx);

typedef struct {
	uint64_t count;        /* frequency (APO:649 ...)                    */
	uint8_t  flag;         /* pattern emitted (count >= min)             */
	uint8_t  severity;     /* 0 low, 1 medium, 2 high                    */
	int64_t  examples[3];  /* first 3 matching record indices, -1 = none */
} orc_pattern;

typedef struct {
	double   sum; uint64_t count; double avg;   /* APO:556-568 */
	uint8_t  low_flag;      /* APO:575: avg < -0.3 && count >= 5 */
	uint8_t  low_severity;  /* APO:591: avg < -0.5 ? high(2) : medium(1) */
	uint8_t  sugg_flag;     /* APO:802: avg < 0 && count >= 3 */
	uint8_t  sugg_priority; /* APO:819: avg < -0.5 ? high(2) : medium(1) */
} orc_dimstat;

typedef struct {
	uint64_t total, good, bad, none;            /* APO:509-516 */
	double   goodRate;                          /* APO:546-547 */
	uint64_t byMode[ORC_NMODE][3];              /* total, good, bad per mode code (APO:519-525) */
	double   byModeGoodRate[ORC_NMODE];         /* APO:541-544 */
	uint64_t withReward;                        /* APO:550 */
	double   rewardSum;                         /* sequential (APO:552) */
	double   avgReward;                         /* NaN when withReward==0 (null) */
	orc_dimstat dim[ORC_NDIM];
	orc_pattern pat[ORC_NPAT];                  /* APO:635-773 */
	uint64_t toolCalls, toolSucc, toolFail;     /* TCS:603-605 */
	double   toolSuccessRate;                   /* TCS:624, NaN = null */
} orc_report;

/* APO:498-625 numeric content + TCS:596-626 stats.  idx_base is added to example
 * indices (global index of records[0]). */
void orc_report_build(const orc_record *recs, uint64_t T, uint64_t idx_base,
                      const double w[ORC_NDIM], orc_report *out);
/* The same on all host threads (persistent pool): integers and first-3 examples identical to
 * orc_report_build, binary64 sums merged per slice (last bits differ from the sequential sums).
 * orc_report_generated walks records [t0, t0+T) of the generator's corpus stream without
 * materialising them. */
void orc_report_build_mt(const orc_record *recs, uint64_t T, uint64_t idx_base,
                         const double w[ORC_NDIM], orc_report *out, int nthreads);
void orc_report_generated(uint64_t seed, uint64_t t0, uint64_t T, uint64_t idx_base, uint32_t agent_permille,
                          const double w[ORC_NDIM], orc_report *out, int nthreads);
/* Exact integer sums (as orc_score_dims_fx) for the listed candidates over generated Form D
 * evaluations t0 .. t0+T-1, never materialised; any thread count gives the same integers. */
void orc_score_generated_fx(uint64_t seed, const uint32_t *cands, uint32_t ncand, uint64_t t0, uint64_t T,
                            uint32_t agent_permille, const double w[ORC_NDIM], uint64_t *lo, int64_t *hi,
                            uint64_t *counts, int nthreads);

/* Form R16 (include/apo_b200.h apo_record16): independent restatement of the 16-byte unpacking. */
typedef struct {
	uint16_t hdr; uint8_t userMsgs, asstMsgs; uint16_t toolCalls, toolFail; uint8_t llmCalls, durClass; uint16_t tokens; float toolDurMs;
} orc_record16;
void orc_unpack16(const orc_record16 *in, uint64_t n, orc_record *out);

/* ---- synthetic generator (build-defined; spec in DESIGN.md "Generator") ---- */
#define ORC_STREAM_CORPUS  1u
#define ORC_STREAM_ROLLOUT 2u
void orc_gen_record(uint64_t seed, uint32_t stream, uint32_t c, uint64_t t,
                    uint32_t agent_permille, orc_record *out);
/* Form D row for (c,t): record -> dims (binary64) -> rounded once to fp32; all-NaN when invalid. */
void orc_gen_dims_row(uint64_t seed, uint32_t c, uint64_t t, uint32_t agent_permille, float out[ORC_NDIM]);
void orc_gen_dims(uint64_t seed, uint32_t c0, uint32_t C, uint64_t t0, uint64_t T, uint64_t pitch_evals,
                  uint32_t agent_permille, float *out, int nthreads);
void orc_gen_records(uint64_t seed, uint32_t stream, uint32_t c0, uint32_t C, uint64_t t0, uint64_t T,
                     uint64_t pitch, uint32_t agent_permille, orc_record *out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
