/*
 * apo_oracle.c — CPU ORACLE (test infrastructure only; see apo_oracle.h header).
 *
 * PARITY PINNED by tests/golden/ref_*.json: outputs of the reference's own method texts executed by
 * oracle/ts_harness (see apo_oracle.h); tests/test_reference_pin.py compares this file with them bit for bit.
 * Plain C, IEEE-754 binary64 throughout (JS `number`), compiled with
 * -ffp-contract=off so no multiply-add is fused: every `a*b + c` below rounds twice,
 * exactly like the TypeScript it restates.
 *
 * TCS = src/vs/workbench/contrib/senweaver/common/traceCollectorService.ts
 * APO = src/vs/workbench/contrib/senweaver/common/apoService.ts
 */
#define _GNU_SOURCE
#include "apo_oracle.h"
#include <math.h>
#include <sched.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* TCS:766-776, in the order the dims are pushed (TCS:679,692,698,708,718,728,736,748,762). */
const double orc_weights[ORC_NDIM] = {
	0.25, /* user_feedback            */
	0.18, /* task_completion          */
	0.12, /* tool_success_rate        */
	0.08, /* tool_call_reliability    */
	0.05, /* tool_call_efficiency     */
	0.05, /* tool_duration_efficiency */
	0.08, /* response_efficiency      */
	0.08, /* token_efficiency         */
	0.11, /* conversation_efficiency  */
};

/* ---------------------------------------------------------------- TCS:668-763 */
uint32_t orc_reward_dims(const orc_record *r, double dims[ORC_NDIM])
{
	uint32_t mask = 0;
	for (int i = 0; i < ORC_NDIM; i++) dims[i] = NAN;

	/* TCS:673-674: only the literal 'agent' selects the agent thresholds. */
	const int isAgent = (r->mode == 2);
	const int hasErrors = (r->flags & ORC_F_ERRORS) != 0;
	const int ended = (r->flags & ORC_F_ENDED) != 0;

	/* d0 user_feedback, TCS:677-679 */
	dims[0] = r->feedback == 1 ? 1.0 : (r->feedback == 2 ? -1.0 : 0.0);
	mask |= 1u << 0;

	/* d1 task_completion, TCS:682-692 — later rule wins */
	double completion = 0.5;
	if (ended && !hasErrors) completion = 0.8;
	if (hasErrors) completion = -0.5;
	if (r->feedback == 1) completion = 1.0;
	dims[1] = completion;
	mask |= 1u << 1;

	/* d2..d5, TCS:695-730 */
	if (r->toolCalls > 0) {
		const double total = (double)r->toolCalls;
		const double rate = (double)r->toolSucc / total;   /* TCS:697 */
		dims[2] = rate * 2 - 1;                              /* TCS:698 */
		mask |= 1u << 2;

		/* TCS:701-708, '>=' comparisons */
		const uint32_t sev = isAgent ? 5 : 3, mod = isAgent ? 3 : 2, mnr = isAgent ? 2 : 1;
		double pen = 1.0;
		if (r->toolFail >= sev) pen = -1.0;
		else if (r->toolFail >= mod) pen = -0.5;
		else if (r->toolFail >= mnr) pen = -0.2;
		dims[3] = pen;
		mask |= 1u << 3;

		/* TCS:711-718, strict '>' */
		const uint32_t exc = isAgent ? 8 : 3, good = isAgent ? 15 : 6, fair = isAgent ? 25 : 10;
		double cnt = 1.0;
		if (r->toolCalls > fair) cnt = -0.8;
		else if (r->toolCalls > good) cnt = -0.3;
		else if (r->toolCalls > exc) cnt = 0.3;
		dims[4] = cnt;
		mask |= 1u << 4;

		/* TCS:721-729.  The record carries totalToolDurationMs as binary32 plus, when the encoder held the JS double, the
		 * results of the binary64 comparisons (durClass); without them the comparisons are made here on the binary32 value. */
		const double dur = (double)r->toolDurMs;
		const int set = (r->durClass & 0x80) != 0;
		if (set ? (r->durClass & 0x04) != 0 : dur > 0) {
			double ds = 1.0;
			if (set) {
				const unsigned lv = r->durClass & 3u;
				ds = lv == 3 ? -0.5 : (lv == 2 ? 0.0 : (lv == 1 ? 0.5 : 1.0));
			} else {
				const double avg = dur / total;
				if (avg > 10000) ds = -0.5;
				else if (avg > 3000) ds = 0.0;
				else if (avg > 1000) ds = 0.5;
			}
			dims[5] = ds;
			mask |= 1u << 5;
		}
	}

	/* d6 response_efficiency, TCS:733-737 */
	if (r->llmCalls > 0) {
		const double thr = isAgent ? 3 : 1;
		double over = (double)r->llmCalls - thr;
		if (!(over > 0)) over = 0;                 /* Math.max(0, n - thr) */
		double e = 1 - over * 0.4;
		if (e < -1) e = -1;                        /* Math.max(-1, ...) */
		dims[6] = e;
		mask |= 1u << 6;
	}

	/* d7 token_efficiency, TCS:740-749 */
	if (r->tokens > 0) {
		const uint32_t exc = isAgent ? 5000 : 2000, good = isAgent ? 15000 : 5000, fair = isAgent ? 30000 : 10000;
		double ts = 1.0;
		if (r->tokens > fair) ts = -0.5;
		else if (r->tokens > good) ts = 0.0;
		else if (r->tokens > exc) ts = 0.5;
		dims[7] = ts;
		mask |= 1u << 7;
	}

	/* d8 conversation_efficiency, TCS:752-763 */
	const uint32_t turns = r->userMsgs < r->asstMsgs ? r->userMsgs : r->asstMsgs;
	if (turns > 0) {
		const uint32_t thr = isAgent ? 3 : 2;
		double tsc = 1.0;
		if (turns > thr * 3) tsc = -0.8;
		else if (turns > thr * 2) tsc = -0.3;
		else if (turns > thr) tsc = 0.3;
		dims[8] = tsc;
		mask |= 1u << 8;
	}
	return mask;
}

/* ---------------------------------------------------------------- TCS:777-784 */
int orc_final_reward(const double dims[ORC_NDIM], uint32_t mask, const double w[ORC_NDIM], double *out)
{
	double weightedSum = 0, totalWeight = 0;
	for (int i = 0; i < ORC_NDIM; i++) {
		if (!(mask & (1u << i))) continue;
		weightedSum += dims[i] * w[i];
		totalWeight += w[i];
	}
	if (!(totalWeight > 0)) return 0;
	*out = weightedSum / totalWeight;
	return 1;
}

int orc_final_reward_f32(const float row[ORC_NDIM], const double w[ORC_NDIM], double *out)
{
	double d[ORC_NDIM];
	uint32_t mask = 0;
	for (int i = 0; i < ORC_NDIM; i++) {
		d[i] = (double)row[i];
		if (!isnan(row[i])) mask |= 1u << i;
	}
	return orc_final_reward(d, mask, w, out);
}

static int record_final(const orc_record *r, const double w[ORC_NDIM], double *out)
{
	if (!(r->flags & ORC_F_VALID)) return 0;   /* finalReward === null, TCS:397 */
	double d[ORC_NDIM];
	uint32_t mask = orc_reward_dims(r, d);
	return orc_final_reward(d, mask, w, out);
}

/* ---------------------------------------------------------------- APO:550-553 per candidate */
static void score_dims_range(const float *row0, uint64_t t0, uint64_t t1, const double w[ORC_NDIM],
                             double *sum, uint64_t *n)
{
	double s = 0; uint64_t k = 0;
	for (uint64_t t = t0; t < t1; t++) {
		double fr;
		if (orc_final_reward_f32(row0 + t * ORC_NDIM, w, &fr)) { s += fr; k++; }
	}
	*sum = s; *n = k;
}

static void score_recs_range(const orc_record *row0, uint64_t t0, uint64_t t1, const double w[ORC_NDIM],
                             double *sum, uint64_t *n)
{
	double s = 0; uint64_t k = 0;
	for (uint64_t t = t0; t < t1; t++) {
		double fr;
		if (record_final(row0 + t, w, &fr)) { s += fr; k++; }
	}
	*sum = s; *n = k;
}

void orc_score_dims(const float *dims, uint32_t C, uint64_t T, uint64_t pitch_evals,
                    const double w[ORC_NDIM], double *scores, uint64_t *counts)
{
	for (uint32_t c = 0; c < C; c++) {
		double s; uint64_t n;
		score_dims_range(dims + (uint64_t)c * pitch_evals * ORC_NDIM, 0, T, w, &s, &n);
		counts[c] = n;
		scores[c] = n > 0 ? s / (double)n : -INFINITY;
	}
}

void orc_score_records(const orc_record *recs, uint32_t C, uint64_t T, uint64_t pitch,
                       const double w[ORC_NDIM], double *scores, uint64_t *counts)
{
	for (uint32_t c = 0; c < C; c++) {
		double s; uint64_t n;
		score_recs_range(recs + (uint64_t)c * pitch, 0, T, w, &s, &n);
		counts[c] = n;
		scores[c] = n > 0 ? s / (double)n : -INFINITY;
	}
}

/* ---- exact fixed-point sums (mirrors the engine's accumulator definition) ---- */
void orc_score_dims_fx(const float *dims, uint32_t C, uint64_t T, uint64_t pitch_evals,
                       const double w[ORC_NDIM], uint64_t *lo, int64_t *hi, uint64_t *counts)
{
	for (uint32_t c = 0; c < C; c++) {
		__int128 acc = 0; uint64_t n = 0;
		const float *row0 = dims + (uint64_t)c * pitch_evals * ORC_NDIM;
		for (uint64_t t = 0; t < T; t++) {
			double fr;
			if (orc_final_reward_f32(row0 + t * ORC_NDIM, w, &fr)) { acc += (__int128)llrint(fr * 4503599627370496.0); n++; }
		}
		lo[c] = (uint64_t)acc; hi[c] = (int64_t)(acc >> 64); counts[c] = n;
	}
}

void orc_score_records_fx(const orc_record *recs, uint32_t C, uint64_t T, uint64_t pitch,
                          const double w[ORC_NDIM], uint64_t *lo, int64_t *hi, uint64_t *counts)
{
	for (uint32_t c = 0; c < C; c++) {
		__int128 acc = 0; uint64_t n = 0;
		for (uint64_t t = 0; t < T; t++) {
			double fr;
			if (record_final(recs + (uint64_t)c * pitch + t, w, &fr)) { acc += (__int128)llrint(fr * 4503599627370496.0); n++; }
		}
		lo[c] = (uint64_t)acc; hi[c] = (int64_t)(acc >> 64); counts[c] = n;
	}
}

/* ---- persistent thread pool (timing legs of bench.py and the bench-side parity checks) ----
 * Workers are created once and parked on a condition variable; a run hands every participant
 * (tid, nthreads) and items are claimed through an atomic counter, so a busy or slow core does not
 * hold the others back.  Workers are pinned to the CPUs of the process's affinity mask (set ORC_PIN=0
 * to leave them floating): pages first touched by the multi-threaded generators stay local to the
 * thread that later reads them. */
typedef void (*pool_fn)(void *arg, int tid, int nthreads);
static struct {
	pthread_mutex_t mu, run_mu;
	pthread_cond_t go, done;
	pthread_t *th;
	int nworkers;              /* threads created so far (tids 1..nworkers) */
	uint64_t gen;
	int active, pending;
	pool_fn fn; void *arg;
} g_pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, NULL, 0, 0, 0, 0, NULL, NULL};

static void *pool_worker(void *p)
{
	const int tid = (int)(intptr_t)p;
	uint64_t seen = 0;
	pthread_mutex_lock(&g_pool.mu);
	for (;;) {
		while (g_pool.gen == seen) pthread_cond_wait(&g_pool.go, &g_pool.mu);
		seen = g_pool.gen;
		if (tid >= g_pool.active) continue;
		pool_fn fn = g_pool.fn; void *arg = g_pool.arg; const int n = g_pool.active;
		pthread_mutex_unlock(&g_pool.mu);
		fn(arg, tid, n);
		pthread_mutex_lock(&g_pool.mu);
		if (--g_pool.pending == 0) pthread_cond_signal(&g_pool.done);
	}
	return NULL;
}

static void pool_run(pool_fn fn, void *arg, int nthreads)
{
	if (nthreads <= 1) { fn(arg, 0, 1); return; }
	pthread_mutex_lock(&g_pool.run_mu);                       /* one run at a time */
	pthread_mutex_lock(&g_pool.mu);
	if (g_pool.nworkers < nthreads - 1) {
		cpu_set_t allowed; int ncpu = 0, cpus[1024];
		const char *pin = getenv("ORC_PIN");
		const int do_pin = !(pin && pin[0] == '0');
		if (do_pin && sched_getaffinity(0, sizeof allowed, &allowed) == 0)
			for (int c = 0; c < CPU_SETSIZE && ncpu < 1024; c++) if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
		g_pool.th = (pthread_t *)realloc(g_pool.th, sizeof(pthread_t) * (size_t)(nthreads - 1));
		while (g_pool.nworkers < nthreads - 1) {
			const int tid = g_pool.nworkers + 1;
			if (pthread_create(&g_pool.th[tid - 1], NULL, pool_worker, (void *)(intptr_t)tid) != 0) break;
			if (ncpu > 0) {
				cpu_set_t one; CPU_ZERO(&one); CPU_SET(cpus[tid % ncpu], &one);
				pthread_setaffinity_np(g_pool.th[tid - 1], sizeof one, &one);
			}
			g_pool.nworkers++;
		}
		if (nthreads - 1 > g_pool.nworkers) nthreads = g_pool.nworkers + 1;   /* thread limit reached: run narrower */
	}
	g_pool.fn = fn; g_pool.arg = arg; g_pool.active = nthreads; g_pool.pending = nthreads - 1;
	g_pool.gen++;
	pthread_cond_broadcast(&g_pool.go);
	pthread_mutex_unlock(&g_pool.mu);
	fn(arg, 0, nthreads);
	pthread_mutex_lock(&g_pool.mu);
	while (g_pool.pending > 0) pthread_cond_wait(&g_pool.done, &g_pool.mu);
	pthread_mutex_unlock(&g_pool.mu);
	pthread_mutex_unlock(&g_pool.run_mu);
}

/* ---- multi-threaded baseline: (candidate, T-slice) work items, merged in slice order ---- */
typedef struct {
	const float *dims; const orc_record *recs;
	uint32_t C; uint64_t T, pitch; const double *w;
	int nslice; double *psum; uint64_t *pcnt;
	uint64_t next;                     /* atomic item counter */
} mt_job;

static void mt_worker(void *arg, int tid, int nthreads)
{
	(void)tid; (void)nthreads;
	mt_job *j = (mt_job *)arg;
	const uint64_t nitems = (uint64_t)j->C * (uint64_t)j->nslice;
	for (;;) {
		/* slice-major claim order: neighbouring claims are the same T-slice of consecutive candidates */
		const uint64_t k = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
		if (k >= nitems) break;
		const int s = (int)(k / j->C);
		const uint32_t c = (uint32_t)(k % j->C);
		const uint64_t it = (uint64_t)c * (uint64_t)j->nslice + (uint64_t)s;
		const uint64_t t0 = j->T * (uint64_t)s / (uint64_t)j->nslice;
		const uint64_t t1 = j->T * (uint64_t)(s + 1) / (uint64_t)j->nslice;
		if (j->dims)
			score_dims_range(j->dims + (uint64_t)c * j->pitch * ORC_NDIM, t0, t1, j->w, &j->psum[it], &j->pcnt[it]);
		else
			score_recs_range(j->recs + (uint64_t)c * j->pitch, t0, t1, j->w, &j->psum[it], &j->pcnt[it]);
	}
}

static void score_mt(const float *dims, const orc_record *recs, uint32_t C, uint64_t T, uint64_t pitch,
                     const double w[ORC_NDIM], double *scores, uint64_t *counts, int nthreads)
{
	if (nthreads < 1) nthreads = 1;
	int nslice = nthreads;            /* every candidate is cut into nthreads T-slices */
	const uint64_t nitems = (uint64_t)C * (uint64_t)nslice;
	double *psum = (double *)calloc(nitems ? nitems : 1, sizeof(double));
	uint64_t *pcnt = (uint64_t *)calloc(nitems ? nitems : 1, sizeof(uint64_t));
	mt_job job = {dims, recs, C, T, pitch, w, nslice, psum, pcnt, 0};
	pool_run(mt_worker, &job, nthreads);
	for (uint32_t c = 0; c < C; c++) {
		double s = 0; uint64_t n = 0;
		for (int k = 0; k < nslice; k++) { s += psum[(uint64_t)c * nslice + k]; n += pcnt[(uint64_t)c * nslice + k]; }
		counts[c] = n;
		scores[c] = n > 0 ? s / (double)n : -INFINITY;
	}
	free(psum); free(pcnt);
}

void orc_score_dims_mt(const float *dims, uint32_t C, uint64_t T, uint64_t pitch_evals,
                       const double w[ORC_NDIM], double *scores, uint64_t *counts, int nthreads)
{ score_mt(dims, NULL, C, T, pitch_evals, w, scores, counts, nthreads); }

void orc_score_records_mt(const orc_record *recs, uint32_t C, uint64_t T, uint64_t pitch,
                          const double w[ORC_NDIM], double *scores, uint64_t *counts, int nthreads)
{ score_mt(NULL, recs, C, T, pitch, w, scores, counts, nthreads); }

/* ---------------------------------------------------------------- top-K */
void orc_topk(const double *scores, uint32_t C, uint32_t K, int32_t *out_idx)
{
	/* stable insertion-selection: K passes, strict '>' keeps the lower index on ties */
	uint8_t *taken = (uint8_t *)calloc(C ? C : 1, 1);
	for (uint32_t k = 0; k < K; k++) {
		int64_t best = -1;
		for (uint32_t c = 0; c < C; c++) {
			if (taken[c]) continue;
			if (best < 0 || scores[c] > scores[best]) best = c;
		}
		out_idx[k] = (int32_t)best;
		if (best >= 0) taken[best] = 1;
	}
	free(taken);
}

/* ---------------------------------------------------------------- APO:498-625, 635-773 */
static void pat_hit(orc_pattern *p, uint64_t idx)
{
	if (p->count < 3) p->examples[p->count] = (int64_t)idx;   /* slice(0,3), APO:652 ... */
	p->count++;
}

/* Everything _buildReport / _analyzePatterns / getStats accumulate over the records [0,T): tallies, byMode,
 * tool totals (APO:509-538, TCS:602-610), the sequential reward / per-dimension sums (APO:550-568) and the six
 * 'bad'-gated predicates with their first three matches (APO:635-773).  `out` must be zeroed with examples = -1. */
static void report_accumulate(const orc_record *recs, uint64_t T, uint64_t idx_base, const double w[ORC_NDIM], orc_report *out)
{
	out->total += T;
	for (uint64_t t = 0; t < T; t++) {
		const orc_record *r = &recs[t];
		if (r->feedback == 1) out->good++;
		else if (r->feedback == 2) out->bad++;
		else out->none++;
		const int m = r->mode < ORC_NMODE ? r->mode : 0;      /* APO:627-633 */
		out->byMode[m][0]++;
		if (r->feedback == 1) out->byMode[m][1]++;
		if (r->feedback == 2) out->byMode[m][2]++;
		out->toolCalls += r->toolCalls;
		out->toolSucc += r->toolSucc;
		out->toolFail += r->toolFail;
	}
	/* APO:550-568: sequential sums over traces with non-null finalReward */
	for (uint64_t t = 0; t < T; t++) {
		const orc_record *r = &recs[t];
		if (!(r->flags & ORC_F_VALID)) continue;
		double d[ORC_NDIM], fr;
		const uint32_t mask = orc_reward_dims(r, d);
		if (!orc_final_reward(d, mask, w, &fr)) continue;
		out->rewardSum += fr;
		out->withReward++;
		for (int i = 0; i < ORC_NDIM; i++) {
			if (!(mask & (1u << i))) continue;
			out->dim[i].sum += d[i];
			out->dim[i].count++;
		}
	}
	/* APO:635-773.  Every predicate ANDs userFeedback==='bad', so with no bad trace nothing matches — which is
	 * exactly the APO:641 early-out (no bad examples -> no patterns at all). */
	for (uint64_t t = 0; t < T; t++) {
		const orc_record *r = &recs[t];
		if (r->feedback != 2) continue;
		const uint64_t gi = idx_base + t;
		if (r->flags & ORC_F_ERRORS) pat_hit(&out->pat[0], gi);          /* P1 APO:644 */
		if (r->flags & ORC_F_FAILSPAN) pat_hit(&out->pat[1], gi);        /* P2 APO:666-670 */
		if (r->tokens > 10000) pat_hit(&out->pat[2], gi);                /* P3 APO:692-694 */
		if (r->llmCalls > 2) pat_hit(&out->pat[3], gi);                  /* P4 APO:712-714 */
		if (r->userMsgs >= 4) pat_hit(&out->pat[4], gi);                 /* P5 APO:732-735 */
		if ((r->durClass & 0x80) ? (r->durClass & 0x08) != 0 : (double)r->toolDurMs > 15000) pat_hit(&out->pat[5], gi);   /* P6 APO:753-755 */
	}
}

static void report_init(orc_report *out)
{
	memset(out, 0, sizeof(*out));
	for (int p = 0; p < ORC_NPAT; p++) for (int k = 0; k < 3; k++) out->pat[p].examples[k] = -1;
}

/* rates, averages, rule flags and severities from the accumulated totals */
static void report_finish(orc_report *out)
{
	for (int m = 0; m < ORC_NMODE; m++) {                     /* APO:541-544 */
		const uint64_t tot = out->byMode[m][1] + out->byMode[m][2];
		out->byModeGoodRate[m] = tot > 0 ? (double)out->byMode[m][1] / (double)tot : 0;
	}
	{
		const uint64_t tot = out->good + out->bad;              /* APO:546-547 */
		out->goodRate = tot > 0 ? (double)out->good / (double)tot : 0;
	}
	out->toolSuccessRate = out->toolCalls > 0 ? (double)out->toolSucc / (double)out->toolCalls : NAN; /* TCS:624 */
	out->avgReward = out->withReward > 0 ? out->rewardSum / (double)out->withReward : NAN;
	for (int i = 0; i < ORC_NDIM; i++) {
		orc_dimstat *ds = &out->dim[i];
		ds->avg = ds->count > 0 ? ds->sum / (double)ds->count : 0;         /* APO:567 */
		ds->low_flag = (ds->count > 0 && ds->avg < -0.3 && ds->count >= 5);  /* APO:575 */
		ds->low_severity = ds->avg < -0.5 ? 2 : 1;                           /* APO:591 */
		ds->sugg_flag = (ds->count > 0 && ds->avg < 0 && ds->count >= 3);    /* APO:802 */
		ds->sugg_priority = ds->avg < -0.5 ? 2 : 1;                          /* APO:819 */
	}
	if (out->bad == 0) return;                                 /* APO:641 */
	/* thresholds and severities: APO:645,650 / 671,676 / 695,700 / 715,720 / 736,741 / 756,761 */
	static const uint64_t minc[ORC_NPAT] = {2, 2, 3, 2, 2, 2};
	for (int p = 0; p < ORC_NPAT; p++) {
		orc_pattern *pp = &out->pat[p];
		pp->flag = pp->count >= minc[p];
		switch (p) {
		case 0: case 1: pp->severity = pp->count >= 5 ? 2 : 1; break;
		case 2: pp->severity = 1; break;
		case 3: pp->severity = 2; break;
		case 4: pp->severity = pp->count >= 4 ? 2 : 1; break;
		default: pp->severity = 1; break;
		}
	}
}

/* single thread, exact reference order: the parity oracle */
void orc_report_build(const orc_record *recs, uint64_t T, uint64_t idx_base,
                      const double w[ORC_NDIM], orc_report *out)
{
	report_init(out);
	report_accumulate(recs, T, idx_base, w, out);
	report_finish(out);
}

/* merge partial `b` (a later, disjoint index range) into `a`: integers add, binary64 sums add in slice order,
 * the first three examples are taken in index order */
static void report_merge(orc_report *a, const orc_report *b)
{
	a->total += b->total; a->good += b->good; a->bad += b->bad; a->none += b->none;
	for (int m = 0; m < ORC_NMODE; m++) for (int k = 0; k < 3; k++) a->byMode[m][k] += b->byMode[m][k];
	a->toolCalls += b->toolCalls; a->toolSucc += b->toolSucc; a->toolFail += b->toolFail;
	a->withReward += b->withReward; a->rewardSum += b->rewardSum;
	for (int i = 0; i < ORC_NDIM; i++) { a->dim[i].sum += b->dim[i].sum; a->dim[i].count += b->dim[i].count; }
	for (int p = 0; p < ORC_NPAT; p++) {
		uint64_t have = a->pat[p].count < 3 ? a->pat[p].count : 3;
		const uint64_t from = b->pat[p].count < 3 ? b->pat[p].count : 3;
		for (uint64_t k = 0; k < from && have < 3; k++) a->pat[p].examples[have++] = b->pat[p].examples[k];
		a->pat[p].count += b->pat[p].count;
	}
}

typedef struct {
	const orc_record *recs;            /* NULL: generate records [t0, t0+T) of the corpus stream on the fly */
	uint64_t seed, t0; uint32_t agent_permille;
	uint64_t T, idx_base; const double *w;
	orc_report *parts;                 /* one per slice */
	int nslice; uint64_t next;
} rep_job;

static void rep_worker(void *arg, int tid, int nthreads)
{
	(void)tid; (void)nthreads;
	rep_job *j = (rep_job *)arg;
	enum { CHUNK = 4096 };
	orc_record *buf = j->recs ? NULL : (orc_record *)malloc(sizeof(orc_record) * CHUNK);
	for (;;) {
		const uint64_t s = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
		if (s >= (uint64_t)j->nslice) break;
		const uint64_t a = j->T * s / (uint64_t)j->nslice, b = j->T * (s + 1) / (uint64_t)j->nslice;
		orc_report *part = &j->parts[s];
		report_init(part);
		if (j->recs) report_accumulate(j->recs + a, b - a, j->idx_base + a, j->w, part);
		else
			for (uint64_t t = a; t < b; t += CHUNK) {
				const uint64_t n = b - t < CHUNK ? b - t : CHUNK;
				for (uint64_t k = 0; k < n; k++) orc_gen_record(j->seed, ORC_STREAM_CORPUS, 0, j->t0 + t + k, j->agent_permille, &buf[k]);
				report_accumulate(buf, n, j->idx_base + t, j->w, part);
			}
	}
	free(buf);
}

static void report_mt(rep_job *j, orc_report *out, int nthreads)
{
	if (nthreads < 1) nthreads = 1;
	j->nslice = nthreads * 4; j->next = 0;
	j->parts = (orc_report *)malloc(sizeof(orc_report) * (size_t)j->nslice);
	pool_run(rep_worker, j, nthreads);
	report_init(out);
	for (int s = 0; s < j->nslice; s++) report_merge(out, &j->parts[s]);
	report_finish(out);
	free(j->parts);
}

/* all host threads: per-slice partials merged in slice order (integers and examples identical to the
 * single-thread report; the binary64 sums differ from the sequential ones in the last bits) */
void orc_report_build_mt(const orc_record *recs, uint64_t T, uint64_t idx_base,
                         const double w[ORC_NDIM], orc_report *out, int nthreads)
{
	rep_job j = {recs, 0, 0, 0, T, idx_base, w, NULL, 0, 0};
	report_mt(&j, out, nthreads);
}

/* the same over records [t0, t0+T) of the generator's corpus stream, generated chunk by chunk (never
 * materialised): bench.py checks the GPU's corpus report of the full benchmark corpus against it */
void orc_report_generated(uint64_t seed, uint64_t t0, uint64_t T, uint64_t idx_base, uint32_t agent_permille,
                          const double w[ORC_NDIM], orc_report *out, int nthreads)
{
	rep_job j = {NULL, seed, t0, agent_permille, T, idx_base, w, NULL, 0, 0};
	report_mt(&j, out, nthreads);
}

/* ---- exact per-candidate sums over generated Form D evaluations (never materialised) ----
 * For each listed candidate: sum over t in [t0, t0+T) of llrint(finalReward * 2^52), finalReward computed from the
 * fp32-rounded dims exactly as orc_score_dims_fx does on a materialised tensor.  Integer partials per (candidate,
 * slice) item: any thread count gives the same integers. */
typedef struct {
	uint64_t seed; const uint32_t *cands; uint32_t ncand; uint64_t t0, T; uint32_t agent_permille; const double *w;
	__int128 *psum; uint64_t *pcnt; int nslice; uint64_t next;
} genfx_job;

static void genfx_worker(void *arg, int tid, int nthreads)
{
	(void)tid; (void)nthreads;
	genfx_job *j = (genfx_job *)arg;
	const uint64_t nitems = (uint64_t)j->ncand * (uint64_t)j->nslice;
	for (;;) {
		const uint64_t it = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
		if (it >= nitems) break;
		const uint32_t ci = (uint32_t)(it / (uint64_t)j->nslice);
		const uint64_t s = it % (uint64_t)j->nslice;
		const uint64_t a = j->T * s / (uint64_t)j->nslice, b = j->T * (s + 1) / (uint64_t)j->nslice;
		__int128 acc = 0; uint64_t n = 0;
		for (uint64_t t = a; t < b; t++) {
			float row[ORC_NDIM]; double fr;
			orc_gen_dims_row(j->seed, j->cands[ci], j->t0 + t, j->agent_permille, row);
			if (orc_final_reward_f32(row, j->w, &fr)) { acc += (__int128)llrint(fr * 4503599627370496.0); n++; }
		}
		j->psum[it] = acc; j->pcnt[it] = n;
	}
}

void orc_score_generated_fx(uint64_t seed, const uint32_t *cands, uint32_t ncand, uint64_t t0, uint64_t T,
                            uint32_t agent_permille, const double w[ORC_NDIM], uint64_t *lo, int64_t *hi,
                            uint64_t *counts, int nthreads)
{
	if (nthreads < 1) nthreads = 1;
	genfx_job j = {seed, cands, ncand, t0, T, agent_permille, w, NULL, NULL, 0, 0};
	j.nslice = (int)((uint64_t)nthreads * 4 / (ncand ? ncand : 1)) + 1;
	const uint64_t nitems = (uint64_t)ncand * (uint64_t)j.nslice;
	j.psum = (__int128 *)calloc(nitems ? nitems : 1, sizeof(__int128));
	j.pcnt = (uint64_t *)calloc(nitems ? nitems : 1, sizeof(uint64_t));
	pool_run(genfx_worker, &j, nthreads);
	for (uint32_t c = 0; c < ncand; c++) {
		__int128 acc = 0; uint64_t n = 0;
		for (int s = 0; s < j.nslice; s++) { acc += j.psum[(uint64_t)c * j.nslice + s]; n += j.pcnt[(uint64_t)c * j.nslice + s]; }
		lo[c] = (uint64_t)acc; hi[c] = (int64_t)(acc >> 64); counts[c] = n;
	}
	free(j.psum); free(j.pcnt);
}

/* ---------------------------------------------------------------- Form R16 -> Form R */
void orc_unpack16(const orc_record16 *in, uint64_t n, orc_record *out)
{
	for (uint64_t i = 0; i < n; i++) {
		const orc_record16 *p = &in[i];
		orc_record *r = &out[i];
		memset(r, 0, sizeof *r);
		r->feedback = (uint8_t)(p->hdr & 3);
		r->flags = (uint8_t)(((p->hdr >> 2) & 1 ? ORC_F_ERRORS : 0) | ((p->hdr >> 3) & 1 ? ORC_F_ENDED : 0) |
		                     ((p->hdr >> 4) & 1 ? ORC_F_VALID : 0) | ((p->hdr >> 5) & 1 ? ORC_F_FAILSPAN : 0));
		r->mode = (uint8_t)((p->hdr >> 6) & 7);
		r->userMsgs = p->userMsgs; r->asstMsgs = p->asstMsgs;
		r->toolCalls = p->toolCalls; r->toolFail = p->toolFail; r->toolSucc = (uint32_t)p->toolCalls - p->toolFail;
		r->llmCalls = p->llmCalls; r->tokens = p->tokens; r->toolDurMs = p->toolDurMs; r->durClass = p->durClass;
	}
}

/* ================================================================ synthetic generator
 * Build-defined (not from the reference).  Integer-only field derivation so the CUDA
 * generator and this one agree bit for bit.  Spec: DESIGN.md "Generator". */
#define GOLD 0x9E3779B97F4A7C15ull

static inline uint64_t mix64(uint64_t z)
{
	z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
	z ^= z >> 27; z *= 0x94D049BB133111EBull;
	z ^= z >> 31;
	return z;
}

static inline uint32_t ctz64(uint64_t x) { return (uint32_t)__builtin_ctzll(x); }

void orc_gen_record(uint64_t seed, uint32_t stream, uint32_t c, uint64_t t,
                    uint32_t agent_permille, orc_record *out)
{
	const uint64_t base = mix64(mix64(seed ^ ((uint64_t)stream * GOLD)) ^ mix64(((uint64_t)c + 1) * GOLD)) + t * GOLD;
	const uint64_t h1 = mix64(base + GOLD);
	const uint64_t h2 = mix64(base + 2 * GOLD);
	const uint64_t h3 = mix64(base + 3 * GOLD);
	const uint64_t h4 = mix64(base + 4 * GOLD);
	const uint32_t qc = stream == ORC_STREAM_CORPUS ? 512u
	                  : (uint32_t)(mix64(seed ^ 0xC0FFEEull ^ ((uint64_t)c * GOLD)) & 1023);

	const uint32_t r = (uint32_t)(h1 & 1023);
	const uint32_t pgood = 192 + (qc >> 2);
	const uint32_t pbad = 256 - (qc >> 3);
	const uint8_t feedback = r < pgood ? 1 : (r < pgood + pbad ? 2 : 0);
	const int hasErrors = ((h1 >> 10) & 1023) < 102;
	const int ended = ((h1 >> 20) & 1023) < 973;
	uint8_t mode;
	if (((h1 >> 30) & 1023) < agent_permille) mode = 2;
	else {
		const uint32_t k = (uint32_t)((h1 >> 40) & 15);
		mode = k < 10 ? 1 : (k < 12 ? 3 : (k < 14 ? 4 : 0));
	}
	const int agent = mode == 2;

	uint32_t toolCalls = 0, toolFail = 0;
	float dur = 0.0f;
	if (!(((h1 >> 44) & 1023) < 307)) {
		const uint32_t g = ctz64(h2 | (1ull << 20));
		toolCalls = agent ? 1 + (uint32_t)((h2 >> 21) & 15) + g : 1 + g;
		const uint32_t n = toolCalls < 16 ? toolCalls : 16;
		for (uint32_t i = 0; i < n; i++) if (((h3 >> (4 * i)) & 15) == 0) toolFail++;
		if (((h2 >> 25) & 15) != 0) {
			const uint32_t e = (uint32_t)((h2 >> 29) & 15) % 10;
			const uint32_t basems = 50u << e;
			const uint32_t frac = (uint32_t)((h2 >> 33) & 4095);
			const uint32_t avg = basems + ((basems * frac) >> 12);
			dur = (float)(avg * toolCalls);
		}
	}
	uint32_t llm = 0;
	if (((h2 >> 45) & 31) != 0)
		llm = 1 + ctz64((h2 >> 50) | (1ull << 13)) + (agent ? (uint32_t)((h2 >> 48) & 3) : 0);
	uint32_t tokens = 0;
	if (!((h4 & 1023) < 205)) {
		const uint32_t e = (uint32_t)((h4 >> 10) & 15) % 9;
		const uint32_t b = 200u << e;
		const uint32_t f = (uint32_t)((h4 >> 14) & 4095);
		tokens = b + ((b * f) >> 12);
	}
	uint32_t userMsgs = 0;
	if (((h4 >> 40) & 63) != 0) userMsgs = 1 + ctz64((h4 >> 26) | (1ull << 12));

	out->feedback = feedback;
	out->flags = (uint8_t)((hasErrors ? ORC_F_ERRORS : 0) | (ended ? ORC_F_ENDED : 0) |
	                       ((ended || feedback) ? ORC_F_VALID : 0) | (toolFail > 0 ? ORC_F_FAILSPAN : 0));
	out->mode = mode;
	{   /* the generator's durations are integers (dur = avg * toolCalls < 2^24): the class follows from them, as on the device */
		const uint32_t dms = (uint32_t)dur, avg = toolCalls ? dms / toolCalls : 0u;
		out->durClass = (uint8_t)(0x80u | (dms > 0u ? 0x04u : 0u) | (dms > 15000u ? 0x08u : 0u) | ((avg > 1000u) + (avg > 3000u) + (avg > 10000u)));
	}
	out->userMsgs = (uint16_t)userMsgs;
	const uint32_t asst = llm + ((((h4 >> 46) & 15) == 0) ? 1u : 0u);   /* 1/16: an assistant span without an llm_call record */
	out->asstMsgs = (uint16_t)(asst < 65535 ? asst : 65535);
	out->toolCalls = toolCalls;
	out->toolSucc = toolCalls - toolFail;
	out->toolFail = toolFail;
	out->llmCalls = llm;
	out->tokens = tokens;
	out->toolDurMs = dur;
}

void orc_gen_dims_row(uint64_t seed, uint32_t c, uint64_t t, uint32_t agent_permille, float out[ORC_NDIM])
{
	orc_record r;
	orc_gen_record(seed, ORC_STREAM_ROLLOUT, c, t, agent_permille, &r);
	double d[ORC_NDIM];
	const uint32_t mask = orc_reward_dims(&r, d);
	for (int i = 0; i < ORC_NDIM; i++)
		out[i] = ((r.flags & ORC_F_VALID) && (mask & (1u << i))) ? (float)d[i] : NAN;
}

typedef struct {
	uint64_t seed; uint32_t stream, c0, C; uint64_t t0, T, pitch; uint32_t ap;
	float *dims; orc_record *recs;
} gen_job;

/* thread k fills the k-th T-slice of every candidate: with pinned workers the pages it first touches are the
 * ones the same thread reads back in score_mt's slice k */
static void gen_worker(void *arg, int tid, int nthreads)
{
	gen_job *j = (gen_job *)arg;
	for (uint32_t c = 0; c < j->C; c++) {
		const uint64_t a = j->T * (uint64_t)tid / (uint64_t)nthreads;
		const uint64_t b = j->T * (uint64_t)(tid + 1) / (uint64_t)nthreads;
		for (uint64_t t = a; t < b; t++) {
			if (j->dims)
				orc_gen_dims_row(j->seed, j->c0 + c, j->t0 + t, j->ap, j->dims + ((uint64_t)c * j->pitch + t) * ORC_NDIM);
			else
				orc_gen_record(j->seed, j->stream, j->c0 + c, j->t0 + t, j->ap, j->recs + (uint64_t)c * j->pitch + t);
		}
	}
}

static void gen_mt(gen_job proto, int nthreads)
{
	if (nthreads < 1) nthreads = 1;
	pool_run(gen_worker, &proto, nthreads);
}

void orc_gen_dims(uint64_t seed, uint32_t c0, uint32_t C, uint64_t t0, uint64_t T, uint64_t pitch_evals,
                  uint32_t agent_permille, float *out, int nthreads)
{
	gen_job p = {seed, ORC_STREAM_ROLLOUT, c0, C, t0, T, pitch_evals, agent_permille, out, NULL};
	gen_mt(p, nthreads);
}

void orc_gen_records(uint64_t seed, uint32_t stream, uint32_t c0, uint32_t C, uint64_t t0, uint64_t T,
                     uint64_t pitch, uint32_t agent_permille, orc_record *out, int nthreads)
{
	gen_job p = {seed, stream, c0, C, t0, T, pitch, agent_permille, NULL, out};
	gen_mt(p, nthreads);
}
