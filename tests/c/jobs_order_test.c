/* CPU-only test of napi/apo_jobs.c: the per-handle FIFO (strict ticket order, never two jobs of one handle at once,
 * with more worker threads than a libuv pool has) and the argument validation every JS-reachable call goes through.
 * Built and run by tests/test_napi_jobs.py; needs no GPU (a detached serial has no engine). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "apo_jobs.h"

#define NJOBS 400
#define NWORK 8
static apo_serial *S;
static apo_job jobs[NJOBS];
static int order[NJOBS], norder, busy, overlap;
static int next_job;

static void hook(void *arg)
{
	if (__atomic_fetch_add(&busy, 1, __ATOMIC_SEQ_CST) != 0) overlap = 1;
	const int id = (int)(intptr_t)arg;
	if (id % 7 == 0) { struct timespec ts = {0, 200000}; nanosleep(&ts, NULL); }
	order[norder++] = id;                       /* no lock: the FIFO is what makes this safe */
	__atomic_fetch_sub(&busy, 1, __ATOMIC_SEQ_CST);
}

static void *worker(void *p)
{
	(void)p;
	for (;;) {                                   /* like libuv: dequeue in submission order, run on whichever thread is free */
		const int i = __atomic_fetch_add(&next_job, 1, __ATOMIC_SEQ_CST);
		if (i >= NJOBS) break;
		if (i % 5 == 0) { struct timespec ts = {0, 50000}; nanosleep(&ts, NULL); }   /* let later tickets reach the lock first */
		apo_job_run(S, &jobs[i]);
	}
	return NULL;
}

static int expect_bad(apo_job j, const char *what)
{
	const int rc = apo_job_validate(&j);
	if (rc != APO_E_ARG || !j.err[0]) { printf("FAIL: %s was accepted (rc=%d)\n", what, rc); return 1; }
	return 0;
}

int main(void)
{
	int fails = 0;
	S = apo_serial_create_detached();
	for (int i = 0; i < NJOBS; i++) {
		memset(&jobs[i], 0, sizeof jobs[i]);
		jobs[i].kind = APO_JOB_TEST_HOOK; jobs[i].hook = hook; jobs[i].hook_arg = (void *)(intptr_t)i;
		if (i == 13) { jobs[i].kind = APO_JOB_SCORE_HOST; }          /* invalid job in the middle: takes its turn, fails, FIFO moves on */
		apo_job_validate(&jobs[i]);
		jobs[i].ticket = apo_serial_ticket(S);
	}
	pthread_t th[NWORK];
	for (int i = 0; i < NWORK; i++) pthread_create(&th[i], NULL, worker, NULL);
	for (int i = 0; i < NWORK; i++) pthread_join(th[i], NULL);
	if (overlap) { printf("FAIL: two jobs of one handle ran at the same time\n"); fails++; }
	if (norder != NJOBS - 1) { printf("FAIL: %d hooks ran, expected %d\n", norder, NJOBS - 1); fails++; }
	for (int i = 0, want = 0; i < norder; i++, want++) {
		if (want == 13) want++;
		if (order[i] != want) { printf("FAIL: position %d ran job %d, expected %d\n", i, order[i], want); fails++; break; }
	}
	if (jobs[13].rc != APO_E_ARG) { printf("FAIL: invalid job rc=%d\n", jobs[13].rc); fails++; }
	/* a compute job on a handle without an engine fails loudly, it does not fall back */
	{
		static float dims[4 * 9];
		apo_job j; memset(&j, 0, sizeof j);
		j.kind = APO_JOB_DIMS_UPLOAD; j.buf = dims; j.buf_bytes = sizeof dims; j.C = 1; j.T = 4;
		if (apo_job_validate(&j) != APO_OK) { printf("FAIL: valid dims upload rejected: %s\n", j.err); fails++; }
		j.ticket = apo_serial_ticket(S);
		apo_job_run(S, &j);
		if (j.rc != APO_E_STATE) { printf("FAIL: engine-less compute job rc=%d\n", j.rc); fails++; }
	}
	apo_serial_destroy(S);

	/* ---- validation: everything the renderer can send over IPC */
	static float small[36];
	apo_job j; memset(&j, 0, sizeof j);
	j.kind = APO_JOB_SCORE_HOST; j.buf = small; j.buf_bytes = sizeof small; j.C = 1; j.T = 4; j.K = 1;
	if (apo_job_validate(&j) != APO_OK) { printf("FAIL: 1 x 4 Form D rejected: %s\n", j.err); fails++; }
	{ apo_job b = j; b.T = 5; fails += expect_bad(b, "dims buffer shorter than C*T*36"); }
	{ apo_job b = j; b.C = 0; fails += expect_bad(b, "C == 0"); }
	{ apo_job b = j; b.K = 2; fails += expect_bad(b, "K > C"); }
	{ apo_job b = j; b.C = 0x80000000u; b.T = 0x4000000000000ull; fails += expect_bad(b, "C*T*36 overflowing 64 bits"); }
	{ apo_job b = j; b.buf = NULL; fails += expect_bad(b, "missing buffer"); }
	{ apo_job b = j; b.corpus = small; b.corpus_bytes = 33; fails += expect_bad(b, "corpus length not a multiple of 32"); }
	{ apo_job b = j; b.kind = APO_JOB_SCORE_HOST_RECORDS; b.row_bytes = 24; fails += expect_bad(b, "rowBytes 24"); }
	{ apo_job b = j; b.kind = APO_JOB_SCORE_HOST_RECORDS; b.row_bytes = 32; b.T = 5; fails += expect_bad(b, "records buffer shorter than C*T*32"); }
	{ apo_job b = j; b.kind = APO_JOB_REWARD_BATCH; b.buf_bytes = 40; fails += expect_bad(b, "record bytes not a multiple of 32"); }
	{ apo_job b = j; b.kind = APO_JOB_COMM_INIT; b.nranks = 2; b.rank = 2; fails += expect_bad(b, "rank >= nranks"); }
	{
		/* Form T: 1 x 4 evaluations, two dictionary entries */
		static uint16_t tl[4]; static uint8_t th[4]; static uint32_t tpc[2], book[8 * 256], d2b[4096]; static uint16_t tpd[2];
		apo_job t; memset(&t, 0, sizeof t);
		t.kind = APO_JOB_SCORE_HOST_TUPLES; t.C = 1; t.T = 4; t.K = 1; t.n_tuples = 2;
		t.buf = tl; t.buf_bytes = sizeof tl; t.aux[0] = th; t.aux_bytes[0] = sizeof th; t.aux[1] = tpc; t.aux_bytes[1] = sizeof tpc;
		t.aux[2] = tpd; t.aux_bytes[2] = sizeof tpd; t.aux[3] = book; t.aux_bytes[3] = sizeof book; t.aux[4] = d2b; t.aux_bytes[4] = sizeof d2b;
		if (apo_job_validate(&t) != APO_OK) { printf("FAIL: valid Form T job rejected: %s\n", t.err); fails++; }
		{ apo_job b = t; b.T = 5; fails += expect_bad(b, "tl plane shorter than C*T*2"); }
		{ apo_job b = t; b.aux_bytes[0] = 3; fails += expect_bad(b, "th plane shorter than C*T"); }
		{ apo_job b = t; b.n_tuples = 3; fails += expect_bad(b, "dictionary shorter than nTuples"); }
		{ apo_job b = t; b.n_tuples = 0; fails += expect_bad(b, "evaluations without a dictionary"); }
		{ apo_job b = t; b.n_tuples = 1u << 24; fails += expect_bad(b, "too many dictionary entries"); }
		{ apo_job b = t; b.aux_bytes[3] = 100; fails += expect_bad(b, "codebook of the wrong size"); }
		{ apo_job b = t; b.aux[4] = NULL; fails += expect_bad(b, "d2book missing"); }
		{ apo_job b = t; b.K = 2; fails += expect_bad(b, "K exceeds C"); }
	}
	{ apo_job b = j; b.kind = APO_JOB_SCORE_RESIDENT; b.first = 3; fails += expect_bad(b, "window start not a multiple of 4"); }
	{ apo_job b = j; b.kind = APO_JOB_SCORE_RESIDENT; b.first = 0; b.K = 20000; b.C = 30000; fails += expect_bad(b, "K above the beam-width limit"); }
	{ apo_job b = j; b.kind = (apo_job_kind)99; fails += expect_bad(b, "unknown kind"); }
	/* create never throws / aborts: without a device it returns NULL and a message */
	{
		char err[256] = {0};
		apo_serial *s = apo_serial_create(0, err, sizeof err);
		if (s) apo_serial_destroy(s);
		else if (!err[0]) { printf("FAIL: create failed without a message\n"); fails++; }
		else printf("create without a usable device: %s\n", err);
	}
	printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
	return fails ? 1 : 0;
}
