/* GPU test of napi/apo_jobs.c through the C ABI: the re-entrancy rule of include/apo_b200.h ("one in-flight call per
 * handle") holds for overlapping JS calls because jobs of a handle run in submission order.
 *
 * Two "JS callers" interleave on ONE handle: A uploads tensor X and scores it resident, B uploads tensor Y and scores it;
 * the submission order is A.upload, A.score, B.upload, A.score2, B.score, reward batch, host score ...  executed by 6
 * worker threads.  Every result must equal the same sequence run on a fresh handle one call at a time. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "apo_jobs.h"

#define NC 6
#define NT 20000
#define NW 6
static apo_serial *S;
static apo_job jobs[64];
static int njobs, next_job;

static void *worker(void *p)
{
	(void)p;
	for (;;) {
		const int i = __atomic_fetch_add(&next_job, 1, __ATOMIC_SEQ_CST);
		if (i >= njobs) break;
		apo_job_run(S, &jobs[i]);
	}
	return NULL;
}

static float *make_dims(unsigned seed)
{
	float *d = (float *)malloc(sizeof(float) * NC * NT * 9);
	for (size_t i = 0; i < (size_t)NC * NT * 9; i++) {
		seed = seed * 1664525u + 1013904223u;
		const unsigned r = seed >> 24;
		d[i] = (r & 7) == 0 ? NAN : (float)((int)(r % 5) - 2) * 0.5f;
	}
	return d;
}

static apo_job *push(apo_job_kind k)
{
	apo_job *j = &jobs[njobs++];
	memset(j, 0, sizeof *j);
	j->kind = k;
	return j;
}

static void seal(apo_job *j)
{
	if (apo_job_validate(j) == APO_OK) apo_job_prepare(j);
	j->ticket = apo_serial_ticket(S);
}

static void build_sequence(float *X, float *Y, apo_record *recs)
{
	njobs = 0;
	apo_job *j;
	j = push(APO_JOB_DIMS_UPLOAD); j->buf = X; j->buf_bytes = sizeof(float) * NC * NT * 9; j->C = NC; j->T = NT; seal(j);          /* 0 A */
	j = push(APO_JOB_SCORE_RESIDENT); j->C = NC; j->K = 3; seal(j);                                                             /* 1 A: scores X */
	j = push(APO_JOB_DIMS_UPLOAD); j->buf = Y; j->buf_bytes = sizeof(float) * NC * NT * 9; j->C = NC; j->T = NT; j->compact = 0; seal(j); /* 2 B */
	j = push(APO_JOB_SCORE_RESIDENT); j->C = NC; j->K = 3; seal(j);                                                             /* 3 A again: must see Y */
	j = push(APO_JOB_SCORE_RESIDENT); j->C = NC; j->K = NC; j->first = 4000; j->count = 8000; seal(j);                          /* 4 B: window of Y */
	j = push(APO_JOB_REWARD_BATCH); j->buf = recs; j->buf_bytes = sizeof(apo_record) * 64; seal(j);                           /* 5 */
	j = push(APO_JOB_SCORE_HOST); j->buf = X; j->buf_bytes = sizeof(float) * NC * NT * 9; j->C = NC; j->T = NT; j->K = 2;
	j->corpus = recs; j->corpus_bytes = sizeof(apo_record) * 64; seal(j);                                                      /* 6 host streaming + corpus */
	j = push(APO_JOB_SCORE_HOST); j->buf = X; j->buf_bytes = 100; j->C = NC; j->T = NT; j->K = 2; seal(j);                       /* 7 invalid: rejected in its turn */
	j = push(APO_JOB_SCORE_RESIDENT); j->C = NC; j->K = 1; seal(j);                                                             /* 8 still Y resident */
}

int main(void)
{
	char err[256];
	float *X = make_dims(1), *Y = make_dims(2);
	apo_record recs[64];
	memset(recs, 0, sizeof recs);
	for (int i = 0; i < 64; i++) { recs[i].feedback = (uint8_t)(i % 3); recs[i].flags = APO_F_VALID | APO_F_ENDED | (i % 5 == 0 ? APO_F_ERRORS : 0); recs[i].mode = 1 + i % 2;
		recs[i].toolCalls = i % 7; recs[i].toolFail = i % 7 ? i % 2 : 0; recs[i].toolSucc = recs[i].toolCalls - recs[i].toolFail; recs[i].llmCalls = 1 + i % 4;
		recs[i].tokens = 1000u * (i % 13); recs[i].userMsgs = 1 + i % 5; recs[i].asstMsgs = 1 + i % 4; recs[i].toolDurMs = 900.0f * (i % 9); }

	/* sequential truth on its own handle */
	S = apo_serial_create(0, err, sizeof err);
	if (!S) { printf("no engine: %s\n", err); return 2; }
	build_sequence(X, Y, recs);
	for (int i = 0; i < njobs; i++) apo_job_run(S, &jobs[i]);
	static apo_job truth[64];
	memcpy(truth, jobs, sizeof(apo_job) * njobs);
	const int n = njobs;
	apo_serial_destroy(S);

	int fails = 0;
	for (int round = 0; round < 5; round++) {
		S = apo_serial_create(0, err, sizeof err);
		build_sequence(X, Y, recs);
		next_job = 0;
		pthread_t th[NW];
		for (int i = 0; i < NW; i++) pthread_create(&th[i], NULL, worker, NULL);
		for (int i = 0; i < NW; i++) pthread_join(th[i], NULL);
		for (int i = 0; i < n; i++) {
			const apo_job *a = &truth[i], *b = &jobs[i];
			if (a->rc != b->rc) { printf("FAIL round %d job %d: rc %d vs %d (%s)\n", round, i, b->rc, a->rc, b->err); fails++; continue; }
			if (a->rc != APO_OK) continue;
			if (a->scores && (memcmp(a->scores, b->scores, 8 * NC) || memcmp(a->counts, b->counts, 8 * NC) || memcmp(a->topk, b->topk, 4 * (a->K < NC ? a->K : NC)))) {
				printf("FAIL round %d job %d: scores / counts / top-K differ from the sequential run\n", round, i); fails++;
			}
			if (a->kind == APO_JOB_SCORE_HOST && memcmp(&a->report, &b->report, sizeof a->report)) { printf("FAIL round %d job %d: report differs\n", round, i); fails++; }
			if (a->finals_out && (memcmp(a->finals_out, b->finals_out, 8 * 64) || memcmp(a->masks_out, b->masks_out, 4 * 64))) { printf("FAIL round %d job %d: rewards differ\n", round, i); fails++; }
		}
		if (round == 0) {
			if (truth[7].rc != APO_E_ARG) { printf("FAIL: undersized buffer accepted\n"); fails++; }
			if (!memcmp(truth[1].scores, truth[3].scores, 8 * NC)) { printf("FAIL: job 3 did not see tensor Y\n"); fails++; }
			if (memcmp(truth[3].scores, truth[8].scores, 8 * NC)) { printf("FAIL: resident tensor changed under the host-streaming call\n"); fails++; }
		}
		for (int i = 0; i < n; i++) apo_job_release(&jobs[i]);
		apo_serial_destroy(S);
	}
	for (int i = 0; i < n; i++) apo_job_release(&truth[i]);
	free(X); free(Y);
	printf(fails ? "FAILED (%d)\n" : "OK: %d overlapping jobs x 5 rounds on %d threads equal the sequential run\n", fails ? fails : n, NW);
	return fails ? 1 : 0;
}
