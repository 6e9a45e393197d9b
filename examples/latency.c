/* Wall-clock latency of the small calls the IDE really makes (<= 1000 traces, TCS:219), measured in C against the
 * C ABI (no Python / ctypes overhead):  gcc -O2 -I include examples/latency.c -L senweaver-ide_b200 -lapo_b200 -lm
 *   apo_reward_one            one trace  -> dims, mask, finalReward       (endTrace / recordUserFeedback, TCS:413,547)
 *   apo_score report-only     1000-trace corpus report                    (_buildReport, APO:498-625)
 *   apo_score 4 x 1000 + K=2  candidates x traces + report + top-K        (configs[0])
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "apo_b200.h"

static double now_us(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }

int main(void)
{
	apo_engine *e = NULL;
	if (apo_create(0, &e) != APO_OK) { printf("apo_create failed: %s\n", apo_last_error(NULL)); return 2; }
	enum { C = 4, T = 1000, REPS = 2000 };
	if (apo_dims_generate(e, 0x5EED0001, 0, C, 0, T, 300) || apo_corpus_generate(e, 0x5EED0001, 0, T, 300)) { printf("%s\n", apo_last_error(e)); return 1; }
	apo_record rec[1];
	apo_corpus_download(e, rec, 0, 1);
	double dims[APO_NDIM], fin, scores[C]; uint32_t mask; uint64_t counts[C]; int32_t topk[2]; apo_corpus_report rep;
	apo_score_opts o; memset(&o, 0, sizeof o);
	double t0, a, b, c;
	for (int i = 0; i < 50; i++) apo_reward_one(e, rec, dims, &mask, &fin);
	t0 = now_us(); for (int i = 0; i < REPS; i++) apo_reward_one(e, rec, dims, &mask, &fin); a = (now_us() - t0) / REPS;
	o.K = 0; o.flags = APO_SCORE_CORPUS; o.count = 4;
	for (int i = 0; i < 50; i++) apo_score(e, &o, scores, counts, NULL, &rep);
	t0 = now_us(); for (int i = 0; i < REPS; i++) apo_score(e, &o, scores, counts, NULL, &rep); b = (now_us() - t0) / REPS;
	o.K = 2; o.count = 0;
	for (int i = 0; i < 50; i++) apo_score(e, &o, scores, counts, topk, &rep);
	t0 = now_us(); for (int i = 0; i < REPS; i++) apo_score(e, &o, scores, counts, topk, &rep); c = (now_us() - t0) / REPS;
	apo_timing tm; apo_last_timing(e, &tm);
	printf("{\"reward_one_us\": %.2f, \"report_1000_traces_us\": %.2f, \"score_4x1000_top2_with_report_us\": %.2f, \"launches_per_score\": %u, \"tail_finalize_us\": %.2f, "
	       "\"tail_publish_us\": %.2f, \"reps\": %d, \"final\": %.17g, \"top\": [%d, %d]}\n",
	       a, b, c, tm.launches, tm.tail_finalize_ms * 1e3, tm.tail_publish_ms * 1e3, REPS, fin, topk[0], topk[1]);
	apo_destroy(e);
	return 0;
}
