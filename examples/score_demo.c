/* score_demo.c — the C ABI used from plain C, the way the N-API addon (napi/apo_napi.c) or any
 * other FFI would: no C++, no CUDA, no torch in sight.
 *
 *   gcc -std=c11 -I include examples/score_demo.c -L senweaver-ide_b200 -lapo_b200 \
 *       -Wl,-rpath,$PWD/senweaver-ide_b200 -lm -o /tmp/score_demo && /tmp/score_demo
 *
 * Scores 8 candidates x 20000 records (device generator), prints the top-4 beam and the six
 * pattern counts, then shows the error convention (negative code + apo_last_error, never abort). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "apo_b200.h"

int main(void)
{
	apo_engine *e = NULL;
	if (apo_create(0, &e) != APO_OK) {
		printf("apo_create failed (expected on a box without a B200): %s\n", apo_last_error(NULL));
		return 2;
	}
	enum { C = 8, K = 4 };
	const uint64_t T = 20000, seed = 0x5EED0001;
	int rc = apo_dims_generate(e, seed, 0, C, 0, T, 300);
	if (rc == APO_OK) rc = apo_corpus_generate(e, seed, 0, T, 300);
	if (rc == APO_OK) rc = apo_dims_compact(e);                 /* optional: 14 B/eval resident layout */
	if (rc != APO_OK && rc != APO_E_STATE) { printf("setup failed: %s\n", apo_last_error(e)); return 1; }

	apo_score_opts o; memset(&o, 0, sizeof o);
	o.K = K; o.source = APO_SRC_DIMS; o.flags = APO_SCORE_CORPUS;
	double scores[C]; uint64_t counts[C]; int32_t topk[K]; apo_corpus_report rep;
	if (apo_score(e, &o, scores, counts, topk, &rep) != APO_OK) { printf("apo_score failed: %s\n", apo_last_error(e)); return 1; }
	printf("layout=%d  top-%d:", apo_dims_layout(e), K);
	for (int k = 0; k < K; k++) printf(" c%d(%.6f)", topk[k], scores[topk[k]]);
	printf("\ncorpus: total=%llu good=%llu bad=%llu avgReward=%.6f  patterns:", (unsigned long long)rep.total,
	       (unsigned long long)rep.good, (unsigned long long)rep.bad, rep.avgReward);
	for (int p = 0; p < APO_NPAT; p++) printf(" %llu%s", (unsigned long long)rep.pat[p].count, rep.pat[p].flag ? "*" : "");
	printf("\n");
	for (int k = 1; k < K; k++) if (!(scores[topk[k - 1]] >= scores[topk[k]])) { printf("top-K not sorted\n"); return 1; }

	o.K = C + 1;                                                 /* error convention: code + message, state untouched */
	rc = apo_score(e, &o, scores, counts, topk, NULL);
	printf("K > C -> rc=%d (%s)\n", rc, apo_last_error(e));
	apo_destroy(e);
	return rc == APO_E_ARG ? 0 : 1;
}
