/*
 * apo_napi.c — thin N-API addon over the C ABI of include/apo_b200.h.
 *
 * Loaded by the IDE's *main process* (the renderer cannot load native modules:
 * eslint.config.js:94 — `common` imports nothing platform specific), behind a named IPC
 * channel exactly like the reference's metrics service
 * (common/metricsService.ts:25-50 + electron-main/metricsMainService.ts:35, registered in
 * src/vs/code/electron-main/app.ts:1124,1259-1260).  See INTEGRATION.md.
 *
 * N-API is a stable C ABI (Electron 34 -> Node 20.18 -> N-API v9).  node_api.h is not
 * present in this build image, so the handful of prototypes used are declared below when the
 * real header is absent; `gcc -fsyntax-only -I../include apo_napi.c` is the check run here,
 * the real build is `node-gyp` / `cmake-js` with the genuine header.
 *
 * Heavy calls (score) run on the libuv pool via napi_create_async_work and resolve a
 * Promise: "anything transmitted over a channel must be async" (metricsService.ts:47), and
 * the Electron main loop must never block.  Typed arrays cross the IPC channel as raw bytes
 * (base/parts/ipc/common/ipc.ts:274-283).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "apo_b200.h"

#if defined(__has_include)
#if __has_include(<node_api.h>)
#include <node_api.h>
#define APO_HAVE_NODE_API 1
#endif
#endif
#ifndef APO_HAVE_NODE_API
/* ---- minimal N-API v9 declarations (subset used here) ---- */
typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_callback_info__ *napi_callback_info;
typedef struct napi_deferred__ *napi_deferred;
typedef struct napi_async_work__ *napi_async_work;
typedef struct napi_ref__ *napi_ref;
typedef enum { napi_ok = 0 } napi_status;
typedef napi_value (*napi_callback)(napi_env, napi_callback_info);
typedef void (*napi_finalize)(napi_env, void *, void *);
typedef void (*napi_async_execute_callback)(napi_env, void *);
typedef void (*napi_async_complete_callback)(napi_env, napi_status, void *);
typedef enum { napi_default = 0 } napi_property_attributes;
typedef struct { const char *utf8name; napi_value name; napi_callback method, getter, setter; napi_value value;
                 napi_property_attributes attributes; void *data; } napi_property_descriptor;
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t *, napi_value *, napi_value *, void **);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t *);
napi_status napi_get_value_uint32(napi_env, napi_value, uint32_t *);
napi_status napi_get_value_double(napi_env, napi_value, double *);
napi_status napi_get_arraybuffer_info(napi_env, napi_value, void **, size_t *);
napi_status napi_create_arraybuffer(napi_env, size_t, void **, napi_value *);
napi_status napi_create_external_arraybuffer(napi_env, void *, size_t, napi_finalize, void *, napi_value *);
napi_status napi_create_object(napi_env, napi_value *);
napi_status napi_create_double(napi_env, double, napi_value *);
napi_status napi_create_int32(napi_env, int32_t, napi_value *);
napi_status napi_create_string_utf8(napi_env, const char *, size_t, napi_value *);
napi_status napi_set_named_property(napi_env, napi_value, const char *, napi_value);
napi_status napi_create_external(napi_env, void *, napi_finalize, void *, napi_value *);
napi_status napi_get_value_external(napi_env, napi_value, void **);
napi_status napi_create_promise(napi_env, napi_deferred *, napi_value *);
napi_status napi_resolve_deferred(napi_env, napi_deferred, napi_value);
napi_status napi_reject_deferred(napi_env, napi_deferred, napi_value);
napi_status napi_create_async_work(napi_env, napi_value, napi_value, napi_async_execute_callback,
                                   napi_async_complete_callback, void *, napi_async_work *);
napi_status napi_queue_async_work(napi_env, napi_async_work);
napi_status napi_delete_async_work(napi_env, napi_async_work);
napi_status napi_create_reference(napi_env, napi_value, uint32_t, napi_ref *);
napi_status napi_delete_reference(napi_env, napi_ref);
napi_status napi_define_properties(napi_env, napi_value, size_t, const napi_property_descriptor *);
napi_status napi_create_error(napi_env, napi_value, napi_value, napi_value *);
napi_status napi_throw_error(napi_env, const char *, const char *);
#define NAPI_AUTO_LENGTH ((size_t)-1)
#define NAPI_MODULE_INIT() napi_value napi_register_module_v1(napi_env env, napi_value exports)
#endif

static void engine_finalize(napi_env env, void *data, void *hint) { (void)env; (void)hint; apo_destroy((apo_engine *)data); }

static apo_engine *get_engine(napi_env env, napi_value v) { void *p = NULL; napi_get_value_external(env, v, &p); return (apo_engine *)p; }

/* create(device:number) -> external handle */
static napi_value Create(napi_env env, napi_callback_info info) {
	size_t argc = 1; napi_value argv[1]; int32_t dev = 0;
	napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
	if (argc > 0) napi_get_value_int32(env, argv[0], &dev);
	apo_engine *e = NULL;
	if (apo_create(dev, &e) != APO_OK) { napi_throw_error(env, "APO_E_CUDA", apo_last_error(NULL)); return NULL; }
	napi_value out; napi_create_external(env, e, engine_finalize, NULL, &out);
	return out;
}

/* rewardBatch(handle, records:ArrayBuffer) -> {dims:ArrayBuffer(f64 n*9), masks:ArrayBuffer(u32 n), finals:ArrayBuffer(f64 n)}
 * = TraceCollectorService._computeRewardSignals for n traces (TCS:668-788). Small and latency bound: synchronous. */
static napi_value RewardBatch(napi_env env, napi_callback_info info) {
	size_t argc = 2; napi_value argv[2];
	napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
	apo_engine *e = get_engine(env, argv[0]);
	void *recs; size_t bytes; napi_get_arraybuffer_info(env, argv[1], &recs, &bytes);
	const uint64_t n = bytes / sizeof(apo_record);
	void *dims, *masks, *finals; napi_value vd, vm, vf, out;
	napi_create_arraybuffer(env, n * APO_NDIM * 8, &dims, &vd);
	napi_create_arraybuffer(env, n * 4, &masks, &vm);
	napi_create_arraybuffer(env, n * 8, &finals, &vf);
	if (apo_reward_batch(e, (const apo_record *)recs, n, (double *)dims, (uint32_t *)masks, (double *)finals) != APO_OK) {
		napi_throw_error(env, "APO", apo_last_error(e)); return NULL;
	}
	napi_create_object(env, &out);
	napi_set_named_property(env, out, "dims", vd); napi_set_named_property(env, out, "masks", vm); napi_set_named_property(env, out, "finals", vf);
	return out;
}

/* allocPinned(bytes:number) -> ArrayBuffer over page-locked memory (apo_host_alloc).  Typed arrays built on it
 * (the dims / record buffers handed to score) are read in place by the streaming calls at PCIe line rate; ordinary
 * ArrayBuffers work too, through the library's staging path, at about half that. */
static void pinned_finalize(napi_env env, void *data, void *hint) { (void)env; (void)hint; apo_host_free(data); }

static napi_value AllocPinned(napi_env env, napi_callback_info info) {
	size_t argc = 1; napi_value argv[1], out; double bytes = 0; void *p = NULL;
	napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
	napi_get_value_double(env, argv[0], &bytes);
	if (bytes < 0 || apo_host_alloc((uint64_t)bytes, &p) != APO_OK) { napi_throw_error(env, "APO_E_NOMEM", apo_last_error(NULL)); return NULL; }
	napi_create_external_arraybuffer(env, p, (size_t)bytes, pinned_finalize, NULL, &out);
	return out;
}

/* recordsFromJson(utf8:ArrayBuffer) -> ArrayBuffer of apo_record[T]
 * utf8 = the string stored under 'senweaver.traceCollector.data' (TCS:296-359), encoded with TextEncoder.
 * Host-side format code (apo_records_from_json); linear in the input, no GPU work. */
static napi_value RecordsFromJson(napi_env env, napi_callback_info info) {
	size_t argc = 1; napi_value argv[1], out;
	napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
	void *text; size_t bytes; uint64_t pos = 0;
	napi_get_arraybuffer_info(env, argv[0], &text, &bytes);
	int64_t n = apo_records_from_json((const char *)text, bytes, NULL, 0, &pos);
	if (n < 0) { napi_throw_error(env, "APO_E_ARG", "malformed trace JSON"); return NULL; }
	void *recs;
	napi_create_arraybuffer(env, (size_t)n * sizeof(apo_record), &recs, &out);
	apo_records_from_json((const char *)text, bytes, (apo_record *)recs, (uint64_t)n, &pos);
	return out;
}

/* ---- score(handle, {dims:ArrayBuffer, C, T, corpus:ArrayBuffer|null, K}) -> Promise<{scores,counts,topk,report}> ---- */
typedef struct {
	apo_engine *e; napi_deferred deferred; napi_async_work work; napi_ref keep_dims, keep_corpus;
	const float *dims; const apo_record *corpus; uint32_t C, K; uint64_t T, Tc;
	double *scores; uint64_t *counts; int32_t *topk; apo_corpus_report report; int rc;
} score_job;

static void score_execute(napi_env env, void *data) {
	(void)env;
	score_job *j = (score_job *)data;
	apo_score_opts o; memset(&o, 0, sizeof o);
	o.K = j->K; o.source = APO_SRC_DIMS; o.flags = j->corpus ? APO_SCORE_CORPUS : 0;
	j->rc = APO_OK;
	if (j->corpus) j->rc = apo_corpus_upload(j->e, j->corpus, j->Tc, 0);
	if (j->rc == APO_OK) j->rc = apo_score_host(j->e, &o, j->dims, j->C, j->T, j->scores, j->counts, j->topk, &j->report);
}

static void score_complete(napi_env env, napi_status st, void *data) {
	(void)st;
	score_job *j = (score_job *)data;
	if (j->rc != APO_OK) {
		napi_value msg, err; napi_create_string_utf8(env, apo_last_error(j->e), NAPI_AUTO_LENGTH, &msg);
		napi_create_error(env, NULL, msg, &err); napi_reject_deferred(env, j->deferred, err);
	} else {
		napi_value out, v; void *p;
		napi_create_object(env, &out);
		napi_create_arraybuffer(env, 8u * j->C, &p, &v); memcpy(p, j->scores, 8u * j->C); napi_set_named_property(env, out, "scores", v);
		napi_create_arraybuffer(env, 8u * j->C, &p, &v); memcpy(p, j->counts, 8u * j->C); napi_set_named_property(env, out, "counts", v);
		napi_create_arraybuffer(env, 4u * j->K, &p, &v); memcpy(p, j->topk, 4u * j->K); napi_set_named_property(env, out, "topk", v);
		napi_create_arraybuffer(env, sizeof j->report, &p, &v); memcpy(p, &j->report, sizeof j->report); napi_set_named_property(env, out, "report", v);
		napi_resolve_deferred(env, j->deferred, out);
	}
	napi_delete_reference(env, j->keep_dims);
	if (j->keep_corpus) napi_delete_reference(env, j->keep_corpus);
	napi_delete_async_work(env, j->work);
	free(j->scores); free(j->counts); free(j->topk); free(j);
}

static napi_value Score(napi_env env, napi_callback_info info) {
	size_t argc = 6; napi_value argv[6], promise, name;
	napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
	score_job *j = (score_job *)calloc(1, sizeof *j);
	j->e = get_engine(env, argv[0]);
	void *p; size_t bytes; double Td = 0;
	napi_get_arraybuffer_info(env, argv[1], &p, &bytes); j->dims = (const float *)p;
	napi_get_value_uint32(env, argv[2], &j->C);
	napi_get_value_double(env, argv[3], &Td); j->T = (uint64_t)Td;
	if (argc > 4 && napi_get_arraybuffer_info(env, argv[4], &p, &bytes) == napi_ok && bytes) { j->corpus = (const apo_record *)p; j->Tc = bytes / sizeof(apo_record); }
	napi_get_value_uint32(env, argv[5], &j->K);
	j->scores = (double *)malloc(8u * (j->C ? j->C : 1)); j->counts = (uint64_t *)malloc(8u * (j->C ? j->C : 1)); j->topk = (int32_t *)malloc(4u * (j->K ? j->K : 1));
	/* the ArrayBuffers stay referenced until the worker is done (ownership rule of INTEGRATION.md) */
	napi_create_reference(env, argv[1], 1, &j->keep_dims);
	if (j->corpus) napi_create_reference(env, argv[4], 1, &j->keep_corpus);
	napi_create_promise(env, &j->deferred, &promise);
	napi_create_string_utf8(env, "apo_score", NAPI_AUTO_LENGTH, &name);
	napi_create_async_work(env, NULL, name, score_execute, score_complete, j, &j->work);
	napi_queue_async_work(env, j->work);
	return promise;
}

NAPI_MODULE_INIT() {
	const napi_property_descriptor props[] = {
	    {"create", NULL, Create, NULL, NULL, NULL, napi_default, NULL},
	    {"rewardBatch", NULL, RewardBatch, NULL, NULL, NULL, napi_default, NULL},
	    {"allocPinned", NULL, AllocPinned, NULL, NULL, NULL, napi_default, NULL},
	    {"recordsFromJson", NULL, RecordsFromJson, NULL, NULL, NULL, napi_default, NULL},
	    {"score", NULL, Score, NULL, NULL, NULL, napi_default, NULL},
	};
	napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
	return exports;
}
