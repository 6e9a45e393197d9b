/*
 * node_api_stub.h — the subset of Node-API (N-API v9: Electron 34 -> Node 20.18) that apo_napi.c uses, declared
 * here because the engine's build image ships no node_api.h.  N-API is a stable C ABI: the real header declares the
 * same prototypes, and apo_napi.c includes it instead of this file whenever it is on the include path.
 * Purpose: keep the addon compilable (gcc -fsyntax-only -Wall -Wextra -Werror) in CI without Node.
 */
#ifndef APO_NODE_API_STUB_H
#define APO_NODE_API_STUB_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_callback_info__ *napi_callback_info;
typedef struct napi_deferred__ *napi_deferred;
typedef struct napi_async_work__ *napi_async_work;
typedef struct napi_ref__ *napi_ref;
typedef enum { napi_ok = 0, napi_invalid_arg = 1 } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function, napi_external, napi_bigint } napi_valuetype;
typedef napi_value (*napi_callback)(napi_env, napi_callback_info);
typedef void (*napi_finalize)(napi_env, void *, void *);
typedef void (*napi_async_execute_callback)(napi_env, void *);
typedef void (*napi_async_complete_callback)(napi_env, napi_status, void *);
typedef enum { napi_default = 0 } napi_property_attributes;
typedef struct { const char *utf8name; napi_value name; napi_callback method, getter, setter; napi_value value;
                 napi_property_attributes attributes; void *data; } napi_property_descriptor;
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t *, napi_value *, napi_value *, void **);
napi_status napi_typeof(napi_env, napi_value, napi_valuetype *);
napi_status napi_is_arraybuffer(napi_env, napi_value, bool *);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t *);
napi_status napi_get_value_uint32(napi_env, napi_value, uint32_t *);
napi_status napi_get_value_double(napi_env, napi_value, double *);
napi_status napi_get_value_bool(napi_env, napi_value, bool *);
napi_status napi_get_arraybuffer_info(napi_env, napi_value, void **, size_t *);
napi_status napi_create_arraybuffer(napi_env, size_t, void **, napi_value *);
napi_status napi_create_external_arraybuffer(napi_env, void *, size_t, napi_finalize, void *, napi_value *);
napi_status napi_create_object(napi_env, napi_value *);
napi_status napi_create_double(napi_env, double, napi_value *);
napi_status napi_create_int32(napi_env, int32_t, napi_value *);
napi_status napi_create_string_utf8(napi_env, const char *, size_t, napi_value *);
napi_status napi_get_null(napi_env, napi_value *);
napi_status napi_get_boolean(napi_env, bool, napi_value *);
napi_status napi_set_named_property(napi_env, napi_value, const char *, napi_value);
napi_status napi_get_named_property(napi_env, napi_value, const char *, napi_value *);
napi_status napi_has_named_property(napi_env, napi_value, const char *, bool *);
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
#define NAPI_AUTO_LENGTH ((size_t)-1)
#define NAPI_MODULE_INIT() napi_value napi_register_module_v1(napi_env env, napi_value exports)
#endif
