/* Compile-only stand-in for Node's <node_api.h>: the declarations (real signatures, Node-API v8) of exactly the calls
 * wat-fft_b200/napi/watfft_napi.cc makes.  The build image has no Node headers; this lets the test-suite at
 * least type-check the shim.  It is never linked or loaded. */
#ifndef WFB_NAPI_STUB_H
#define WFB_NAPI_STUB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_ref__ *napi_ref;
typedef struct napi_callback_info__ *napi_callback_info;
typedef enum { napi_ok, napi_invalid_arg, napi_generic_failure } napi_status;
typedef enum { napi_default = 0 } napi_property_attributes;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);
typedef struct {
    const char *utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter;
    napi_value value; napi_property_attributes attributes; void *data;
} napi_property_descriptor;
napi_status napi_throw_error(napi_env env, const char *code, const char *msg);
napi_status napi_get_value_int64(napi_env env, napi_value value, int64_t *result);
napi_status napi_get_value_external(napi_env env, napi_value value, void **result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value *result);
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t *argc, napi_value *argv, napi_value *this_arg, void **data);
napi_status napi_create_external(napi_env env, void *data, napi_finalize finalize_cb, void *finalize_hint, napi_value *result);
napi_status napi_create_external_arraybuffer(napi_env env, void *external_data, size_t byte_length, napi_finalize finalize_cb, void *finalize_hint, napi_value *result);
napi_status napi_get_arraybuffer_info(napi_env env, napi_value arraybuffer, void **data, size_t *byte_length);
napi_status napi_detach_arraybuffer(napi_env env, napi_value arraybuffer);
napi_status napi_create_reference(napi_env env, napi_value value, uint32_t initial_refcount, napi_ref *result);
napi_status napi_delete_reference(napi_env env, napi_ref ref);
napi_status napi_get_reference_value(napi_env env, napi_ref ref, napi_value *result);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count, const napi_property_descriptor *properties);
#define NODE_GYP_MODULE_NAME watfft_napi
#define NAPI_MODULE(modname, regfunc) napi_value wfb_napi_stub_register(napi_env env, napi_value exports) { return regfunc(env, exports); }
#ifdef __cplusplus
}
#endif
#endif
