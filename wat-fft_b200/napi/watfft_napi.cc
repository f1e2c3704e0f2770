// watfft_napi.cc -- thin N-API shim over the C ABI (include/watfft_b200.h).
//
// This translation unit only translates JavaScript values to the C ABI and back; it contains no
// transform logic.  It is compiled by js/build.js where Node's headers exist (node-gyp or
// `g++ -I$(node -p "process.execPath")/../include/node`); the build image of this repository has
// neither Node nor node_api.h, so it is guarded and not part of the default build.
//
// Exposed to JS (see js/index.js):
//   deviceCount() -> number
//   requireB200(device) -> throws when no sm_100 device
//   planCreate(kind, precision, layout, n, batch, device) -> External<wfb_plan>
//   planDestroy(plan)
//   hostBuffer(plan, which) -> ArrayBuffer aliasing the plan's PINNED host memory (no copy; the
//                              ArrayBuffer is valid until planDestroy -- like memory.buffer views,
//                              index.js:78-83)
//   exec(plan, direction[, flags]) -> undefined, throws Error(wfb_strerror) on failure
#if __has_include(<node_api.h>)
#include <node_api.h>

#include <cstdint>
#include <cstdio>

#include "../../include/watfft_b200.h"

#define NAPI_OK(call)                                              \
    do {                                                           \
        if ((call) != napi_ok) {                                   \
            napi_throw_error(env, nullptr, "N-API failure: " #call); \
            return nullptr;                                        \
        }                                                          \
    } while (0)

static napi_value throw_wfb(napi_env env, int code) {
    char msg[512];
    snprintf(msg, sizeof msg, "watfft_b200: %s%s%s (code %d)", wfb_strerror(code),
             code == WFB_ERR_CUDA || code == WFB_ERR_NO_DEVICE ? " -- " : "",
             code == WFB_ERR_CUDA || code == WFB_ERR_NO_DEVICE ? wfb_last_cuda_error() : "", code);
    napi_throw_error(env, nullptr, msg);
    return nullptr;
}

static bool get_i64(napi_env env, napi_value v, int64_t *out) { return napi_get_value_int64(env, v, out) == napi_ok; }

static wfb_plan *get_plan(napi_env env, napi_value v) {
    void *p = nullptr;
    if (napi_get_value_external(env, v, &p) != napi_ok) return nullptr;
    return static_cast<wfb_plan *>(p);
}

static napi_value DeviceCount(napi_env env, napi_callback_info) {
    napi_value r;
    NAPI_OK(napi_create_int32(env, wfb_device_count(), &r));
    return r;
}

static napi_value RequireB200(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    int64_t dev = 0;
    if (argc > 0) get_i64(env, argv[0], &dev);
    int rc = wfb_require_b200((int)dev);
    if (rc != WFB_OK) return throw_wfb(env, rc);
    return nullptr;
}

static napi_value PlanCreate(napi_env env, napi_callback_info info) {
    size_t argc = 6;
    napi_value argv[6];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    int64_t a[6] = {0, 0, 0, 0, 1, 0};
    for (size_t i = 0; i < argc && i < 6; i++) get_i64(env, argv[i], &a[i]);
    int err = 0;
    wfb_plan *pl = wfb_plan_create((int)a[0], (int)a[1], (int)a[2], (int)a[3], (long)a[4], (int)a[5], &err);
    if (!pl) return throw_wfb(env, err);
    napi_value ext;
    NAPI_OK(napi_create_external(env, pl, nullptr, nullptr, &ext));   // freed explicitly by dispose()
    return ext;
}

static napi_value PlanDestroy(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    wfb_plan_destroy(get_plan(env, argv[0]));
    return nullptr;
}

static napi_value HostBuffer(napi_env env, napi_callback_info info) {
    size_t argc = 2;
    napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    wfb_plan *pl = get_plan(env, argv[0]);
    int64_t which = 0;
    get_i64(env, argv[1], &which);
    void *ptr = wfb_host_buffer(pl, (int)which);
    size_t bytes = wfb_host_bytes(pl, (int)which);
    if (!ptr) return throw_wfb(env, WFB_ERR_NO_HOST_BUFFERS);
    napi_value ab;
    // external ArrayBuffer over pinned memory: zero-copy views, never detached while the plan lives
    NAPI_OK(napi_create_external_arraybuffer(env, ptr, bytes, nullptr, nullptr, &ab));
    return ab;
}

static napi_value Exec(napi_env env, napi_callback_info info) {
    size_t argc = 3;
    napi_value argv[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    wfb_plan *pl = get_plan(env, argv[0]);
    int64_t dir = 0, flags = WFB_EXEC_DEFAULT;
    get_i64(env, argv[1], &dir);
    if (argc > 2) get_i64(env, argv[2], &flags);
    int rc = wfb_exec(pl, (int)dir, (int)flags);
    if (rc != WFB_OK) return throw_wfb(env, rc);
    return nullptr;
}

static napi_value Init(napi_env env, napi_value exports) {
    napi_property_descriptor props[] = {
        {"deviceCount", nullptr, DeviceCount, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"requireB200", nullptr, RequireB200, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"planCreate", nullptr, PlanCreate, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"planDestroy", nullptr, PlanDestroy, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"hostBuffer", nullptr, HostBuffer, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"exec", nullptr, Exec, nullptr, nullptr, nullptr, napi_default, nullptr},
    };
    napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
    return exports;
}

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
#else
// No Node headers on this machine: nothing to build (see the header comment).
#endif
