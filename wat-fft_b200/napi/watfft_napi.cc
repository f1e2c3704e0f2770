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
//   planCreate(kind, precision, layout, n, batch, device[, flags]) -> External<Plan>
//   planDestroy(plan)             idempotent; detaches the ArrayBuffers handed out by hostBuffer(plan, ..)
//   hostBuffer(plan, which) -> ArrayBuffer aliasing the plan's PINNED host memory (no copy; valid
//                              until planDestroy, which detaches it -- like memory.buffer views,
//                              index.js:78-83)
//   exec(plan, direction[, flags]) -> undefined, throws Error(wfb_strerror) on failure
//   hostAlloc(bytes) -> ArrayBuffer over pinned, device-mapped, zeroed host memory (wfb_host_alloc): `exports.memory.buffer`
//   hostFree(arrayBuffer)          detaches the ArrayBuffer and returns the memory
//   execHost(plan, direction, arrayBuffer, inOff0, inOff1, outOff0, outOff1[, flags])
//                                 runs the plan on plane offsets inside a hostAlloc'ed buffer (offset -1 = no such plane);
//                                 SYNCHRONOUS: results are in the buffer when it returns
//   planSetOption(plan, option, value), planLastPath(plan)
// Every call that takes a plan throws after planDestroy (the external carries a small heap box whose pointer is
// nulled on destroy: no use-after-free, no double free).
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

// The external handed to JS is a box, not the plan: planDestroy nulls box->plan, so a stale handle is detectable.
struct PlanBox {
    wfb_plan *plan;
    napi_ref buffers[2];      // ArrayBuffers handed out by hostBuffer (detached on destroy)
};

static PlanBox *get_box(napi_env env, napi_value v) {
    void *p = nullptr;
    if (napi_get_value_external(env, v, &p) != napi_ok) return nullptr;
    return static_cast<PlanBox *>(p);
}

// live plan or nullptr (after throwing)
static wfb_plan *get_plan(napi_env env, napi_value v) {
    PlanBox *b = get_box(env, v);
    if (!b || !b->plan) {
        napi_throw_error(env, nullptr, "watfft_b200: plan used after dispose()");
        return nullptr;
    }
    return b->plan;
}

static void box_finalize(napi_env, void *data, void *) {
    PlanBox *b = static_cast<PlanBox *>(data);
    if (b->plan) wfb_plan_destroy(b->plan);     // a context that was garbage collected without dispose()
    delete b;
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
    size_t argc = 7;
    napi_value argv[7];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    int64_t a[7] = {0, 0, 0, 0, 1, 0, 0};
    for (size_t i = 0; i < argc && i < 7; i++) get_i64(env, argv[i], &a[i]);
    int err = 0;
    wfb_plan *pl = wfb_plan_create_ex((int)a[0], (int)a[1], (int)a[2], (int)a[3], (long)a[4], (int)a[5], (int)a[6], &err);
    if (!pl) return throw_wfb(env, err);
    PlanBox *box = new PlanBox{pl, {nullptr, nullptr}};
    napi_value ext;
    if (napi_create_external(env, box, box_finalize, nullptr, &ext) != napi_ok) {
        wfb_plan_destroy(pl);
        delete box;
        napi_throw_error(env, nullptr, "N-API failure: napi_create_external");
        return nullptr;
    }
    return ext;
}

static napi_value PlanDestroy(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    PlanBox *b = get_box(env, argv[0]);
    if (!b || !b->plan) return nullptr;                      // idempotent
    for (int i = 0; i < 2; i++)
        if (b->buffers[i]) {                                 // views over the pinned buffers must not outlive them
            napi_value ab;
            if (napi_get_reference_value(env, b->buffers[i], &ab) == napi_ok && ab) napi_detach_arraybuffer(env, ab);
            napi_delete_reference(env, b->buffers[i]);
            b->buffers[i] = nullptr;
        }
    wfb_plan_destroy(b->plan);
    b->plan = nullptr;
    return nullptr;
}

static napi_value HostBuffer(napi_env env, napi_callback_info info) {
    size_t argc = 2;
    napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    wfb_plan *pl = get_plan(env, argv[0]);
    if (!pl) return nullptr;
    PlanBox *box = get_box(env, argv[0]);
    int64_t which = 0;
    get_i64(env, argv[1], &which);
    if (which < 0 || which > 1) return throw_wfb(env, WFB_ERR_BAD_ARG);
    napi_value ab;
    if (box->buffers[which]) {                               // one ArrayBuffer per pinned buffer
        NAPI_OK(napi_get_reference_value(env, box->buffers[which], &ab));
        if (ab) return ab;
    }
    void *ptr = wfb_host_buffer(pl, (int)which);
    size_t bytes = wfb_host_bytes(pl, (int)which);
    if (!ptr) return throw_wfb(env, WFB_ERR_NO_HOST_BUFFERS);
    // external ArrayBuffer over pinned memory: zero-copy views, detached by planDestroy
    NAPI_OK(napi_create_external_arraybuffer(env, ptr, bytes, nullptr, nullptr, &ab));
    NAPI_OK(napi_create_reference(env, ab, 1, &box->buffers[which]));
    return ab;
}

static napi_value Exec(napi_env env, napi_callback_info info) {
    size_t argc = 3;
    napi_value argv[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    wfb_plan *pl = get_plan(env, argv[0]);
    if (!pl) return nullptr;
    int64_t dir = 0, flags = WFB_EXEC_DEFAULT;
    get_i64(env, argv[1], &dir);
    if (argc > 2) get_i64(env, argv[2], &flags);
    int rc = wfb_exec(pl, (int)dir, (int)flags);
    if (rc != WFB_OK) return throw_wfb(env, rc);
    return nullptr;
}

static napi_value HostAlloc(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    int64_t bytes = 0;
    get_i64(env, argv[0], &bytes);
    if (bytes <= 0) return throw_wfb(env, WFB_ERR_BAD_ARG);
    void *p = wfb_host_alloc((size_t)bytes);
    if (!p) return throw_wfb(env, wfb_require_b200(0) != WFB_OK ? WFB_ERR_NO_DEVICE : WFB_ERR_ALLOC);
    napi_value ab;
    // freed by hostFree (instance.dispose()); the finalizer covers instances that were garbage collected without it
    if (napi_create_external_arraybuffer(env, p, (size_t)bytes, [](napi_env, void *data, void *) { wfb_host_free(data); }, nullptr, &ab) != napi_ok) {
        wfb_host_free(p);
        napi_throw_error(env, nullptr, "N-API failure: napi_create_external_arraybuffer");
        return nullptr;
    }
    return ab;
}

static napi_value HostFree(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    void *p = nullptr;
    size_t len = 0;
    if (napi_get_arraybuffer_info(env, argv[0], &p, &len) != napi_ok || !p) return nullptr;   // already detached
    napi_detach_arraybuffer(env, argv[0]);     // stale views read as empty instead of dangling
    wfb_host_free(p);                          // (wfb_host_free ignores pointers it does not own: the finalizer's later call is a no-op)
    return nullptr;
}

static napi_value ExecHost(napi_env env, napi_callback_info info) {
    size_t argc = 8;
    napi_value argv[8];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    if (argc < 7) return throw_wfb(env, WFB_ERR_BAD_ARG);
    wfb_plan *pl = get_plan(env, argv[0]);
    if (!pl) return nullptr;
    int64_t dir = 0, off[4] = {-1, -1, -1, -1}, flags = WFB_SYNC;
    get_i64(env, argv[1], &dir);
    void *base = nullptr;
    size_t len = 0;
    if (napi_get_arraybuffer_info(env, argv[2], &base, &len) != napi_ok || !base) return throw_wfb(env, WFB_ERR_BAD_ARG);
    for (int i = 0; i < 4; i++) get_i64(env, argv[3 + i], &off[i]);
    if (argc > 7) get_i64(env, argv[7], &flags);
    const void *in[2] = {nullptr, nullptr};
    void *out[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; i++) {
        if (off[i] >= 0 && (size_t)off[i] < len) in[i] = static_cast<char *>(base) + off[i];
        if (off[2 + i] >= 0 && (size_t)off[2 + i] < len) out[i] = static_cast<char *>(base) + off[2 + i];
    }
    int rc = wfb_exec_host(pl, (int)dir, in, out, (int)flags);      // (range-checks the planes against the allocation)
    if (rc != WFB_OK) return throw_wfb(env, rc);
    return nullptr;
}

static napi_value PlanSetOption(napi_env env, napi_callback_info info) {
    size_t argc = 3;
    napi_value argv[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    wfb_plan *pl = get_plan(env, argv[0]);
    if (!pl) return nullptr;
    int64_t opt = 0, val = 0;
    get_i64(env, argv[1], &opt);
    get_i64(env, argv[2], &val);
    int rc = wfb_plan_set_option(pl, (int)opt, (long)val);
    if (rc != WFB_OK) return throw_wfb(env, rc);
    return nullptr;
}

static napi_value PlanLastPath(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    wfb_plan *pl = get_plan(env, argv[0]);
    if (!pl) return nullptr;
    napi_value r;
    NAPI_OK(napi_create_int32(env, wfb_plan_last_path(pl), &r));
    return r;
}

static napi_value Init(napi_env env, napi_value exports) {
    napi_property_descriptor props[] = {
        {"deviceCount", nullptr, DeviceCount, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"requireB200", nullptr, RequireB200, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"planCreate", nullptr, PlanCreate, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"planDestroy", nullptr, PlanDestroy, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"hostBuffer", nullptr, HostBuffer, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"exec", nullptr, Exec, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"hostAlloc", nullptr, HostAlloc, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"hostFree", nullptr, HostFree, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"execHost", nullptr, ExecHost, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"planSetOption", nullptr, PlanSetOption, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"planLastPath", nullptr, PlanLastPath, nullptr, nullptr, nullptr, napi_default, nullptr},
    };
    napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
    return exports;
}

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
#else
// No Node headers on this machine: nothing to build (see the header comment).
#endif
