// index.js -- wat-fft's module surface backed by the B200 engine.
//
// Drop-in for the reference's index.js: the same eight factories with the same shapes
//   createFFTInstance / createFFTf32Instance / createRFFTInstance / createRFFTf32Instance   (reference index.js:28-58)
//       -> raw module-shaped exports {memory, precompute_*, fft|rfft, ifft|irfft}
//   createFFT / createFFTf32 / createRFFT / createRFFTf32                                    (reference index.js:69-178)
//       -> {size, exports, getInputBuffer, getOutputBuffer, forward, inverse}
// and a context is BUILT the way the reference builds it: instance -> exports.precompute_*(size) -> typed views over
// exports.memory.buffer (fresh on every call, as there) -> forward() = exports.fft(size).
//
// What an "instance" is here: `memory.buffer` is an ArrayBuffer over pinned, device-mapped host memory laid out like the
// WAT module's linear memory (same page count, same offsets); every transform export is ONE SYNCHRONOUS native call that
// runs one CUDA kernel reading and writing those very bytes (no copies) and returns when the results are in memory.
// Exports are synchronous exactly like the `(param i32)` WASM exports: the reference's suites write memory, call
// `wasm.fft_split(n)` and read memory on the next line (tests/fft_split_native.test.js:78-114).  Only the factories are
// async, as in the reference.
//
// Added: a `batch` option on the context factories (rows are contiguous per transform; batch = 1 is the reference
// shape), split-format factories for the flagship module (createFFTf32SplitInstance / createFFTf32Split /
// createRFFTf32Split), and ctx.dispose() / exports.dispose() (pinned and device memory are not garbage collected).
// There is NO CPU fallback: every factory rejects when no B200 (sm_100) device is present.  The WASM modules remain in
// the reference repository as oracle and baseline only.
//
// Python twin of this file (the one the build image can execute): ../contexts.py.
// Native side: libwatfft_b200.so (CUDA, C ABI in include/watfft_b200.h) through the N-API addon napi/watfft_napi.cc.
import { createRequire } from "module";
import { fileURLToPath } from "url";
import { dirname, join } from "path";

const __dirname = dirname(fileURLToPath(import.meta.url));
const require = createRequire(import.meta.url);

const C2C = 0, R2C = 1;
const F32 = 0, F64 = 1;
const SPLIT = 0, INTERLEAVED = 1;
const FORWARD = 0, INVERSE = 1;
const BUF_TIME = 0, BUF_SPECTRUM = 1;
const PLAN_NO_HOST_BUFFERS = 1;
const PAGE = 65536;

let addon = null;
function native() {
  if (addon === null) {
    // throws if the addon is missing: the GPU engine is not optional
    addon = require(join(__dirname, "..", "build", "watfft_napi.node"));
  }
  return addon;
}

// Memory maps of the reference modules (SURVEY.md appendix A; `(memory (export "memory") N)` of each .wat)
const MODULES = {
  fft_combined: { pages: 6, f64: true },
  fft_stockham_f32_dual: { pages: 4, f64: false },
  fft_real_combined: { pages: 8, f64: true },
  fft_real_f32_dual: { pages: 6, f64: false },
  fft_split_native_f32: { pages: 8, f64: false },
};
const REAL_OFFSET = 0, IMAG_OFFSET = 32768; // modules/fft_split_native_f32.wat:60-61

/**
 * A raw module-shaped instance (synchronous to build: the native plan and memory calls are synchronous).
 * `precompute_*` (re)targets the instance at a size -- calling it with a new n silently re-targets, as in the reference
 * (tests/boundary.test.js:304-332); a transform call for a size that was never precomputed builds its plan on demand.
 */
function makeInstanceSync(moduleName, { device = 0 } = {}) {
  const n = native();
  n.requireB200(device); // throws Error("... no B200 (sm_100) device available ...")
  const { pages, f64 } = MODULES[moduleName];
  const buffer = n.hostAlloc(pages * PAGE); // ArrayBuffer over pinned, device-mapped, zeroed host memory
  const plans = new Map();
  let disposed = false;
  const alive = () => { if (disposed) throw new Error("watfft_b200: instance used after dispose()"); };
  const plan = (kind, layout, size) => {
    alive();
    const key = `${kind}:${layout}:${size}`;
    let p = plans.get(key);
    if (p === undefined) {
      p = n.planCreate(kind, f64 ? F64 : F32, layout, size, 1, device, PLAN_NO_HOST_BUFFERS);
      plans.set(key, p);
    }
    return p;
  };
  // in place on `memory`: plane 0 at off0, plane 1 (split only) at off1
  const run = (p, direction, off0, off1 = -1) => { n.execHost(p, direction, buffer, off0, off1, off0, off1); };
  const off = (v) => Object.assign(Object(v), { value: v }); // readable as a number and as `.value` (benchmarks/lib/wat-contexts.js:70-72)

  const common = {
    memory: { buffer },
    dispose() {
      if (disposed) return;
      disposed = true;
      for (const p of plans.values()) n.planDestroy(p);
      plans.clear();
      n.hostFree(buffer); // the ArrayBuffer is detached: stale views read as empty instead of dangling
    },
  };
  switch (moduleName) {
    case "fft_combined":
    case "fft_stockham_f32_dual":
      return {
        ...common,
        precompute_twiddles(size) { plan(C2C, INTERLEAVED, size); },
        fft(size) { run(plan(C2C, INTERLEAVED, size), FORWARD, 0); },
        ifft(size) { run(plan(C2C, INTERLEAVED, size), INVERSE, 0); },
      };
    case "fft_real_combined":
    case "fft_real_f32_dual":
      return {
        ...common,
        precompute_rfft_twiddles(size) { plan(R2C, INTERLEAVED, size); },
        rfft(size) { run(plan(R2C, INTERLEAVED, size), FORWARD, 0); },
        // f64: extension -- the reference's fft_real_combined has no irfft export (its index.js:145-147 calls a missing one)
        irfft(size) { run(plan(R2C, INTERLEAVED, size), INVERSE, 0); },
      };
    default: // fft_split_native_f32
      return {
        ...common,
        REAL_OFFSET: off(REAL_OFFSET),
        IMAG_OFFSET: off(IMAG_OFFSET),
        precompute_twiddles_split(size) { plan(C2C, SPLIT, size); },
        precompute_rfft_twiddles_split(size) { plan(R2C, INTERLEAVED, size); },
        fft_split(size) { run(plan(C2C, SPLIT, size), FORWARD, REAL_OFFSET, IMAG_OFFSET); },
        ifft_split(size) { run(plan(C2C, SPLIT, size), INVERSE, REAL_OFFSET, IMAG_OFFSET); },
        rfft_split(size) { run(plan(R2C, INTERLEAVED, size), FORWARD, 0); },
        irfft_split(size) { run(plan(R2C, INTERLEAVED, size), INVERSE, 0); },
      };
  }
}

// =============================================================================
// Low-level instance factories (return raw module-shaped exports) -- reference index.js:28-58
// =============================================================================
/** Raw instance for complex FFT (f64): memory, precompute_twiddles, fft, ifft. */
export async function createFFTInstance(options) { return makeInstanceSync("fft_combined", options); }
/** Raw instance for complex FFT (f32, interleaved). */
export async function createFFTf32Instance(options) { return makeInstanceSync("fft_stockham_f32_dual", options); }
/** Raw instance for real FFT (f64): memory, precompute_rfft_twiddles, rfft (+ irfft, an extension). */
export async function createRFFTInstance(options) { return makeInstanceSync("fft_real_combined", options); }
/** Raw instance for real FFT (f32): memory, precompute_rfft_twiddles, rfft, irfft. */
export async function createRFFTf32Instance(options) { return makeInstanceSync("fft_real_f32_dual", options); }
/** Raw instance of the split-format module (the reference exposes it to its tests and benchmarks only). */
export async function createFFTf32SplitInstance(options) { return makeInstanceSync("fft_split_native_f32", options); }
/** Round-1 name of createFFTf32SplitInstance. */
export const createSplitExports = createFFTf32SplitInstance;

// =============================================================================
// High-level context factories -- reference index.js:69-178
// =============================================================================
// batch = 1: the reference's construction, on an instance.  batch > 1: a plan with batch-sized pinned buffers of its own
// (rows contiguous per transform); `exports` is then a module instance of its own, created on first use -- its memory is one
// transform wide, so it cannot be the batch buffers.
function makeContext({ moduleName, kind, layout, precision, size, options, precompute, fwd, inv, views }) {
  const n = native();
  const batch = options?.batch ?? 1;
  const device = options?.device ?? 0;
  let disposed = false;
  const alive = () => { if (disposed) throw new Error("watfft_b200: context used after dispose()"); };
  let exportsObj = null;
  let plan = null;
  let bufs = null; // batch > 1: ArrayBuffers of the plan's pinned host buffers
  if (batch === 1) {
    exportsObj = makeInstanceSync(moduleName, { device });
    exportsObj[precompute](size);
  } else {
    n.requireB200(device);
    plan = n.planCreate(kind, precision, layout, size, batch, device, 0);
    bufs = [n.hostBuffer(plan, 0), layout === SPLIT || kind === R2C ? n.hostBuffer(plan, 1) : null];
  }
  const ctx = {
    size,
    batch,
    get exports() {
      alive();
      if (exportsObj === null) exportsObj = makeInstanceSync(moduleName, { device });
      return exportsObj;
    },
    /** staging knobs of wfb_exec (OPTIONS below = WFB_OPT_* of include/watfft_b200.h); contexts with batch > 1 */
    setOption(option, value) {
      alive();
      if (plan === null) throw new Error("watfft_b200: setOption applies to contexts created with batch > 1");
      n.planSetOption(plan, option, value);
    },
    forward() { alive(); if (batch === 1) exportsObj[fwd](size); else n.exec(plan, FORWARD); },
    inverse() { alive(); if (batch === 1) exportsObj[inv](size); else n.exec(plan, INVERSE); },
    dispose() { // idempotent; later use throws instead of touching freed memory
      if (disposed) return;
      disposed = true;
      if (exportsObj !== null) exportsObj.dispose();
      if (plan !== null) n.planDestroy(plan); // detaches the ArrayBuffers handed out by hostBuffer
      plan = null;
      bufs = null;
    },
  };
  // fresh typed-array views on every call, like the reference (index.js:78-83)
  const memoryOf = () => { alive(); return batch === 1 ? [exportsObj.memory.buffer, exportsObj.memory.buffer] : bufs; };
  Object.assign(ctx, views(memoryOf, batch));
  return ctx;
}

function complexContext(size, precision, options) {
  const TA = precision === F64 ? Float64Array : Float32Array;
  return makeContext({
    moduleName: precision === F64 ? "fft_combined" : "fft_stockham_f32_dual",
    kind: C2C, layout: INTERLEAVED, precision, size, options,
    precompute: "precompute_twiddles", fwd: "fft", inv: "ifft",
    views: (mem, batch) => ({
      getInputBuffer: () => new TA(mem()[0], 0, batch * 2 * size),
      getOutputBuffer: () => new TA(mem()[0], 0, batch * 2 * size), // same bytes: in-place contract
    }),
  });
}

function realContext(size, precision, options) {
  const TA = precision === F64 ? Float64Array : Float32Array;
  return makeContext({
    moduleName: precision === F64 ? "fft_real_combined" : "fft_real_f32_dual",
    kind: R2C, layout: INTERLEAVED, precision, size, options,
    precompute: "precompute_rfft_twiddles", fwd: "rfft", inv: "irfft",
    views: (mem, batch) => ({
      // batch = 1: both views start at memory offset 0, exactly like the reference (index.js:136-141);
      // batch > 1: distinct buffers (row strides differ: size vs size + 2)
      getInputBuffer: () => new TA(mem()[BUF_TIME], 0, batch * size),
      getOutputBuffer: () => new TA(mem()[BUF_SPECTRUM], 0, batch * (size + 2)),
    }),
  });
}

/** f64 interleaved complex FFT (reference index.js:69-91). */
export async function createFFT(size, options) { return complexContext(size, F64, options); }
/** f32 interleaved complex FFT (reference index.js:98-120). */
export async function createFFTf32(size, options) { return complexContext(size, F32, options); }
/** f64 real FFT (reference index.js:127-149); inverse() is an extension. */
export async function createRFFT(size, options) { return realContext(size, F64, options); }
/** f32 real FFT (reference index.js:156-178): N reals in, N/2+1 interleaved bins out, N >= 8. */
export async function createRFFTf32(size, options) { return realContext(size, F32, options); }
export const createRFFTf32Split = createRFFTf32;

/** f32 split-format complex FFT: the flagship fft_split / ifft_split path. */
export async function createFFTf32Split(size, options) {
  return makeContext({
    moduleName: "fft_split_native_f32",
    kind: C2C, layout: SPLIT, precision: F32, size, options,
    precompute: "precompute_twiddles_split", fwd: "fft_split", inv: "ifft_split",
    views: (mem, batch) => {
      const re = () => new Float32Array(mem()[0], batch === 1 ? REAL_OFFSET : 0, batch * size);
      const im = () => new Float32Array(mem()[1], batch === 1 ? IMAG_OFFSET : 0, batch * size);
      return {
        getRealBuffer: re,
        getImagBuffer: im,
        getInputBuffer: () => [re(), im()],
        getOutputBuffer: () => [re(), im()],
      };
    },
  });
}

export function deviceCount() { return native().deviceCount(); }

/** Option ids for ctx.setOption(): zero-copy threshold, chunk size / streams / ramp of the H2D -> kernel -> D2H pipeline. */
export const OPTIONS = Object.freeze({ MAPPED_MAX_BYTES: 0, STAGE_CHUNK_BYTES: 1, STAGE_STREAMS: 2, STAGE_RAMP: 3 });
