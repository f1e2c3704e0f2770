// index.js -- wat-fft's context surface backed by the B200 engine.
//
// Same factories and context shape as the reference's index.js (createFFT / createFFTf32 /
// createRFFT / createRFFTf32 returning {size, getInputBuffer, getOutputBuffer, forward, inverse},
// reference index.js:69-178), plus:
//   * a `batch` option (rows are contiguous per transform; batch = 1 reproduces the reference's
//     in-place views exactly),
//   * split-format factories for the flagship module (createFFTf32Split / createRFFTf32Split),
//   * ctx.dispose() (device and pinned memory are not garbage collected),
//   * an exports-shaped facade (createSplitExports) so the reference's own test suites, which call
//     raw module exports, can be pointed at the GPU by swapping their loadWasm().
// There is NO CPU fallback: every factory throws when no B200 (sm_100) device is present.  The
// WASM modules remain in the reference repository as oracle and baseline only.
//
// The heavy lifting is native/: libwatfft_b200.so (CUDA, C ABI in include/watfft_b200.h) reached
// through the N-API addon napi/watfft_napi.cc built by build.js.
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

let addon = null;
function native() {
  if (addon === null) {
    // throws if the addon is missing: the GPU engine is not optional
    addon = require(join(__dirname, "..", "build", "watfft_napi.node"));
  }
  return addon;
}

function makePlan(kind, precision, layout, size, { batch = 1, device = 0 } = {}) {
  const n = native();
  n.requireB200(device); // throws Error("... no B200 (sm_100) device available ...")
  return n.planCreate(kind, precision, layout, size, batch, device);
}

function complexContext(size, precision, options) {
  const n = native();
  const plan = makePlan(C2C, precision, INTERLEAVED, size, options);
  const TA = precision === F64 ? Float64Array : Float32Array;
  const buf = new TA(n.hostBuffer(plan, 0)); // batch * 2 * size values, pinned
  return {
    size,
    batch: options?.batch ?? 1,
    getInputBuffer: () => buf,
    getOutputBuffer: () => buf, // same bytes: in-place contract (reference index.js:78-83)
    forward() { n.exec(plan, FORWARD); },
    inverse() { n.exec(plan, INVERSE); },
    dispose() { n.planDestroy(plan); },
  };
}

function realContext(size, precision, options) {
  const n = native();
  const batch = options?.batch ?? 1;
  const plan = makePlan(R2C, precision, INTERLEAVED, size, options);
  const TA = precision === F64 ? Float64Array : Float32Array;
  const specAB = n.hostBuffer(plan, BUF_SPECTRUM);
  const spectrum = new TA(specAB); // batch * (size + 2)
  // batch = 1: the input view aliases the first `size` values of the output view, exactly like the
  // reference's two views over memory offset 0 (index.js:136-141)
  const time = batch === 1 ? new TA(specAB, 0, size) : new TA(n.hostBuffer(plan, BUF_TIME));
  return {
    size,
    batch,
    getInputBuffer: () => time,
    getOutputBuffer: () => spectrum,
    forward() { n.exec(plan, FORWARD); },
    inverse() { n.exec(plan, INVERSE); }, // f64: extension (the reference export is missing, F5)
    dispose() { n.planDestroy(plan); },
  };
}

/** f64 interleaved complex FFT (reference index.js:69-91). */
export async function createFFT(size, options) { return complexContext(size, F64, options); }
/** f32 interleaved complex FFT (reference index.js:98-120). */
export async function createFFTf32(size, options) { return complexContext(size, F32, options); }
/** f64 real FFT (reference index.js:127-149). */
export async function createRFFT(size, options) { return realContext(size, F64, options); }
/** f32 real FFT, rfft_split contract: N reals in, N/2+1 interleaved bins out, N >= 32. */
export async function createRFFTf32(size, options) { return realContext(size, F32, options); }
export const createRFFTf32Split = createRFFTf32;

/** f32 split-format complex FFT: the flagship fft_split / ifft_split path. */
export async function createFFTf32Split(size, options) {
  const n = native();
  const plan = makePlan(C2C, F32, SPLIT, size, options);
  const re = new Float32Array(n.hostBuffer(plan, 0));
  const im = new Float32Array(n.hostBuffer(plan, 1));
  return {
    size,
    batch: options?.batch ?? 1,
    getRealBuffer: () => re,
    getImagBuffer: () => im,
    getInputBuffer: () => [re, im],
    getOutputBuffer: () => [re, im],
    forward() { n.exec(plan, FORWARD); },
    inverse() { n.exec(plan, INVERSE); },
    dispose() { n.planDestroy(plan); },
  };
}

/**
 * Exports-shaped facade of fft_split_native_f32 (memory, REAL_OFFSET, IMAG_OFFSET,
 * precompute_*, fft_split, ...) over a host buffer laid out like the WAT memory map, so the
 * reference's suites run unmodified against the GPU.  REAL_OFFSET/IMAG_OFFSET are readable both
 * as numbers and as `.value` (benchmarks/lib/wat-contexts.js:70-72).
 */
export async function createSplitExports({ device = 0 } = {}) {
  const memory = { buffer: new ArrayBuffer(8 * 65536) };
  const plans = new Map();
  const get = async (key, make) => { if (!plans.has(key)) plans.set(key, await make()); return plans.get(key); };
  const off = (v) => Object.assign(Object(v), { value: v });
  const c2c = async (n, inverse) => {
    const ctx = await get(`c${n}`, () => createFFTf32Split(n, { device }));
    ctx.getRealBuffer().set(new Float32Array(memory.buffer, 0, n));
    ctx.getImagBuffer().set(new Float32Array(memory.buffer, 32768, n));
    inverse ? ctx.inverse() : ctx.forward();
    new Float32Array(memory.buffer, 0, n).set(ctx.getRealBuffer());
    new Float32Array(memory.buffer, 32768, n).set(ctx.getImagBuffer());
  };
  const real = async (n, inverse) => {
    const ctx = await get(`r${n}`, () => createRFFTf32(n, { device }));
    if (!inverse) {
      ctx.getInputBuffer().set(new Float32Array(memory.buffer, 0, n));
      ctx.forward();
      new Float32Array(memory.buffer, 0, n + 2).set(ctx.getOutputBuffer());
    } else {
      ctx.getOutputBuffer().set(new Float32Array(memory.buffer, 0, n + 2));
      ctx.inverse();
      new Float32Array(memory.buffer, 0, n).set(ctx.getInputBuffer());
    }
  };
  return {
    memory,
    REAL_OFFSET: off(0),
    IMAG_OFFSET: off(32768),
    precompute_twiddles_split: (n) => get(`c${n}`, () => createFFTf32Split(n, { device })),
    precompute_rfft_twiddles_split: (n) => get(`r${n}`, () => createRFFTf32(n, { device })),
    fft_split: (n) => c2c(n, false),
    ifft_split: (n) => c2c(n, true),
    rfft_split: (n) => real(n, false),
    irfft_split: (n) => real(n, true),
    dispose() { for (const p of plans.values()) p.dispose(); plans.clear(); },
  };
}

export function deviceCount() { return native().deviceCount(); }
