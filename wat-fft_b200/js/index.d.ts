// Types of the B200-backed surface.  Everything the reference's index.d.ts declares (index.d.ts:6-250) is declared
// here under the same name and with the same shape; additions are marked NEW.  Differences in meaning:
//   * `memory` is `{ buffer: ArrayBuffer }` over pinned, device-mapped host memory laid out like the module's linear
//     memory (same page count and offsets), not a WebAssembly.Memory: it never grows, and `buffer` is detached by dispose().
//   * every factory REJECTS when no B200 (sm_100) device is present: there is no CPU fallback.

// =============================================================================
// Low-level instance types (raw module-shaped exports; all calls are synchronous)
// =============================================================================

/** NEW: the `memory` export of an instance. */
export interface InstanceMemory {
  /** Pinned, device-mapped host bytes laid out like the WAT module's memory; stable for the life of the instance. */
  readonly buffer: ArrayBuffer;
}

/** NEW: options of every factory. */
export interface GpuOptions {
  /** rows per call (default 1 = the reference shape); only the context factories take it */
  batch?: number;
  /** CUDA device index (default 0) */
  device?: number;
}

/** Raw exports for complex FFT (f64) -- replaces fft_combined.wasm */
export interface FFTExports {
  memory: InstanceMemory;
  precompute_twiddles(n: number): void;
  fft(n: number): void;
  ifft(n: number): void;
  /** NEW: frees the plans and the pinned memory (idempotent) */
  dispose(): void;
}

/** Raw exports for complex FFT (f32) -- replaces fft_stockham_f32_dual.wasm */
export interface FFTf32Exports {
  memory: InstanceMemory;
  precompute_twiddles(n: number): void;
  fft(n: number): void;
  ifft(n: number): void;
  dispose(): void;
}

/** Raw exports for real FFT (f64) -- replaces fft_real_combined.wasm */
export interface RFFTExports {
  memory: InstanceMemory;
  precompute_rfft_twiddles(n: number): void;
  rfft(n: number): void;
  /** Provided here; the reference declares it (index.d.ts:27) but its module has no such export. */
  irfft(n: number): void;
  dispose(): void;
}

/** Raw exports for real FFT (f32) -- replaces fft_real_f32_dual.wasm */
export interface RFFTf32Exports {
  memory: InstanceMemory;
  precompute_rfft_twiddles(n: number): void;
  rfft(n: number): void;
  irfft(n: number): void;
  dispose(): void;
}

/** NEW: raw exports of the split-format module -- replaces fft_split_native_f32.wasm (modules/fft_split_native_f32.wat:60-61 for the offsets) */
export interface FFTf32SplitExports {
  memory: InstanceMemory;
  /** readable as a number and as `.value`, like a WebAssembly.Global (benchmarks/lib/wat-contexts.js:70-72) */
  REAL_OFFSET: number & { value: number };
  IMAG_OFFSET: number & { value: number };
  precompute_twiddles_split(n: number): void;
  precompute_rfft_twiddles_split(n: number): void;
  fft_split(n: number): void;
  ifft_split(n: number): void;
  rfft_split(n: number): void;
  irfft_split(n: number): void;
  dispose(): void;
}

// =============================================================================
// High-level FFT context types
// =============================================================================

/** High-level complex FFT context (f64) */
export interface FFT {
  /** FFT size (must be power of 2) */
  readonly size: number;
  /** NEW: rows per call; rows are contiguous per transform */
  readonly batch: number;

  /**
   * Get the input buffer for writing samples.
   * Format: interleaved complex [re0, im0, re1, im1, ...]
   * Length: batch * size * 2
   */
  getInputBuffer(): Float64Array;

  /**
   * Get the output buffer for reading results (the same bytes as the input buffer: in place).
   * Format: interleaved complex [re0, im0, re1, im1, ...]
   * Length: batch * size * 2
   */
  getOutputBuffer(): Float64Array;

  /** Execute forward FFT in-place (synchronous) */
  forward(): void;

  /** Execute inverse FFT in-place (synchronous, 1/size normalised) */
  inverse(): void;

  /** The module-shaped instance the context runs on (batch = 1), or an instance of its own (batch > 1) */
  readonly exports: FFTExports;

  /** NEW: frees device and pinned host memory; idempotent; any later use throws */
  /** Staging knob of the H2D -> kernel -> D2H pipeline (see OPTIONS); contexts created with batch > 1 only. */
  setOption(option: number, value: number): void;
  dispose(): void;
}

/** High-level complex FFT context (f32) */
export interface FFTf32 {
  readonly size: number;
  readonly batch: number;
  /** interleaved complex, length batch * size * 2 */
  getInputBuffer(): Float32Array;
  getOutputBuffer(): Float32Array;
  forward(): void;
  inverse(): void;
  readonly exports: FFTf32Exports;
  /** Staging knob of the H2D -> kernel -> D2H pipeline (see OPTIONS); contexts created with batch > 1 only. */
  setOption(option: number, value: number): void;
  dispose(): void;
}

/** High-level real FFT context (f64) */
export interface RFFT {
  readonly size: number;
  readonly batch: number;

  /**
   * Get the input buffer for writing real samples.
   * Length: batch * size.  At batch = 1 it starts at the same address as the output buffer, as in the reference.
   */
  getInputBuffer(): Float64Array;

  /**
   * Get the output buffer for reading complex results.
   * Format: interleaved complex [re0, im0, re1, im1, ...], (size / 2 + 1) bins per row
   * Length: batch * (size / 2 + 1) * 2
   */
  getOutputBuffer(): Float64Array;

  /** Execute forward real FFT */
  forward(): void;

  /** Execute inverse real FFT (reads the output buffer, writes the input buffer).  Works here; the reference's f64 module lacks it. */
  inverse(): void;

  readonly exports: RFFTExports;
  /** Staging knob of the H2D -> kernel -> D2H pipeline (see OPTIONS); contexts created with batch > 1 only. */
  setOption(option: number, value: number): void;
  dispose(): void;
}

/** High-level real FFT context (f32) */
export interface RFFTf32 {
  readonly size: number;
  readonly batch: number;
  /** length batch * size */
  getInputBuffer(): Float32Array;
  /** length batch * (size / 2 + 1) * 2 */
  getOutputBuffer(): Float32Array;
  forward(): void;
  inverse(): void;
  readonly exports: RFFTf32Exports;
  /** Staging knob of the H2D -> kernel -> D2H pipeline (see OPTIONS); contexts created with batch > 1 only. */
  setOption(option: number, value: number): void;
  dispose(): void;
}

/** NEW: high-level split-format complex FFT context (f32): the flagship fft_split / ifft_split path */
export interface FFTf32Split {
  readonly size: number;
  readonly batch: number;
  /** batch * size reals */
  getRealBuffer(): Float32Array;
  /** batch * size reals */
  getImagBuffer(): Float32Array;
  getInputBuffer(): [Float32Array, Float32Array];
  getOutputBuffer(): [Float32Array, Float32Array];
  forward(): void;
  inverse(): void;
  readonly exports: FFTf32SplitExports;
  /** Staging knob of the H2D -> kernel -> D2H pipeline (see OPTIONS); contexts created with batch > 1 only. */
  setOption(option: number, value: number): void;
  dispose(): void;
}

// =============================================================================
// High-level factory functions (recommended)
// =============================================================================

/**
 * Create a complex FFT context with f64 precision.
 *
 * @param size - FFT size (power of 2, 4..8192)
 * @example
 * ```ts
 * const fft = await createFFT(1024);            // or createFFT(1024, { batch: 4096 })
 * const input = fft.getInputBuffer();
 * input[0] = 1.0; input[1] = 0.0;
 * fft.forward();
 * const output = fft.getOutputBuffer();
 * fft.dispose();
 * ```
 */
export function createFFT(size: number, options?: GpuOptions): Promise<FFT>;

/** Create a complex FFT context with f32 precision (interleaved). */
export function createFFTf32(size: number, options?: GpuOptions): Promise<FFTf32>;

/** Create a real FFT context with f64 precision (size: power of 2, 8..16384). */
export function createRFFT(size: number, options?: GpuOptions): Promise<RFFT>;

/** Create a real FFT context with f32 precision (size: power of 2, 8..16384). */
export function createRFFTf32(size: number, options?: GpuOptions): Promise<RFFTf32>;

/** NEW: alias of createRFFTf32 (the f32 real transform follows the rfft_split contract) */
export function createRFFTf32Split(size: number, options?: GpuOptions): Promise<RFFTf32>;

/** NEW: split-format complex FFT context (f32) */
export function createFFTf32Split(size: number, options?: GpuOptions): Promise<FFTf32Split>;

// =============================================================================
// Low-level factory functions (advanced)
// =============================================================================

/** Create a raw instance for complex FFT (f64).  For users who need direct memory control. */
export function createFFTInstance(options?: GpuOptions): Promise<FFTExports>;

/** Create a raw instance for complex FFT (f32). */
export function createFFTf32Instance(options?: GpuOptions): Promise<FFTf32Exports>;

/** Create a raw instance for real FFT (f64). */
export function createRFFTInstance(options?: GpuOptions): Promise<RFFTExports>;

/** Create a raw instance for real FFT (f32). */
export function createRFFTf32Instance(options?: GpuOptions): Promise<RFFTf32Exports>;

/** NEW: raw instance of the split-format module */
export function createFFTf32SplitInstance(options?: GpuOptions): Promise<FFTf32SplitExports>;

/** NEW (round-1 name of createFFTf32SplitInstance) */
export function createSplitExports(options?: GpuOptions): Promise<FFTf32SplitExports>;

/** NEW: number of CUDA devices visible to the process */
export function deviceCount(): number;

/** Option ids for `ctx.setOption()` (WFB_OPT_* of include/watfft_b200.h). */
export const OPTIONS: Readonly<{ MAPPED_MAX_BYTES: 0; STAGE_CHUNK_BYTES: 1; STAGE_STREAMS: 2; STAGE_RAMP: 3 }>;
