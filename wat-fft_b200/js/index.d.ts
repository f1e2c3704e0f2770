// Types of the B200-backed surface: the reference's context shape (index.d.ts:42-153) + batch.
export interface GpuOptions { batch?: number; device?: number; }

export interface FFT<T extends Float32Array | Float64Array = Float64Array> {
  readonly size: number;
  readonly batch: number;
  /** batch * 2 * size interleaved values; same memory as getOutputBuffer() (in place). */
  getInputBuffer(): T;
  getOutputBuffer(): T;
  forward(): void;
  inverse(): void;
  /** frees device and pinned host memory (new obligation vs the WASM contexts). */
  dispose(): void;
}
export type FFTf32 = FFT<Float32Array>;

export interface RFFT<T extends Float32Array | Float64Array = Float64Array> {
  readonly size: number;
  readonly batch: number;
  /** batch * size reals.  At batch = 1 it aliases the first `size` values of the output view. */
  getInputBuffer(): T;
  /** batch * (size + 2) values: size/2+1 interleaved bins per row. */
  getOutputBuffer(): T;
  forward(): void;
  inverse(): void;
  dispose(): void;
}
export type RFFTf32 = RFFT<Float32Array>;

export interface FFTf32Split {
  readonly size: number;
  readonly batch: number;
  getRealBuffer(): Float32Array;
  getImagBuffer(): Float32Array;
  getInputBuffer(): [Float32Array, Float32Array];
  getOutputBuffer(): [Float32Array, Float32Array];
  forward(): void;
  inverse(): void;
  dispose(): void;
}

/** All factories reject/throw when no B200 (sm_100) device is present: there is no CPU fallback. */
export function createFFT(size: number, options?: GpuOptions): Promise<FFT>;
export function createFFTf32(size: number, options?: GpuOptions): Promise<FFTf32>;
export function createRFFT(size: number, options?: GpuOptions): Promise<RFFT>;
export function createRFFTf32(size: number, options?: GpuOptions): Promise<RFFTf32>;
export function createRFFTf32Split(size: number, options?: GpuOptions): Promise<RFFTf32>;
export function createFFTf32Split(size: number, options?: GpuOptions): Promise<FFTf32Split>;

export interface SplitExports {
  memory: { buffer: ArrayBuffer };
  REAL_OFFSET: number & { value: number };
  IMAG_OFFSET: number & { value: number };
  precompute_twiddles_split(n: number): Promise<unknown>;
  precompute_rfft_twiddles_split(n: number): Promise<unknown>;
  fft_split(n: number): Promise<void>;
  ifft_split(n: number): Promise<void>;
  rfft_split(n: number): Promise<void>;
  irfft_split(n: number): Promise<void>;
  dispose(): void;
}
export function createSplitExports(options?: { device?: number }): Promise<SplitExports>;
export function deviceCount(): number;
