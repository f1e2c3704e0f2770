// Registry entries for the B200 backend, in the shape of the reference's benchmark surface registry
// (benchmarks/shared/wat-surfaces.mjs:51-215: name / precision / layout / minSize / maxSize / flagship), so a
// maintainer can append them to SURFACES[...].entries and list the bench file below in BENCH_COVERAGE (:217-231);
// tests/benchmark-coverage.test.js then polices the GPU surface like every other one.
//
// Differences from a WASM entry, by necessity:
//   backend: "b200"   there is no `module` (.wasm) to instantiate: `create(size, {batch, device})` returns the context
//   batch             rows per call; the WASM entries transform one row per call
// The layouts are the registry's own (LAYOUTS, :35-39), applied per row: rows are contiguous.
//
// This file cannot be exercised in the build image (no Node runtime, SURVEY F1); its Python twin
// (tests/test_reference_suites.py with the "gpu" backend) runs the same checks against the same C ABI.
import { createFFT, createFFTf32, createFFTf32Split, createRFFT, createRFFTf32 } from "./index.js";

export const B200_SURFACES = {
  "complex-forward": [
    { name: "wat-fft b200 (f32 split)", backend: "b200", precision: "f32", layout: "complex-split", create: createFFTf32Split,
      run: "forward", minSize: 4, maxSize: 8192, flagship: true },
    { name: "wat-fft b200 (f32)", backend: "b200", precision: "f32", layout: "complex-interleaved", create: createFFTf32,
      run: "forward", minSize: 4, maxSize: 8192, flagship: false },
    { name: "wat-fft b200 (f64)", backend: "b200", precision: "f64", layout: "complex-interleaved", create: createFFT,
      run: "forward", minSize: 4, maxSize: 8192, flagship: true },
  ],
  "complex-inverse": [
    { name: "wat-ifft b200 (f32 split)", backend: "b200", precision: "f32", layout: "complex-split", create: createFFTf32Split,
      run: "inverse", roundtripWith: "forward", minSize: 4, maxSize: 8192, flagship: true },
    { name: "wat-ifft b200 (f64)", backend: "b200", precision: "f64", layout: "complex-interleaved", create: createFFT,
      run: "inverse", roundtripWith: "forward", minSize: 4, maxSize: 8192, flagship: true },
  ],
  "real-forward": [
    { name: "wat-rfft b200 (f32)", backend: "b200", precision: "f32", layout: "real-packed", create: createRFFTf32,
      run: "forward", minSize: 8, maxSize: 16384, flagship: true },
    { name: "wat-rfft b200 (f64)", backend: "b200", precision: "f64", layout: "real-packed", create: createRFFT,
      run: "forward", minSize: 8, maxSize: 16384, flagship: true },
  ],
  "real-inverse": [
    { name: "wat-irfft b200 (f32)", backend: "b200", precision: "f32", layout: "real-spectrum", create: createRFFTf32,
      run: "inverse", spectrumVia: "forward", minSize: 8, maxSize: 16384, flagship: true },
  ],
};

export const B200_BENCH_COVERAGE = [
  { file: "wat-fft_b200/js/b200.bench.mjs", surface: "complex-forward" },
  { file: "wat-fft_b200/js/b200.bench.mjs", surface: "complex-inverse" },
  { file: "wat-fft_b200/js/b200.bench.mjs", surface: "real-forward" },
  { file: "wat-fft_b200/js/b200.bench.mjs", surface: "real-inverse" },
];

/** Entries of a surface that support `size` (the registry's own filter, wat-surfaces.mjs `entriesFor`). */
export function b200EntriesFor(surface, size) {
  return (B200_SURFACES[surface] || []).filter((e) => size >= e.minSize && size <= e.maxSize);
}
