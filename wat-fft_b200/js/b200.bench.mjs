// Node benchmark of the B200 backend over the registry surfaces (b200-surfaces.mjs), following the reference's
// harness conventions: mulberry32 inputs seeded with N (benchmarks/lib/harness.js:99-108, wat-contexts.js:34-68),
// >= 10 warm-up calls, 10 samples of >= 150 ms each, median reported (harness.js:27-32,70-76).  Like the reference's
// contexts (wat-contexts.js:1-10) the timed call includes staging: forward()/inverse() copy the pinned host buffers
// to the device, transform and copy back, synchronously.
//
//   node wat-fft_b200/js/b200.bench.mjs [--sizes 256,1024,4096] [--bytes 268435456]
//
// Not runnable in the build image (no Node, SURVEY F1); bench.py is the executable twin (same C ABI, same inputs).
import { B200_SURFACES, b200EntriesFor } from "./b200-surfaces.mjs";

function mulberry32(seed) {
  let s = seed >>> 0;
  return () => {
    s = (s + 0x6d2b79f5) >>> 0;
    let t = Math.imul(s ^ (s >>> 15), s | 1);
    t ^= t + Math.imul(t ^ (t >>> 7), t | 61);
    return ((t ^ (t >>> 14)) >>> 0) / 4294967296;
  };
}

function fill(view, rnd) {
  for (let i = 0; i < view.length; i++) view[i] = rnd() * 2 - 1;
}

function median(xs) {
  const s = [...xs].sort((a, b) => a - b);
  return s[s.length >> 1];
}

// `restore` puts the pristine input back between calls (the transforms are in place; repeating one direction would
// overflow).  It runs outside the timed intervals: the reference charges its per-row `.set()` to the call
// (wat-contexts.js:125-129), which for a 256 MiB batch would measure the host memcpy, not the transform path.
function timeCalls(fn, restore) {
  for (let i = 0; i < 10; i++) { fn(); restore(); }
  const samples = [];
  for (let s = 0; s < 10; s++) {
    let calls = 0;
    let busy = 0;
    do {
      const t0 = process.hrtime.bigint();
      fn();
      busy += Number(process.hrtime.bigint() - t0) / 1e9;
      calls++;
      restore();
    } while (busy < 0.15);
    samples.push(calls / busy);
  }
  return median(samples);
}

const args = process.argv.slice(2);
const opt = (name, dflt) => {
  const i = args.indexOf(name);
  return i >= 0 ? args[i + 1] : dflt;
};
const sizes = opt("--sizes", "16,64,256,1024,4096").split(",").map(Number);
const inputBytes = Number(opt("--bytes", String(1 << 28)));

for (const surface of Object.keys(B200_SURFACES)) {
  for (const size of sizes) {
    for (const e of b200EntriesFor(surface, size)) {
      const elem = e.precision === "f64" ? 8 : 4;
      const perRow = (e.layout.startsWith("complex") ? 2 : 1) * elem * size;
      const batch = Math.max(1, Math.floor(inputBytes / perRow));
      const ctx = await e.create(size, { batch });
      const rnd = mulberry32(size);
      if (e.layout === "complex-split") {
        // the reference draws re then im per element (wat-contexts.js:47-55)
        const re = ctx.getRealBuffer(), im = ctx.getImagBuffer();
        for (let i = 0; i < re.length; i++) { re[i] = rnd() * 2 - 1; im[i] = rnd() * 2 - 1; }
      } else {
        fill(ctx.getInputBuffer(), rnd);
      }
      if (e.spectrumVia) ctx[e.spectrumVia](); // real-inverse: a valid Hermitian spectrum as input
      // pristine copies of whatever the timed direction reads
      const views = e.layout === "complex-split" ? [ctx.getRealBuffer(), ctx.getImagBuffer()]
        : [e.layout === "real-spectrum" ? ctx.getOutputBuffer() : ctx.getInputBuffer()];
      const saved = views.map((v) => v.slice());
      const restore = () => views.forEach((v, i) => v.set(saved[i]));
      const callsPerSec = timeCalls(() => ctx[e.run](), restore);
      console.log(JSON.stringify({ surface, name: e.name, size, batch, transforms_per_s: callsPerSec * batch }));
      ctx.dispose();
    }
  }
}
