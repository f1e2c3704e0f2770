// build.js -- the reference's build step (build.js:23-40: `wasm-tools parse modules/X.wat`) is
// replaced by: (1) nvcc -gencode arch=compute_100a,code=sm_100a for the CUDA engine, (2) the N-API shim.
import { execFileSync } from "child_process";
import { mkdirSync } from "fs";
import { dirname, join } from "path";
import { fileURLToPath } from "url";

const here = dirname(fileURLToPath(import.meta.url));
const root = join(here, "..");
const run = (cmd, args, cwd) => { console.log(`$ ${cmd} ${args.join(" ")}`); execFileSync(cmd, args, { cwd, stdio: "inherit" }); };

// 1. libwatfft_b200.so (every kernel, sm_100a only; see ../Makefile)
run("make", ["-C", root, "all"], root);

// 2. the addon: one translation unit + the shared library
mkdirSync(join(root, "build"), { recursive: true });
const nodeInclude = join(dirname(process.execPath), "..", "include", "node");
run("g++", ["-O2", "-fPIC", "-shared", "-std=c++17", `-I${nodeInclude}`, join(root, "napi", "watfft_napi.cc"),
  "-o", join(root, "build", "watfft_napi.node"), `-L${root}`, "-lwatfft_b200", `-Wl,-rpath,${root}`], root);
console.log("built build/watfft_napi.node");
