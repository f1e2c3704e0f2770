"""CPU-side checks of the boundary: the library loads, exports exactly the symbols
include/watfft_b200.h declares, validates sizes without a GPU, builds reference-exact twiddles,
and refuses to run without a B200 (no CPU fallback)."""
import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "watfft_b200.h").read_text()


def declared_symbols():
    return sorted(set(re.findall(r"WFB_API\s+[\w\s\*]+?\b(wfb_\w+)\s*\(", HEADER)))


def test_header_and_binding_agree(wf):
    decl = declared_symbols()
    assert len(decl) >= 20
    assert sorted(wf._cabi.SYMBOLS) == decl


def test_library_exports_every_declared_symbol(wf):
    out = subprocess.check_output(["nm", "-D", "--defined-only", str(wf._cabi.LIB_PATH)], text=True)
    exported = set(re.findall(r"\sT\s+(wfb_\w+)", out))
    assert set(declared_symbols()) <= exported
    lib = wf._cabi.lib()
    for name in declared_symbols():
        assert getattr(lib, name)
    # nothing from the oracle is linked into the product
    assert "wfo_" not in out and "watref_" not in out


def test_size_ranges(wf):
    C = wf._cabi
    lib = C.lib()
    lo, hi = ctypes.c_int(), ctypes.c_int()
    assert lib.wfb_size_range(C.C2C, C.F32, C.SPLIT, lo, hi) == 0 and (lo.value, hi.value) == (4, 8192)
    assert lib.wfb_size_range(C.C2C, C.F32, C.INTERLEAVED, lo, hi) == 0 and (lo.value, hi.value) == (4, 8192)
    assert lib.wfb_size_range(C.R2C, C.F32, 0, lo, hi) == 0 and (lo.value, hi.value) == (8, 16384)
    assert lib.wfb_size_range(C.C2C, C.F64, C.INTERLEAVED, lo, hi) == 0
    assert lib.wfb_size_range(C.C2C, C.F64, C.SPLIT, lo, hi) == C.ERR_UNSUPPORTED
    assert lib.wfb_size_range(7, C.F32, 0, lo, hi) == C.ERR_UNSUPPORTED


def test_strerror_covers_codes(wf):
    lib = wf._cabi.lib()
    msgs = {lib.wfb_strerror(c).decode() for c in range(0, -8, -1)}
    assert len(msgs) == 8 and "no CPU fallback" in lib.wfb_strerror(-1).decode()


def test_plan_validation_order(wf):
    """Bad sizes are rejected before the device is touched, so this holds with or without a GPU."""
    C = wf._cabi
    for n in (0, 3, 12, 1000, 2, 16384):
        with pytest.raises(wf.WatFFTError) as e:
            wf.Plan(C.C2C, C.F32, C.SPLIT, n)
        assert e.value.code == C.ERR_BAD_SIZE
    with pytest.raises(wf.WatFFTError) as e:
        wf.Plan(C.R2C, C.F32, 0, 4)           # real f32 contexts start at n = 8
    assert e.value.code == C.ERR_BAD_SIZE
    with pytest.raises(wf.WatFFTError) as e:
        wf.Plan(C.C2C, C.F32, C.SPLIT, 64, batch=0)
    assert e.value.code == C.ERR_BAD_ARG


def test_no_cpu_fallback(wf):
    """Without a B200 every factory throws WFB_ERR_NO_DEVICE (north_star: the GPU context throws)."""
    lib = wf._cabi.lib()
    if lib.wfb_device_count() > 0 and lib.wfb_require_b200(0) == 0:
        pytest.skip("a B200 is present")
    for factory in (wf.createFFT, wf.createFFTf32, wf.createRFFT, wf.createRFFTf32, wf.createFFTf32Split):
        with pytest.raises(wf.WatFFTError) as e:
            factory(1024)
        assert e.value.code == wf._cabi.ERR_NO_DEVICE
    assert lib.wfb_kernel_launch_count() == 0


@pytest.mark.parametrize("n", [8, 16, 64, 1024, 4096, 8192])
def test_product_twiddles_bit_exact_with_oracle(wf, oracle, n):
    """The product's own host trig (csrc/wfb_twiddle.h) must reproduce the reference tables bit for bit;
    the oracle's tables are themselves pinned to the reference modules (test_oracle_pinning.py)."""
    lib = wf._cabi.lib()
    for flavour, kind, dt in ((0, "split", np.float32), (1, "dual", np.float32), (2, "f64", np.float64)):
        re, im = np.zeros(n, dt), np.zeros(n, dt)
        assert lib.wfb_reference_twiddles(flavour, n, n, re.ctypes.data, im.ctypes.data) == 0
        ore, oim = oracle.twiddles(kind, n)
        assert np.array_equal(re, ore) and np.array_equal(im, oim), (kind, n)


def test_twiddle_tables_match_reference_module_memory(wf, watref):
    """...and directly against the bytes the reference's precompute export writes."""
    lib = wf._cabi.lib()
    n = 1024
    watref.call("fft_split_native_f32", "precompute_twiddles_split", n)
    ref_re = watref.view("fft_split_native_f32", np.float32, 0x20000, n).copy()
    ref_im = watref.view("fft_split_native_f32", np.float32, 0x28000, n).copy()
    re, im = np.zeros(n, np.float32), np.zeros(n, np.float32)
    lib.wfb_reference_twiddles(0, n, n, re.ctypes.data, im.ctypes.data)
    assert np.array_equal(re, ref_re) and np.array_equal(im, ref_im)
    watref.call("fft_combined", "precompute_twiddles", n)
    t = watref.view("fft_combined", np.float64, 262144, 2 * n).copy()
    re, im = np.zeros(n), np.zeros(n)
    lib.wfb_reference_twiddles(2, n, n, re.ctypes.data, im.ctypes.data)
    assert np.array_equal(re, t[0::2]) and np.array_equal(im, t[1::2])


def test_f64_rfft_table_mirror_symmetry(wf, watref):
    """The f64 r2c kernels form T[M-k] from T[k] in registers (ld_tw_mirror in wfb_kernels.cuh) instead of loading the
    reference's separately tabulated entry (fft_real_combined.wat:502-503,533-534).  That is sound because the reference's
    range reduction makes the two entries mirror images to within one rounding -- checked here on the bytes the reference's
    own precompute_rfft_twiddles writes, for every size, k = M/2 (which pairs with itself and keeps its entry) excluded."""
    lib = wf._cabi.lib()
    worst = 0.0
    for n in [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]:
        m = n // 2
        re, im = np.zeros(m + 1), np.zeros(m + 1)
        lib.wfb_reference_twiddles(2, n, m + 1, re.ctypes.data, im.ctypes.data)
        if n <= 8192:                                          # (the module's 8 pages end where the N = 16384 table would)
            watref.call("fft_real_combined", "precompute_rfft_twiddles", n)
            t = watref.view("fft_real_combined", np.float64, 393216, 2 * (m + 1)).copy()
            assert np.array_equal(re, t[0::2]) and np.array_equal(im, t[1::2]), n
        k = np.arange(1, m)
        k = k[k != m // 2]
        worst = max(worst, float(np.max(np.abs(re[m - k] + re[k]))), float(np.max(np.abs(im[m - k] - im[k]))))
    assert worst <= 4.5e-16, worst             # measured: 3.3e-16 (1.5 ulp of 1.0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under wat-fft_b200/ may reference it."""
    for p in (ROOT / "wat-fft_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cc", ".js", ".mjs") and p.is_file():
            text = p.read_text(errors="ignore")
            assert "libwatfft_oracle" not in text and "watfft_oracle.c" not in text and "libwatref" not in text, p
            assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), p


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (what cgo / N-API / ctypes-style binders consume),
    and a C translation unit must link against the library's exported symbols."""
    import subprocess
    root = Path(__file__).resolve().parent.parent
    src = tmp_path / "use_header.c"
    src.write_text(
        '#include "watfft_b200.h"\n'
        "int main(void) {\n"
        "    int lo = 0, hi = 0;\n"
        "    if (wfb_size_range(WFB_C2C, WFB_F32, WFB_SPLIT, &lo, &hi) != WFB_OK) return 1;\n"
        "    return (lo == 4 && hi == 8192 && wfb_strerror(WFB_ERR_NO_DEVICE)[0]) ? 0 : 2;\n"
        "}\n")
    exe = tmp_path / "use_header"
    lib_dir = root / "wat-fft_b200"
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", str(root / "include"), str(src), "-o", str(exe),
                        "-L", str(lib_dir), "-lwatfft_b200", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)       # pure host logic: runs without a GPU
    assert r.returncode == 0, (r.returncode, r.stderr)



def test_stage_schedule_properties(wf):
    """The chunk schedule of wfb_exec's H2D / kernel / D2H pipeline (pure host logic, exported for this test): for any
    batch, row width, chunk size and ramp setting the chunks tile the batch exactly; every chunk start is 16-byte aligned for
    the (n+2)-wide spectrum rows (all chunks but the last hold an even number of rows); steady-state chunks hold whole kernel
    tiles; and with the ramp on, the first and the last three chunks are the 1/8, 1/4, 1/2 steps around full-size ones."""
    import ctypes
    lib = wf._cabi.lib()
    rng = np.random.default_rng(11)
    out = (ctypes.c_long * 4096)()
    seen_ramped = seen_flat = seen_single = 0
    for _ in range(4000):
        row = int(rng.choice([8, 64, 264, 1024, 4104, 16392, 65536, 131088]))
        chunk_bytes = int(rng.choice([4096, 32 << 10, 1 << 20, 32 << 20]))
        batch = int(rng.integers(1, 200000)) if rng.random() < 0.8 else int(2 ** rng.integers(0, 24))
        ramp = int(rng.integers(0, 2))
        n = lib.wfb_stage_schedule(batch, row, chunk_bytes, ramp, out, 4096)
        chunk = max(1, chunk_bytes // row)
        chunk = chunk & ~255 if chunk >= 512 else (chunk & ~1 if chunk >= 2 else chunk)
        if batch <= 2 * chunk:
            assert n == 0
            seen_single += 1
            continue
        if n > 4096:
            continue                                          # (tiny chunks of a huge batch: only the count is returned)
        rows = list(out[:n])
        assert n >= 3 and sum(rows) == batch and min(rows) >= 1, (batch, row, chunk_bytes, ramp, rows[:8])
        if chunk >= 2:
            assert all(r % 2 == 0 for r in rows[:-1]), (batch, row, chunk_bytes, ramp)
        assert max(rows) <= chunk + 1
        ramped = ramp and chunk >= 16 and rows[0] < chunk
        if ramped:
            seen_ramped += 1
            up = rows[:3]
            assert up[0] <= up[1] <= up[2] <= chunk and up[2] >= chunk // 2 - 255
            down = rows[-3:]
            assert down[0] >= down[1] >= down[2] - 1
            assert rows[3] == chunk                           # at least two full chunks in the middle
            if chunk >= 512:
                assert all(r % 256 == 0 for r in rows[3:-4] if r == chunk)
        else:
            seen_flat += 1
            assert all(r == chunk for r in rows[:-1])
    assert seen_ramped > 200 and seen_flat > 200 and seen_single > 50
    assert lib.wfb_stage_schedule(0, 8, 4096, 1, out, 4096) < 0
    assert lib.wfb_stage_schedule(10, 0, 4096, 1, out, 4096) < 0
