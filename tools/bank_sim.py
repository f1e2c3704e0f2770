#!/usr/bin/env python3
"""Shared-memory bank-conflict simulator for the exchange patterns of wfb_kernels.cuh.

For a plan (N, T, pass codes) and a pad rule (one pad slot per PADQ complex elements) it replays
every warp-wide smem write (tid + e*T) and read (Rp*s'*j + t' + k*s') of every exchange and reports
the worst wavefront multiplicity per instruction.  Model: 32 banks x 4 B; an access of W bytes per
thread is served in phases of 128/W threads (16 for float2, 8 for double2); within a phase the cost
is the max number of distinct 4-byte words that fall in one bank.
"""
import sys
from itertools import product


def nsub(code):
    n = 0
    while code:
        n += 1
        code >>= 4
    return n


def radix(code, q):
    return (code >> (4 * (nsub(code) - 1 - q))) & 0xF


def rp(code):
    r = 1
    while code:
        r *= code & 0xF
        code >>= 4
    return r


def slot_to_out(code, k):
    g, w, mult, out = nsub(code), rp(code), 1, 0
    for a in range(g):
        r = radix(code, a)
        w //= r
        d = (k // w) % r
        out += d * mult
        mult *= r
    return out


def conflicts(addrs_words, width_words):
    """addrs_words: list of 32 starting word addresses (one per lane); returns wavefronts needed."""
    per_phase = 32 // width_words if width_words > 1 else 32
    per_phase = {1: 32, 2: 16, 4: 8}[width_words]
    total = 0
    for ph in range(0, 32, per_phase):
        banks = {}
        for a in addrs_words[ph:ph + per_phase]:
            if a is None:
                continue
            for w in range(width_words):
                banks.setdefault((a + w) % 32, set()).add(a + w)
        total += max((len(v) for v in banks.values()), default=0)
    return total, 32 // per_phase


def simulate(N, T, codes, padq, elem_words, X=None, verbose=False):
    E = N // T
    X = X or max(1, 256 // T)
    S = N + (N // padq if padq else 0)
    pad = (lambda p: p + p // padq) if padq else (lambda p: p)
    worst_w, worst_r = 1.0, 1.0
    nthreads = T * X
    lin = 1
    for P in range(len(codes) - 1):
        code = codes[P]
        RP = rp(code)
        NB = E // RP
        lin *= RP
        # writes
        for i, k in product(range(NB), range(RP)):
            e = i + NB * slot_to_out(code, k)
            for w0 in range(0, nthreads, 32):
                addrs = []
                for lane in range(32):
                    th = w0 + lane
                    xi, tid = th // T, th % T
                    addrs.append((xi * S + pad(tid + e * T)) * elem_words)
                c, ideal = conflicts(addrs, elem_words)
                worst_w = max(worst_w, c / ideal)
        code2 = codes[P + 1]
        RP2 = rp(code2)
        NB2 = E // RP2
        SP2 = N // (lin * RP2)
        for i, k in product(range(NB2), range(RP2)):
            for w0 in range(0, nthreads, 32):
                addrs = []
                for lane in range(32):
                    th = w0 + lane
                    xi, tid = th // T, th % T
                    b = tid + i * T
                    j, t = b // SP2, b % SP2
                    addrs.append((xi * S + pad(RP2 * SP2 * j + t + k * SP2)) * elem_words)
                c, ideal = conflicts(addrs, elem_words)
                worst_r = max(worst_r, c / ideal)
    return worst_w, worst_r


PLANS_F32 = {
    32: (2, [0x2, 0x44]), 64: (4, [0x4, 0x44]), 128: (8, [0x24, 0x44]), 256: (16, [0x44, 0x44]),
    512: (32, [0x2, 0x44, 0x44]), 1024: (64, [0x4, 0x44, 0x44]), 2048: (128, [0x24, 0x44, 0x44]),
    4096: (256, [0x44, 0x44, 0x44]), 8192: (512, [0x2, 0x44, 0x44, 0x44]),
}
PLANS_F64 = {
    32: (2, [0x2, 0x2222]), 64: (4, [0x4, 0x44]), 128: (8, [0x222, 0x2222]), 256: (16, [0x44, 0x44]),
    512: (32, [0x2, 0x2222, 0x2222]), 1024: (64, [0x4, 0x44, 0x44]), 2048: (128, [0x222, 0x2222, 0x2222]),
    4096: (256, [0x44, 0x44, 0x44]), 8192: (512, [0x2, 0x2222, 0x2222, 0x2222]),
}

# the wider plans of csrc/wfb_registry.h: (N, T, pass codes, pad quantum, words per element)
EXTRA_PLANS = {
    "P32_512": (512, 16, [0x244, 0x44], 16, 2), "P32_1024": (1024, 32, [0x244, 0x244], 32, 2),
    "P32_8192": (8192, 256, [0x244, 0x44, 0x44], 16, 2),
    "P64_4096": (4096, 64, [0x444, 0x444], 64, 2), "P64_2048": (2048, 32, [0x244, 0x444], 64, 2),
    "P64_1024": (1024, 16, [0x44, 0x444], 64, 2),
    "D32_512": (512, 16, [0x22222, 0x2222], 16, 4),
}


def registered_plans():
    """every multi-pass plan with the pad quantum its kernels are instantiated with"""
    out = {}
    for n, (t, codes) in PLANS_F32.items():
        out[f"F32_{n}"] = (n, t, codes, 16, 2)
    for n, (t, codes) in PLANS_F64.items():
        out[f"F64_{n}"] = (n, t, codes, 16, 4)
    out.update(EXTRA_PLANS)
    return out


# ----------------------------------------------------------------------------------------
# Tile I/O: the accesses with which the threads read a staged tile into registers (and write the result tile back).
# The exchange patterns above are private to one transform; these depend on how the ROWS of a tile lie in the stage
# buffer relative to each other, because with T < 32 threads per transform a warp touches several rows at once.
# ----------------------------------------------------------------------------------------
def c2c_tile_rows(N, T, X, RG):
    """k_c2c_pipe: thread group xi -> (row of the tile, element offset of that row in its plane).
    RG = 0: dense tile; RG > 0: groups of RG rows, (T if T < 32 else N/16) elements of padding behind each group, the
    32/T thread groups of a warp in different groups."""
    out = []
    for xi in range(X):
        if RG:
            ngr = X // RG
            xr = (xi % ngr) * RG + xi // ngr
            gstr = RG * N + (T if T < 32 else N // 16)
            out.append((xr, (xr // RG) * gstr + (xr % RG) * N))
        else:
            out.append((xi, xi * N))
    return out


def c2c_tile_io(N, T, X, RG, layout):
    """worst wavefront multiplicity of the input reads / result writes of a k_c2c_pipe tile.
    layout 'split': one 32-bit access per plane and element; 'il': one 64-bit access per element."""
    E = N // T
    rows = c2c_tile_rows(N, T, X, RG)
    assert sorted(r for r, _ in rows) == list(range(X))
    words = 1 if layout == "split" else 2
    worst = 1.0
    for e in range(E):
        for w0 in range(0, T * X, 32):
            addrs = []
            for lane in range(32):
                th = w0 + lane
                if th >= T * X:
                    addrs.append(None)
                    continue
                xi, tid = th // T, th % T
                addrs.append((rows[xi][1] + tid + e * T) * words)
            c, ideal = conflicts(addrs, words)
            worst = max(worst, c / ideal)
    return worst


def real_tile_rows(M, T, X, elem_bytes, grouped):
    """k_real_pipe with bulk stores: thread group xi -> (row, offset of the row on the DENSE side in bins).  With fewer
    than a phase of lanes per transform the groups sharing a phase take rows T apart (ROWMAP); `grouped`: the dense side in
    groups of T rows with 64 bytes of padding behind each."""
    phl = 128 // (2 * elem_bytes)
    out = []
    for xi in range(X):
        xr = xi
        if T < phl and X % phl == 0:
            gp = phl // T
            q = xi // gp
            xr = (q // T) * phl + q % T + T * (xi % gp)
        if grouped:
            gstr = T * M + 64 // (2 * elem_bytes)
            out.append((xr, (xr // T) * gstr + (xr % T) * M))
        else:
            out.append((xr, xr * M))
    return out


def real_tile_io(M, T, X, elem_bytes, grouped):
    """worst multiplicity on the dense side (r2c input reads / c2r result writes) and on the (M+1)-bin side"""
    E = M // T
    rows = real_tile_rows(M, T, X, elem_bytes, grouped)
    words = 2 * elem_bytes // 4
    worst_dense = worst_odd = 1.0
    for e in range(E):
        for w0 in range(0, T * X, 32):
            dense, odd = [], []
            for lane in range(32):
                th = w0 + lane
                if th >= T * X:
                    dense.append(None); odd.append(None)
                    continue
                xi, tid = th // T, th % T
                dense.append((rows[xi][1] + tid + e * T) * words)
                odd.append((rows[xi][0] * (M + 1) + tid + e * T) * words)
            c, ideal = conflicts(dense, words)
            worst_dense = max(worst_dense, c / ideal)
            c, ideal = conflicts(odd, words)
            worst_odd = max(worst_odd, c / ideal)
    return worst_dense, worst_odd


# default kernels of csrc/wfb_variants_f32_pipe.cu: (N, T, rows per tile, rows per copy group for split / interleaved)
C2C_F32_TILES = {128: (8, 16, 2, 2), 256: (16, 8, 4, 0), 512: (16, 4, 2, 0), 1024: (32, 2, 0, 0), 2048: (128, 1, 0, 0), 4096: (256, 1, 0, 0)}
# k_real_pipe defaults whose thread groups share a phase: (M, T, X, bytes per real)
REAL_TILES_GROUPED = {"r2c/c2r f32 N=256": (128, 8, 16, 4), "r2c/c2r f64 N=128": (64, 4, 16, 8)}


if __name__ == "__main__":
    for name, plans, ew in (("f32", PLANS_F32, 2), ("f64", PLANS_F64, 4)):
        for padq in (0, 4, 8, 16, 32):
            row = []
            for n, (t, codes) in plans.items():
                w, r = simulate(n, t, codes, padq, ew)
                row.append(f"{n}:{w:.0f}/{r:.0f}")
            print(name, "padq", padq, " ".join(row))
