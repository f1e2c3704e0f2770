#!/usr/bin/env python3
"""Turns the raw captures a `tools/profile_round.sh` run leaves in gpurun_out/ into the tracked summaries under profiles/.

    python tools/make_profiles.py r01

Inputs (gpurun_out/):  launches.csv (ncu launch list of bench.py), traffic.csv (ncu DRAM bytes of one bench step),
full_*.raw.csv / full_*.source.csv.gz (pages of the ncu --set full captures), sweep_burst.jsonl / sweep_inv.jsonl / sweep_sustained.jsonl (tools/sweep.py).
Outputs (profiles/):   <round>_launches.{csv,md}, <round>_traffic.csv, traffic.json, <round>_ncu_full.md, <round>_sweep.md
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G = ROOT / "gpurun_out"
P = ROOT / "profiles"
RND = sys.argv[1] if len(sys.argv) > 1 else "r01"
SIZES = [16, 32, 64, 128, 256, 512, 1024, 2048, 4096]


def ncu_rows(path):
    """rows of an ncu --csv log (skips the ==PROF== banner lines)"""
    text = [l for l in open(path, errors="replace") if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("".join(text))))


def short(name):
    name = re.sub(r"\(int\)|\(bool\)|wfb::|void ", "", name)
    return name.replace("(KParams)", "").replace("(const KParams)", "").strip()


def launches():
    src = G / "launches.csv"
    if not src.exists():
        return
    (P / f"{RND}_launches.csv").write_text("".join(l for l in open(src, errors="replace") if l.startswith('"')))
    rows = ncu_rows(src)
    per = OrderedDict()
    other = [0.0, 0]
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        if per and "reduce_kernel" in r["Kernel Name"]:
            break          # bench.py's round-trip drift check: the device-resident phase is over (the staged e2e phase follows)
        v = float(r["Metric Value"].replace(",", ""))
        us = v / 1000.0 if r["Metric Unit"] in ("nsecond", "ns") else v
        k = short(r["Kernel Name"])
        if k.startswith("k_"):
            d = per.setdefault(k, [0.0, 0]); d[0] += us; d[1] += 1
        else:
            other[0] += us; other[1] += 1
    total = sum(v[0] for v in per.values())
    out = [f"# Round {RND[1:].lstrip('0')} — ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline`", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv`; listed: torch's input generation, the warm-up and the",
           "timed steps (the capture continues into the staged end-to-end phase, whose per-chunk launches are left out here).",
           "Per-launch times are cold-cache and serialised under the profiler, so compare SHARES, not absolutes.",
           f"Raw CSV: `profiles/{RND}_launches.csv`.  The same command ran to exit 0 without ncu first.", "",
           "| kernel (template arguments: Plan<N,T,passes…>, rows/CTA, …, IO 0=split, INV) | launches | total us | avg us | share |",
           "|---|---|---|---|---|"]
    for k, (us, n) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        out.append(f"| `{k}` | {n} | {us:.1f} | {us / n:.1f} | {100 * us / total:.1f}% |")
    out += ["", f"Other kernels in the window (torch RNG / elementwise building the synthetic input, outside the timed region): "
            f"{other[0]:.0f} us over {other[1]} launches.", "",
            "Every step is 18 launches of equal algorithmic bytes (2 GiB each), so shares are near-uniform; the ordering matches the live",
            "per-kernel CUDA-event times `bench.py` reports (`per_kernel`, `roofline.share_of_step`)."]
    (P / f"{RND}_launches.md").write_text("\n".join(out) + "\n")


def traffic():
    src = G / "traffic.csv"
    if not src.exists():
        return
    (P / f"{RND}_traffic.csv").write_text("".join(l for l in open(src, errors="replace") if l.startswith('"')))
    per_id = OrderedDict()
    for r in ncu_rows(src):
        if not short(r["Kernel Name"]).startswith("k_c2c"):
            continue
        per_id.setdefault(r["ID"], 0)
        per_id[r["ID"]] += int(float(r["Metric Value"].replace(",", "")))
    vals = list(per_id.values())
    # bench.py launches, per step, fwd then inv for each size in order; take the LAST full step in the capture
    vals = vals[-2 * len(SIZES):]
    if len(vals) != 2 * len(SIZES):
        print("traffic.csv: expected", 2 * len(SIZES), "c2c launches, got", len(vals)); return
    out = {}
    for i, n in enumerate(SIZES):
        out[f"k_c2c<f32,N={n},split,fwd>"] = vals[2 * i]
        out[f"k_c2c<f32,N={n},split,inv>"] = vals[2 * i + 1]
    (P / "traffic.json").write_text(json.dumps(dict(sorted(out.items())), indent=1) + "\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def full():
    reps = sorted(G.glob("full_*.raw.csv"))
    if not reps:
        return
    import gzip
    out = [f"# Round {RND[1:].lstrip('0')} — `ncu --set full` captures (1 GiB in / 1 GiB out per launch)", "",
           "Command per capture: `ncu --set full --clock-control none --import-source on -k regex:k_ -s 2 -c 1 python tools/prof_one.py <kind>:<N>`",
           "(third launch of the kernel; each command first ran to exit 0 without ncu).  The .ncu-rep files stay out of git; the numbers are",
           "from `ncu -i <rep> --page raw --csv` and the stall / instruction mix from `--page source --csv`.", "",
           "Algorithmic bytes per launch = 2 x 2^30 = 2.147 GB (c2c) ; DRAM traffic = `dram__bytes_read.sum + dram__bytes_write.sum`.", ""]
    for rep in reps:
        rows = list(csv.reader(open(rep, errors="replace")))
        if len(rows) < 3:
            continue
        h, units, r = rows[0], rows[1], rows[-1]
        ix = {n: i for i, n in enumerate(h)}
        out += [f"## `{short(r[ix['Kernel Name']])}`  ({rep.name[5:-8]})", "", "| metric | value |", "|---|---|"]
        for m in WANT:
            if m in ix:
                out.append(f"| {m} | {r[ix[m]]} {units[ix[m]]} |")

        def gb(m):
            v = float(r[ix[m]].replace(",", "")); u = units[ix[m]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        try:
            dram = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
            out.append(f"| **DRAM traffic** | {dram / 1e9:.3f} GB |")
        except Exception:
            pass
        sfile = Path(str(rep)[:-8] + ".source.csv.gz")
        srows = list(csv.reader(io.StringIO(gzip.open(sfile, "rt", errors="replace").read()))) if sfile.exists() else []
        hi = [i for i, x in enumerate(srows) if x and x[0] == "Address"]
        if hi:
            sh = srows[hi[0]]; sx = {n: i for i, n in enumerate(sh)}
            stalls = {n: 0.0 for n in sh if n.startswith("stall_") and "Not Issued" not in n}
            ops = {}
            for x in srows[hi[0] + 1:]:
                if len(x) < len(sh):
                    continue
                for n in stalls:
                    stalls[n] += float(x[sx[n]] or 0)
                s_ = x[sx["Source"]].strip()
                if not s_:
                    continue
                op = (s_.split()[1] if s_.startswith("@") else s_.split()[0]).split(".")[0]
                ops[op] = ops.get(op, 0) + float(x[sx["Instructions Executed"]] or 0)
            tot = sum(stalls.values()) or 1.0
            top = ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
            itot = sum(ops.values()) or 1.0
            mix = ", ".join(f"{k} {100 * v / itot:.0f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:8])
            out += [f"| warp-state samples | {top} |", f"| instruction mix (warp-level) | {mix} |"]
        out.append("")
    (P / f"{RND}_ncu_full.md").write_text("\n".join(out) + "\n")


def sweep():
    files = [("burst clocks (10 launches, median)", G / "sweep_burst.jsonl"), ("inverse c2c, burst clocks", G / "sweep_inv.jsonl"),
             ("power-capped steady state (1 s of back-to-back launches per variant)", G / "sweep_sustained.jsonl")]
    if not any(f.exists() for _, f in files):
        return
    peak = None
    try:
        peak = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        pass
    out = [f"# Round {RND[1:].lstrip('0')} — variant sweep (`tools/sweep.py`), 1 GiB of input per launch, device-resident", "",
           f"`frac` = algorithmic GB/s ÷ the measured HBM copy peak ({peak} GB/s, MEASURED_PEAKS.json).  The first variant listed for a",
           "(kind, N) is the plan's default; the others are the compiled alternates (`wfb_plan_set_variant`).", ""]
    for title, f in files:
        if not f.exists():
            continue
        out += [f"## {title}", "", "| kind | N | variant | ms | GB/s | frac | M transforms/s |", "|---|---|---|---|---|---|---|"]
        for l in open(f):
            if l.startswith("{"):
                r = json.loads(l)
                out.append(f"| {r['kind']} | {r['n']} | {r['variant']} | {r['ms']} | {r['GBs']} | {r['frac']} | {r['Mtransforms_s']} |")
        out.append("")
    (P / f"{RND}_sweep.md").write_text("\n".join(out) + "\n")


def pcie():
    """host-link ceiling (tools/pcie_diag.sh on an 8-GPU box): pcie_peak.jsonl + pcie_topology.txt (+ bench_n8.json)"""
    src = G / "pcie_peak.jsonl"
    if not src.exists():
        return
    rows = [json.loads(l) for l in open(src) if l.startswith("{")]
    out = [f"# Round {RND[1:].lstrip('0')} — concurrent pinned-copy ceiling of the host links (`tools/microbench/pcie_peak.cu`)", "",
           "One C process, one host thread per GPU, plain `cudaMemcpyAsync` of 256 MiB pinned buffers (6 copies per direction), all GPUs",
           "released by a pthread barrier, CUDA-event timed per GPU.  GB/s per direction: `min` = the slowest GPU, `sum` = all GPUs.",
           "This is the roofline of every end-to-end (host-buffer) number: each transform crosses the link once each way.", "",
           "| GPUs | pinned memory | thread affinity | H2D alone min / sum | D2H alone min / sum | duplex H2D min / sum | duplex D2H min / sum |",
           "|---|---|---|---|---|---|---|"]
    for r in rows:
        out.append(f"| {r['gpus']} | {r['alloc']} | {r['affinity']} | {r['h2d_min']} / {r['h2d_sum']} | {r['d2h_min']} / {r['d2h_sum']} | "
                   f"{r['duplex_h2d_min']} / {r['duplex_h2d_sum']} | {r['duplex_d2h_min']} / {r['duplex_d2h_sum']} |")
    best = {}
    for r in rows:
        b = best.setdefault(r["gpus"], r)
        if r["duplex_h2d_sum"] + r["duplex_d2h_sum"] > b["duplex_h2d_sum"] + b["duplex_d2h_sum"]:
            best[r["gpus"]] = r
    one = best.get(1)
    out += ["", "## What it says", ""]
    if one:
        out.append(f"* One GPU: {one['h2d_sum']} / {one['d2h_sum']} GB/s alone, {one['duplex_h2d_sum']} / {one['duplex_d2h_sum']} GB/s with both directions running "
                   "(PCIe 5.0 x16).")
        for g in sorted(best):
            if g == 1:
                continue
            b = best[g]
            eff = min(b["duplex_h2d_sum"], b["duplex_d2h_sum"]) / (g * min(one["duplex_h2d_sum"], one["duplex_d2h_sum"]))
            out.append(f"* {g} GPUs: duplex {b['duplex_h2d_sum']} / {b['duplex_d2h_sum']} GB/s in total = **{eff:.2f}** of {g} x the one-GPU rate; the slowest GPU gets "
                       f"{b['duplex_h2d_min']} / {b['duplex_d2h_min']} GB/s, so a max-over-ranks time cannot beat {g} x {min(b['duplex_h2d_min'], b['duplex_d2h_min'])} = "
                       f"{g * min(b['duplex_h2d_min'], b['duplex_d2h_min']):.1f} GB/s per direction.")
        out += ["* How the host memory is pinned (cudaHostAlloc, write-combined, transparent-hugepage + cudaHostRegister) and where the issuing threads",
                "  run makes no difference: the box is a KVM guest with ONE NUMA node (`pcie_topology.txt` below: every GPU reports numa -1, CPU affinity",
                "  0-31), and its host fabric -- not the per-GPU links, not this library's staging -- stops scaling: two and four GPUs share what one",
                "  GPU gets in duplex, eight GPUs reach ~1.6 x.  The end-to-end scaling of ANY host-buffer workload on this pool is bounded by this table."]
    b8 = G / "bench_n8.json"
    if b8.exists():
        try:
            d = json.load(open(b8)); e = d["e2e"]
            g8 = best.get(8)
            out += ["", f"## `bench.py --gpus 8` end-to-end leg on the same box", "",
                    f"`createFFTf32Split(n, batch).forward()/inverse()` on pinned host buffers, 8 ranks: {e['GBs_per_direction_per_gpu']} GB/s per direction per GPU"
                    f" = {e['GBs_per_direction_per_gpu'] * 8:.1f} GB/s in total ({e['value'] / 1e6:.1f} M transforms/s)."]
            if g8:
                out.append(f"Against the table: {e['GBs_per_direction_per_gpu'] * 8 / min(g8['duplex_h2d_sum'], g8['duplex_d2h_sum']):.2f} of the duplex sum, "
                           f"{e['GBs_per_direction_per_gpu'] / min(g8['duplex_h2d_min'], g8['duplex_d2h_min']):.2f} of the slowest-GPU rate (the timing is a max over ranks).")
        except Exception as ex:
            out.append(f"(bench_n8.json unreadable: {ex!r})")
    pairs = [json.loads(l) for f in ("pcie_pairs.jsonl", "pcie_spread.jsonl") if (G / f).exists() for l in open(G / f) if l.startswith("{")]
    pairs = [r for r in pairs if r["affinity"] == "none"]
    if pairs:
        out += ["", "## Which GPUs share a host bridge (`pcie_peak 256 6 pairs` / `spread`, tools/pcie_diag2.sh)", "",
                "| GPUs | H2D alone sum | D2H alone sum | duplex H2D sum | duplex D2H sum |", "|---|---|---|---|---|"]
        seen = set()
        for r in pairs:
            if r["devices"] in seen:
                continue
            seen.add(r["devices"])
            out.append(f"| {r['devices']} | {r['h2d_sum']} | {r['d2h_sum']} | {r['duplex_h2d_sum']} | {r['duplex_d2h_sum']} |")
        out += ["", "GPU 0 paired with GPU 1, 2 or 3 gets the one-GPU duplex rate in total; paired with GPU 4..7 it gets 1.5-1.6 x: the board's",
                "GPUs hang off two host bridges (0-3, 4-7), and each bridge's path to host memory carries about what one GPU can use.  Hence",
                "`bench.py` places rank i on GPU i * visible/world when the box shows more GPUs than ranks (WFB_BENCH_DEVICE_ORDER=seq",
                "restores rank i -> GPU i).  Four GPUs spread over both bridges reach 83 / 92 GB/s, eight 77 / 82: the box as a whole tops out there."]
        for n in (2, 4):
            f = G / f"bench_n{n}_spread.json"
            if f.exists():
                try:
                    d = json.load(open(f)); e = d["e2e"]
                    out.append(f"* `bench.py --gpus {n}` ({d['devices']['order']}): e2e {e['value'] / 1e6:.1f} M transforms/s, {e['GBs_per_direction_per_gpu']} GB/s per direction per GPU; "
                               f"{e['frac']} of the duplex sum measured in that run, {e['frac_of_slowest_rank_ceiling']} of n x the slowest rank.")
                except Exception as ex:
                    out.append(f"* bench_n{n}_spread.json unreadable: {ex!r}")
    topo = G / "pcie_topology.txt"
    if topo.exists():
        out += ["", "## Box topology (`tools/pcie_diag.sh`)", "", "```", topo.read_text().strip(), "```"]
    (P / f"{RND}_pcie.md").write_text("\n".join(out) + "\n")


if __name__ == "__main__":
    P.mkdir(exist_ok=True)
    launches(); traffic(); full(); sweep(); pcie()
    print("profiles written:", ", ".join(sorted(p.name for p in P.iterdir())))
