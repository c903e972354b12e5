#!/usr/bin/env python
"""Turn the raw artefacts (ncu reports under gpurun_out/, JSON / text results under profiles/) into the tracked summary
profiles/README.md: ncu key metrics per kernel with roofline fractions against MEASURED_PEAKS.json, the GEMM table, the
collective tables (np = 2 / 4 / 8), NVLink link counters, end-to-end benches, GBDT, sanitizers.  Runs on the CPU box
(``ncu -i`` only reads reports).  Layout of profiles/: see the "Files" section the script writes."""
import csv
import glob
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")

METRICS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "lts__t_bytes.sum", "smsp__cycles_active.avg"]

# FLOPs of the launches captured by `ncu_targets.py gemm`, in capture order (-s 3 -c 5)
GEMM_ROWS = [("fwd + bias + GELU 8192x4096x4096", 2.0 * 8192 * 4096 * 4096), ("dgrad 8192x4096x4096", 2.0 * 8192 * 4096 * 4096),
             ("wgrad 8192x4096x4096 (split-K)", 2.0 * 8192 * 4096 * 4096), ("fwd + bias + GELU 4096x4096x1024 (BERT FFN-up)", 2.0 * 4096 * 4096 * 1024),
             ("fwd 4096x30528x1024 (BERT vocabulary projection)", 2.0 * 4096 * 30528 * 1024)]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def ncu_rows(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        return [], {}
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(m) for m in METRICS if m in hdr]
    out = [{hdr[i]: row[i] for i in idx} for row in rows[2:]]
    return out, {hdr[i]: units[i] for i in idx}


def to_us(v, unit):
    v = float(v)
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)


def to_mb(v, unit):
    v = float(v)
    return {"byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(unit, v)


def load(path):
    try:
        return json.load(open(path))
    except (OSError, ValueError):
        return None


def ncu_section(lines, hbm, tf_burst, how):
    lines += ["## ncu key metrics (round 2 kernels)\n",
              "`ncu --set full --clock-control none --import-source on`, one B200, `benchmarks/gpu_batch.sh ncuall` "
              "(drivers in `benchmarks/ncu_targets.py`).  Numbers under the profiler (cold caches, serialised replays) are "
              "for shares and counters only; timing claims elsewhere in this file are CUDA-event timed.  CSV per report in "
              "`profiles/ncu/`.\n",
              "| report | kernel | time (us) | DRAM rd+wr (MB) | achieved GB/s | %% of %s HBM copy | tensor pipe %% | regs | grid x block |" % how,
              "|---|---|---|---|---|---|---|---|---|"]
    os.makedirs(os.path.join(OUT, "ncu"), exist_ok=True)
    for rep in sorted(glob.glob(os.path.join(GO, "r2_*.ncu-rep"))):
        name = os.path.basename(rep)[:-8]
        rows, units = ncu_rows(rep)
        if not rows:
            continue
        with open(os.path.join(OUT, "ncu", "%s.csv" % name), "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
            w.writeheader()
            w.writerow({k: units.get(k, "") for k in rows[0]})
            for r in rows:
                w.writerow(r)
        for i, r in enumerate(rows):
            t = to_us(r["gpu__time_duration.sum"], units["gpu__time_duration.sum"])
            mb = to_mb(r["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) + \
                to_mb(r["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
            gbs = mb / 1e3 / (t * 1e-6)
            k = r["Kernel Name"].split("(")[0].replace("void ", "")
            extra = ""
            if name == "r2_gemm" and i < len(GEMM_ROWS):
                tfl = GEMM_ROWS[i][1] / (t * 1e-6) / 1e12
                extra = " %s: **%.0f TFLOP/s = %.0f%% of %s cuBLAS burst**" % (GEMM_ROWS[i][0], tfl, 100 * tfl / tf_burst, how)
            lines.append("| %s | `%s`%s | %.1f | %.1f | %.0f | %.0f%% | %s | %s | %s x %s |" % (
                name, k, extra, t, mb, gbs, 100 * gbs / hbm,
                r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "-")[:5],
                r["launch__registers_per_thread"], r["launch__grid_size"], r["launch__block_size"]))
    lines.append("")


def gemm_section(lines):
    d = load(os.path.join(OUT, "gemm", "gemm_bench.json"))
    if not d:
        return
    lines += ["## tcgen05 GEMM vs cuBLAS (`benchmarks/gemm_bench.py`)\n",
              "CUDA-graph timed (no host launch cost), ring of operand sets larger than L2, better of two interleaved "
              "measurements per arm; `lib` = `torch` (cuBLAS / cuBLASLt) on the same box, same shapes, same dtype; "
              "`dgrad*dgelu` has no single library call (cuBLAS dgrad + an elementwise pass).\n",
              "| M | N | K | op | ours us | lib us | ours TFLOP/s | lib TFLOP/s | ours / lib |", "|---|---|---|---|---|---|---|---|---|"]
    for r in d["rows"] if isinstance(d, dict) else d:
        lines.append("| %d | %d | %d | %s | %s | %s | %s | %s | %s |" % (
            r["M"], r["N"], r["K"], r["op"], r["ours_us"], r["lib_us"], r["ours_tflops"], r["lib_tflops"], r["ours_over_lib"]))
    lines.append("")


def allreduce_section(lines):
    for np_ in (8, 4, 2):
        d = load(os.path.join(OUT, "sweeps", "allreduce_sweep_np%d.json" % np_))
        if not d:
            continue
        small = load(os.path.join(OUT, "sweeps", "allreduce_sweep_np%d_small_ll.json" % np_))
        ll = {(r["dtype"], r["bytes"]): r.get("ll_us") for r in small["rows"]} if small else {}
        lines += ["## Allreduce on symmetric buffers, np=%d (device-timed, max over ranks; transport %s, multicast %s)\n" % (
            np_, d["info"]["transport"], d["info"]["multicast"]),
            "busbw = 2(N-1)/N x bytes / t.  `LL` = barrier-free flag-in-data kernel (<= 64 KB), `one-shot` / `p2p` = peer "
            "ld/st kernels, `nvls` = multimem.ld_reduce / multimem.st kernel, `wire` = 16-bit / block-scaled fp8 wire "
            "formats of an fp32 tensor, NCCL = torch.distributed (2.28.9) on the same GPUs."
            + ("  The `nvls` column of this np=8 run used 32 CTAs; the sweep's CTA variants (JSON keys `nvls_ctas16_busbw` ...) "
               "showed 16 CTAs to be faster up to 64 MB (32 MB: 90.5 us instead of 98.7), which is now the default.\n" if np_ == 8 else "\n"),
            "| dtype | bytes | LL us | one-shot us | p2p us | nvls us | wire bf16 us | wire fp8 us | NCCL us | best busbw GB/s | NCCL busbw | best / NCCL |",
            "|---|---|---|---|---|---|---|---|---|---|---|---|"]
        for r in d["rows"]:
            llv = r.get("ll_us") or ll.get((r["dtype"], r["bytes"]))
            cands = [x for x in (llv, r.get("oneshot_us"), r.get("twoshot_p2p_us"), r.get("nvls_us")) if x]
            ours = min(cands)
            bus = 2 * (np_ - 1) / np_ * r["bytes"] / (ours * 1e-6) / 1e9
            lines.append("| %s | %d | %s | %s | %s | %s | %s | %s | %s | %.1f | %s | %s |" % (
                r["dtype"], r["bytes"], llv or "-", r.get("oneshot_us", "-"), r.get("twoshot_p2p_us", "-"), r.get("nvls_us", "-"),
                r.get("wire_bf16_us", "-"), r.get("wire_fp8_us", "-"), r.get("nccl_us", "-"), bus, r.get("nccl_busbw", "-"),
                ("%.2fx" % (r["nccl_us"] / ours)) if r.get("nccl_us") else "-"))
        lines.append("")


def collective_section(lines):
    for np_ in (8, 4, 2):
        d = load(os.path.join(OUT, "sweeps", "collective_sweep_np%d.json" % np_))
        if not d:
            continue
        lines += ["## allgather / broadcast / reducescatter through the `hvd` API, np=%d\n" % np_,
                  "Public API (`hvd.allgather_async` ... -> C++ negotiation engine -> peer-memory kernels) against "
                  "`torch.distributed` NCCL collectives on preallocated tensors; the engine adds a negotiation round trip "
                  "and a staging copy per call, so small messages are slower than NCCL here (the symmetric-buffer kernels "
                  "above are the fast path the optimizer uses).\n",
                  ("Note: this np=%d table was taken before `hvd.allgather` stopped doing a blocking host-side size exchange per "
                   "call (one negotiation now, see the np=2 table, taken after) and before the benchmark warmed the "
                   "allocator with the same number of operations in flight; no 8-GPU budget was left to repeat it.\n" % np_)
                  if np_ in (4, 8) else "",
                  "| bytes | allgather ours us | NCCL us | broadcast ours us | NCCL us | reducescatter ours us | NCCL us |", "|---|---|---|---|---|---|---|"]
        for r in d["rows"]:
            lines.append("| %d | %s | %s | %s | %s | %s | %s |" % (
                r["bytes"], r.get("allgather_ours_us"), r.get("allgather_nccl_us"), r.get("broadcast_ours_us"),
                r.get("broadcast_nccl_us"), r.get("reducescatter_ours_us"), r.get("reducescatter_nccl_us")))
        lines.append("")


def nvlink_section(lines):
    any_ = False
    for np_ in (8, 2):
        d = load(os.path.join(OUT, "nvlink", "nvlink_np%d.json" % np_))
        if not d:
            continue
        if not any_:
            lines += ["## NVLink byte counters around the fused compute+collective kernels (`benchmarks/nvlink_counters.py`)\n",
                      "ncu cannot replay a kernel that spins on flags written by a peer process (and the profiling recipe "
                      "forbids wrapping a multi-rank command), so the link evidence is the driver's own counters: NVML "
                      "`NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX/RX` deltas over 50 back-to-back launches, device-timed, max "
                      "over ranks.  `expected` = algorithmic bytes per direction; link rate = max(tx, rx) / time against "
                      "the 770 GB/s measured peer copy / 900 GB/s nominal per direction.\n"]
            any_ = True
        lines += ["np=%d (%s):\n" % (np_, d["info"]["counter"]),
                  "| kernel | us | TX bytes / launch | RX bytes / launch | expected / direction | link GB/s | of 770 | of 900 |",
                  "|---|---|---|---|---|---|---|---|"]
        for r in d["rows"]:
            lines.append("| %s | %s | %d | %d | %d | %s | %.0f%% | %.0f%% |" % (
                r["kernel"], r["us"], r["nvlink_tx_bytes_per_launch"], r["nvlink_rx_bytes_per_launch"],
                r["expected_bytes_per_direction"], r["link_GBps_per_direction"], 100 * r["frac_of_770"], 100 * r["frac_of_900"]))
        lines.append("")


def bench_section(lines):
    lines += ["## End-to-end benches (`bench.py`, one JSON line each; `profiles/bench/`)\n",
              "`ours` = this framework through `sparkdl.HorovodRunner` + `hvd.DistributedOptimizer(fused=True)`; `reference` = "
              "the unmodified reference package (`baseline/_ref`) through its own `HorovodRunner`, stock torch model code, "
              "gradient exchange by torch DDP over NCCL.  Same model, batch, optimizer, dtype, synthetic data.\n",
              "| run | impl | N | value | ms/step | e2e value | e2e ms/step | our kernel launches | clocks (MHz, reasons) |",
              "|---|---|---|---|---|---|---|---|---|"]
    for p in sorted(glob.glob(os.path.join(OUT, "bench", "*.json"))):
        d = load(p)
        if not d or "value" not in d:
            continue
        ck = d.get("clocks", {})
        lines.append("| %s | %s | %s | %.1f %s | %s | %.1f | %s | %s | %s %s |" % (
            os.path.basename(p)[:-5], d.get("impl"), d.get("n_gpus"), d["value"], d["unit"], d["ms_per_step"],
            d["e2e"]["value"], d["e2e"].get("ms_per_step", "-"), d.get("gpu_launches"), ck.get("sm_mhz"), ck.get("reasons")))
    lines.append("")


def xgb_section(lines):
    rows = []
    for np_ in (1, 2, 8):
        for tag, name in (("", "shared-memory histograms"), ("_globalatomics", "global atomics (round-1 kernel)")):
            d = load(os.path.join(OUT, "xgboost", "xgboost_bench_np%d%s.json" % (np_, tag)))
            if d:
                rows.append((np_, name, d))
    if not rows:
        return
    lines += ["## GBDT histogram build + allreduce (K8), 1 M rows x 64 features per worker, depth 6 (`benchmarks/xgboost_bench.py`)\n",
              "`fused` = one kernel: shared-memory histogram build -> grid barrier -> two-shot int64 allreduce over peer "
              "memory; `unfused` = histogram kernel + the generic int64 allreduce kernel; `unfused + NCCL` = histogram "
              "kernel + `torch.distributed.all_reduce` of the same bytes.  Level of 32 nodes (8 MB of histograms).  "
              "`sec/round` = median gap between per-round callbacks (sketch and binning excluded).\n",
              "| np | histogram kernel | fused us / level | unfused us | unfused + NCCL us | sec / round | rows / s | sklearn HGB sec / round (CPU, 1 shard) |",
              "|---|---|---|---|---|---|---|---|"]
    for np_, name, d in rows:
        sk = d.get("sklearn_hgb_sec_per_round_1shard_cpu")
        lines.append("| %d | %s | %.1f | %.1f | %s | %s | %s | %s |" % (
            np_, name, d["hist_level_fused_us"], d["hist_level_unfused_us"],
            ("%.1f" % d["hist_level_unfused_plus_nccl_us"]) if isinstance(d.get("hist_level_unfused_plus_nccl_us"), float) else "-",
            ("%.4f" % d["sec_per_round"]) if d["sec_per_round"] > 1e-6 else "(timing fixed later)",
            ("%.3g" % d["rows_per_sec"]) if d["sec_per_round"] > 1e-6 else "-", ("%.3f" % sk) if isinstance(sk, float) else "-"))
    lines.append("")


def sanitizer_section(lines):
    files = sorted(glob.glob(os.path.join(OUT, "sanitizer", "*.txt")))
    if not files:
        return
    lines += ["## Sanitizers (`benchmarks/sanitize.sh`; `profiles/sanitizer/`)\n", "| log | result |", "|---|---|"]
    for p in files:
        txt = open(p, errors="replace").read().strip().splitlines()
        tail = [t for t in txt if "SUMMARY" in t or "passed" in t or "failed" in t or "ThreadSanitizer" in t or "GEMM_OK" in t]
        lines.append("| %s | %s |" % (os.path.basename(p), "; ".join(t.strip("= ").strip() for t in tail[-3:]) or "(see file)"))
    lines.append("")


def main():
    os.makedirs(OUT, exist_ok=True)
    hbm, tf_burst, tf_sus, how = peaks()
    lines = ["# Profiles (round 2)\n",
             "Roofline denominators (%s): HBM copy %.1f GB/s, bf16 cuBLAS GEMM %.1f TFLOP/s burst / %.1f sustained "
             "(`MEASURED_PEAKS.json`); NVLink 770 GB/s measured peer copy / 900 nominal per direction.  Everything in this "
             "directory was produced on B200 boxes through `gpurun` by `benchmarks/gpu_batch.sh <stage>` and summarised here by "
             "`benchmarks/make_profiles.py` / `benchmarks/make_sass.py`.\n" % (how, hbm, tf_burst, tf_sus),
             "## Files\n",
             "| path | what |", "|---|---|",
             "| `ncu/` | key metrics per captured kernel (CSV), from the `--set full` reports |",
             "| `sass/` | `SUMMARY.md` + key-instruction listings (`UTCHMMA.2CTA`, `UTMALDG.2D.2CTA`, `LDGMC`, peer `LDG/STG...SYS`) |",
             "| `gemm/` | `gemm_bench.json` (ours vs cuBLAS), `gemm_lab2.txt` (epilogue experiments) |",
             "| `sweeps/` | allreduce and allgather/broadcast/reducescatter sweeps at np = 2, 4, 8 vs NCCL |",
             "| `nvlink/` | NVML link byte counters around the collective kernels at np = 2, 8 |",
             "| `bench/` | `bench.py` JSON lines, both arms, N = 1, 2, 8 |",
             "| `kernels/` | per-kernel time tables of one training step (torch profiler) and small lab results (BN variants, stem, forked outputs) |",
             "| `xgboost/` | GBDT bench JSON (np = 1, 2, 8; shared-memory vs global-atomic histograms), training profile |",
             "| `tests/` | logs of the multi-GPU pytest module at np = 2, 4, 8, collective fuzzers, the 1-GPU suite |",
             "| `sanitizer/` | compute-sanitizer memcheck / racecheck / synccheck and the engine's ThreadSanitizer run |",
             "| `round1/` | the previous round's artefacts, kept for comparison |", ""]
    ncu_section(lines, hbm, tf_burst, how)
    gemm_section(lines)
    bench_section(lines)
    allreduce_section(lines)
    nvlink_section(lines)
    collective_section(lines)
    xgb_section(lines)
    sanitizer_section(lines)
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:30]))


if __name__ == "__main__":
    main()
