#!/usr/bin/env python
"""Turn the raw artefacts under gpurun_out/ (ncu reports, sweep / bench JSON) into the tracked summaries
under profiles/: ncu key metrics per kernel (CSV), roofline fractions against MEASURED_PEAKS.json, the
allreduce table and the bench table.  Runs on the CPU box (ncu -i only reads reports)."""
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


def main():
    os.makedirs(OUT, exist_ok=True)
    hbm, tf_burst, tf_sus, how = peaks()
    lines = ["# Profiles\n",
             "Roofline denominators (%s): HBM copy %.1f GB/s, bf16 GEMM %.1f TFLOP/s burst / %.1f sustained "
             "(`MEASURED_PEAKS.json`); NVLink 770 GB/s measured peer copy / 900 nominal per direction.\n" % (how, hbm, tf_burst, tf_sus),
             "All `ncu` captures: `ncu --set full --clock-control none --import-source on` on one B200 "
             "(`benchmarks/gpu_call19.sh`, drivers in `benchmarks/ncu_targets.py`); numbers under the profiler are for "
             "shares and counters only - the timing numbers in the tables further down are CUDA-event timed.\n",
             "## ncu key metrics\n",
             "| report | kernel | time (us) | DRAM rd+wr (MB) | achieved GB/s | %% of %s HBM | tensor pipe %% | regs | grid x block |"
             % how, "|---|---|---|---|---|---|---|---|---|"]
    flops = {"prof_gemm": 2.0 * 8192 * 4096 * 4096, "prof_gemm_bert": 2.0 * 8192 * 4096 * 1024}
    # prof_allreduce (single GPU, world=1) from an earlier capture is kept only if present
    for rep in sorted(glob.glob(os.path.join(GO, "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[:-8]
        rows, units = ncu_rows(rep)
        if not rows:
            continue
        with open(os.path.join(OUT, "ncu_%s.csv" % name), "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
            w.writeheader()
            w.writerow({k: units.get(k, "") for k in rows[0]})
            for r in rows:
                w.writerow(r)
        for r in rows:
            t = to_us(r["gpu__time_duration.sum"], units["gpu__time_duration.sum"])
            mb = to_mb(r["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) + \
                to_mb(r["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
            gbs = mb / 1e3 / (t * 1e-6)
            k = r["Kernel Name"].split("(")[0].replace("void ", "")
            extra = ""
            if name in flops and "gemm" in k:
                tfl = flops[name] / (t * 1e-6) / 1e12
                extra = " **%.0f TFLOP/s = %.0f%% of %s burst**" % (tfl, 100 * tfl / tf_burst, how)
            lines.append("| %s | `%s`%s | %.1f | %.1f | %.0f | %.0f%% | %s | %s | %s x %s |" % (
                name, k, extra, t, mb, gbs, 100 * gbs / hbm,
                r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "-")[:5],
                r["launch__registers_per_thread"], r["launch__grid_size"], r["launch__block_size"]))
    # allreduce sweep
    for np_ in (8, 2):
        p = os.path.join(OUT, "allreduce_sweep_np%d.json" % np_)
        if not os.path.exists(p):
            continue
        d = json.load(open(p))
        lines += ["\n## Allreduce, np=%d (device-timed, max over ranks; transport %s, multicast %s)\n" % (
            np_, d["info"]["transport"], d["info"]["multicast"]),
            "busbw = 2(N-1)/N x bytes / t.  `nvls` = multimem.ld_reduce/st kernel, `p2p` = two-shot peer ld/st kernel, "
            "NCCL = torch.distributed (2.28.9), the NCCL-via-torch stand-in for horovod+NCCL.\n",
            "| dtype | bytes | one-shot us | p2p us | nvls us | NCCL us | p2p busbw | nvls busbw | NCCL busbw | best/NCCL | nvls vs 900 GB/s |",
            "|---|---|---|---|---|---|---|---|---|---|---|"]
        for r in d["rows"]:
            ours = min(x for x in (r.get("oneshot_us"), r.get("twoshot_p2p_us"), r.get("nvls_us")) if x)
            nv = r.get("nvls_busbw")
            lines.append("| %s | %d | %s | %s | %s | %s | %s | %s | %s | %.2fx | %s |" % (
                r["dtype"], r["bytes"], r.get("oneshot_us", "-"), r.get("twoshot_p2p_us", "-"), r.get("nvls_us", "-"),
                r.get("nccl_us", "-"), r.get("twoshot_p2p_busbw", "-"), nv if nv else "-", r.get("nccl_busbw", "-"),
                (r["nccl_us"] / ours) if r.get("nccl_us") else float("nan"), ("%.0f%%" % (100 * nv / 900)) if nv else "-"))
    # bench table
    lines += ["\n## End-to-end benches (bench.py, one JSON line each; `profiles/bench/`)\n",
              "| run | impl | N | value | ms/step | e2e value | our kernel launches |", "|---|---|---|---|---|---|---|"]
    for p in sorted(glob.glob(os.path.join(OUT, "bench", "*.json"))):
        try:
            d = json.load(open(p))
        except ValueError:
            continue
        lines.append("| %s | %s | %s | %.1f %s | %s | %.1f | %s |" % (
            os.path.basename(p)[:-5], d.get("impl"), d.get("n_gpus"), d["value"], d["unit"], d["ms_per_step"],
            d["e2e"]["value"], d.get("gpu_launches")))
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
