#!/usr/bin/env python
"""Turns the raw ncu output of tests/gpu/profile_step.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/:

  <tag>_launches.csv            every launch of ONE step: id, kernel, grid, block, microseconds
  <tag>_launch_shares.txt       per-kernel totals and shares of that step
  <tag>_kernel_counters.csv     per launch: time, DRAM bytes read / written, achieved GB/s, instructions (glue + GEMM kernels)
  <tag>_hbm_table.txt           the row-stream / elementwise kernels against the measured HBM peak
  <tag>_ncu_conv_pair_summary.txt, r02_ncu_conv_res.json   the `--set full` capture of the residual-block convolution

usage: python tests/gpu/summarize_profiles.py <tag>          (needs `ncu` for the .ncu-rep import; no GPU)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02c"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM = peaks["hbm_gbs"]


def short(name):
    name = re.sub(r"^(void )?(sggan::)?", "", name)            # leading qualifiers only (the parameter list names sggan:: too)
    name = re.sub(r"\((int|bool)\)", "", name)                 # row_stream_kernel<(int)2, (int)2> -> row_stream_kernel<2, 2>
    m = re.match(r"[A-Za-z_0-9]+(<[^>]*>)?", name)
    return m.group(0) if m else name


def rows_of(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


def us(v, unit):
    v = float(v.replace(",", ""))
    return v / 1000.0 if unit == "ns" else (v * 1000.0 if unit == "ms" else v)


def nbytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


# ---- 1. launch list of one step
rs = [r for r in rows_of(os.path.join(G, tag + "_launches_all.csv")) if r["Metric Name"] == "gpu__time_duration.sum"]
launches = [(int(r["ID"]), short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], us(r["Metric Value"], r["Metric Unit"])) for r in rs]
marks = [i for i, x in enumerate(launches) if x[1] == "bump_step_kernel"]
step = launches[marks[2] + 1:marks[3] + 1]  # the 4th step (three warm-up steps before it)
with open(os.path.join(P, tag + "_launches.csv"), "w") as f:
    f.write("id,kernel,grid,block,us\n")
    for i, k, g, b, t in step:
        f.write('%d,%s,"%s","%s",%.2f\n' % (i, k, g, b, t))
agg = collections.OrderedDict()
for _, k, _, _, t in step:
    d = agg.setdefault(k, [0, 0.0])
    d[0] += 1
    d[1] += t
tot = sum(d[1] for d in agg.values())
with open(os.path.join(P, tag + "_launch_shares.txt"), "w") as f:
    f.write("# one training step (256x512, batch 8, C=34, loss p2p), eager launches under\n"
            "# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES, not absolutes)\n"
            "# %d launches, %.1f us serialised\n" % (len(step), tot))
    for k, d in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-34s %4d launches %9.1f us %5.1f %%  %7.1f us/launch\n" % (k, d[0], d[1], 100 * d[1] / tot, d[1] / d[0]))

# ---- 2. per-launch counters
by_id = collections.OrderedDict()
for r in rows_of(os.path.join(G, tag + "_glue_counters.csv")):
    e = by_id.setdefault(r["ID"], {"kernel": short(r["Kernel Name"]), "full": r["Kernel Name"], "grid": r["Grid Size"]})
    e[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
with open(os.path.join(P, tag + "_kernel_counters.csv"), "w") as f:
    f.write("id,kernel,grid,us,dram_read_MB,dram_write_MB,dram_GBps,frac_of_hbm_peak,inst_executed,sm_throughput_pct,dram_throughput_pct\n")
    table = collections.OrderedDict()
    for i, e in by_id.items():
        t = us(*e["gpu__time_duration.sum"])
        rd, wr = nbytes(*e["dram__bytes_read.sum"]), nbytes(*e["dram__bytes_write.sum"])
        gbs = (rd + wr) / (t * 1e-6) / 1e9
        name = e["kernel"]
        f.write('%s,%s,"%s",%.2f,%.2f,%.2f,%.0f,%.3f,%s,%s,%s\n' % (
            i, name, e["grid"], t, rd / 1e6, wr / 1e6, gbs, gbs / HBM, e["smsp__inst_executed.sum"][0].replace(",", ""),
            e["sm__throughput.avg.pct_of_peak_sustained_elapsed"][0], e["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0]))
        if (rd + wr) >= 20e6 and ("row_stream" in name or name in ("act_bwd_kernel", "fake_grad_kernel")):
            d = table.setdefault(name, [0, 0.0, 0.0])
            d[0] += 1
            d[1] += t
            d[2] += rd + wr
with open(os.path.join(P, tag + "_hbm_table.txt"), "w") as f:
    f.write("# HBM-bound passes of one step, launches that move >= 20 MB (ncu: gpu__time_duration, dram__bytes_read + write;\n"
            "# cache control = flush before every launch and six metrics collected per launch, so these are cold-cache, replayed\n"
            "# numbers -- the bench line's roofline_hbm objects are the CUDA-event timings inside the step); row_stream<mode, streams>:\n"
            "# mode 0 apply, 1 backward reduce, 2 backward apply; peak = %.1f GB/s (MEASURED_PEAKS.json)\n" % HBM)
    f.write("%-22s %8s %12s %14s %10s %8s\n" % ("kernel", "launches", "us/launch", "MB/launch", "GB/s", "of peak"))
    for k, d in table.items():
        f.write("%-22s %8d %12.1f %14.1f %10.0f %8.2f\n" % (k, d[0], d[1] / d[0], d[2] / d[0] / 1e6, d[2] / d[1] / 1e3, d[2] / d[1] / 1e3 / HBM))

# ---- 3. the full capture of the residual-block convolution
rep = os.path.join(G, tag + "_conv_res.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
keys = ["Kernel Name", "Grid Size", "Block Size", "launch__cluster_dim_x", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum"]
traffic = []
with open(os.path.join(P, tag + "_ncu_conv_pair_summary.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:conv_gemm_pair_kernel -s 4 -c 2 (tests/gpu/profile_step.sh)\n"
            "# residual-block 3x3 256->256 convolution at 64x128, batch 8, inside a training step (eager launches); read here with\n"
            "# ncu -i gpurun_out/%s_conv_res.ncu-rep --page raw --csv\n" % tag)
    for r in rr[2:]:
        vals = {}
        for k in keys:
            for i, h in enumerate(hdr):
                if h == k:
                    f.write("%-72s %s %s\n" % (k, r[i][:70], units[i]))
                    vals[k] = (r[i], units[i])
        traffic.append(nbytes(*vals["dram__bytes_read.sum"]) + nbytes(*vals["dram__bytes_write.sum"]))
        f.write("\n")
json.dump({"kernel": "conv_gemm_pair_kernel, residual-block 3x3 256->256, 64x128, batch 8", "dram_bytes_per_launch": sum(traffic) / len(traffic),
           "source": "profiles/%s_ncu_conv_pair_summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, cold cache)" % tag},
          open(os.path.join(P, "r02_ncu_conv_res.json"), "w"), indent=1)
print("wrote summaries for", tag)
