#!/usr/bin/env python3
"""Regenerates the tables of profiles/README.md from the raw ncu outputs.

    python profiles/summarize.py profiles/r2_launches.csv gpurun_out/prof_r2.ncu-rep [r2]

(1) launch list -> per-kernel share; (2) `ncu --set full` report -> one line per captured kernel
(profiles/r1_ncu_full.txt) and profiles/traffic.json (DRAM read+write bytes per launch, used by bench.py's
roofline.traffic)."""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def launch_shares(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        name = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for n, a in agg.items() if "synth" not in n)
    print("| kernel | launches | total µs | share |\n|---|---|---|---|")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if "synth" not in n:
            print(f"| `{n}` | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / tot:.1f} % |")


KEEP = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.sum", "sm__inst_executed_pipe_uniform.sum"]


def key_of(name):
    for pat, key in (("tc_blur", "tc_blur"), ("adaptive_fix", "adaptive_gauss_fix"), ("adaptive_tail", "adaptive_gauss_tail"), ("adaptive", "adaptive_gauss"), ("lut16", "pw_lut"), ("otsu", "scalars_otsu"), ("warp_persp", "warp_perspective_c3"), ("blur", "blur_gauss"),
                     ("morph_march", "morph_march"), ("mask_blend", "mask_blend"), ("warp_affine", "warp_affine")):
        if pat in name:
            return key
    return name


def full_report(rep, tag="r2"):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = [k for k in KEEP if k in hdr]                 # metric names differ a little between ncu versions
    idx = [hdr.index(k) for k in keep]
    traffic = {}
    with open(os.path.join(HERE, f"{tag}_ncu_full.txt"), "w") as f:
        f.write("\t".join(keep) + "\n" + "\t".join(units[i] for i in idx) + "\n")
        for r in rows[2:]:
            f.write("\t".join(r[i][:90] for i in idx) + "\n")
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[hdr.index("dram__bytes_read.sum")]]
            t = (float(r[hdr.index("dram__bytes_read.sum")]) + float(r[hdr.index("dram__bytes_write.sum")])) * scale
            kname = r[hdr.index("Kernel Name")].replace("void <unnamed>::", "")
            traffic.setdefault(key_of(kname), t)
            print(f"{kname[:64]:64s} {float(r[hdr.index('gpu__time_duration.sum')]):8.1f} {units[hdr.index('gpu__time_duration.sum')]}  "
                  f"dram {t / 1e6:8.1f} MB  issue {float(r[hdr.index('smsp__issue_active.avg.pct_of_peak_sustained_active')]):5.1f} %  "
                  f"warps {float(r[hdr.index('sm__warps_active.avg.pct_of_peak_sustained_active')]):5.1f} %")
    path = os.path.join(HERE, "traffic.json")
    # the capture command runs bench.py --pages 32: one launch = 32 pages; bench.py scales to its own launch size
    json.dump({"pages_per_launch": 32, "dram_bytes_per_launch": traffic}, open(path, "w"), indent=1)


if __name__ == "__main__":
    launch_shares(sys.argv[1])
    if len(sys.argv) > 2:
        full_report(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "r2")
