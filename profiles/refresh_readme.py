#!/usr/bin/env python3
"""Rewrites the machine-generated blocks of profiles/README.md (headline table rows that come from the bench JSONs, the
per-kernel CUDA-event table, the ncu launch-list table) from the committed evidence files.  The hand-written readings
(the `ncu --set full` table, the history) are left alone.

    python profiles/summarize.py profiles/r1_launches.csv gpurun_out/<capture>.ncu-rep > /tmp/summ.txt
    python profiles/refresh_readme.py /tmp/summ.txt
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    readme = os.path.join(HERE, "README.md")
    s = open(readme).read()
    d = json.load(open(os.path.join(HERE, "r1_bench_1gpu.json")))
    ref = json.load(open(os.path.join(HERE, "r1_bench_reference_arm.json")))
    kern, sk = d["roofline"]["kernels"], d["with_skew_estimate"]
    skrows = ", ".join(f"`{k}` {v}" for k, v in sk["kernels_ms"].items())
    rows = {
        "| `value` (device-resident, 1 GPU) |":
            f"| `value` (device-resident, 1 GPU) | **{d['value'] / 1e3:.1f} k input-MP/s**, {d['ms_per_step']:.2f} ms per 256-page step "
            f"({d['ms_per_step'] / 256 * 1e3:.0f} µs/page), {d['gpu_launches']} launches per 5 steps | `r1_bench_1gpu.json` |",
        "| `e2e` (pinned host buffers":
            f"| `e2e` (pinned host buffers through the C ABI, H2D ∥ kernels ∥ D2H) | **{d['e2e']['value'] / 1e3:.1f} k MP/s** — "
            f"{d['e2e']['h2d_bytes_per_step'] / 1e9:.2f} GB in (only the part of each photo under its quad; 9.22 GB before) + "
            f"{d['e2e']['d2h_bytes_per_step'] / 1e9:.2f} GB out per step ≈ 52 GB/s over PCIe: link-bound (raw contiguous pinned copies on the same "
            f"box: 52–56 GB/s up, 56–57 GB/s down, `tools/pcie_bandwidth.py`) | `r1_bench_1gpu.json` |",
        "| `with_skew_estimate`":
            f"| `with_skew_estimate` (every page's deskew angle estimated on the device: Canny + HoughLines + median) | "
            f"{sk['value'] / 1e3:.1f} k MP/s, {sk['ms_per_step']:.1f} ms per step; kernels (ms per 256 pages): {skrows} | `r1_bench_1gpu.json` |",
        "| `cpu_baseline`":
            f"| `cpu_baseline` (cv2 chain of DocScanner.py, {d['cpu_baseline']['cores']} host cores, page-parallel) | "
            f"{d['cpu_baseline']['value'] / 1e3:.2f} k MP/s | `r1_bench_1gpu.json` |",
        "| `--impl reference` arm":
            f"| `--impl reference` arm, same box | {ref['value'] / 1e3:.2f} k MP/s | `r1_bench_reference_arm.json` |",
    }
    lines = s.split("\n")
    for i, line in enumerate(lines):
        for prefix, new in rows.items():
            if line.startswith(prefix):
                lines[i] = new
    s = "\n".join(lines)
    a = s.index("| kernel | launches | ms | algorithmic GB/s |")
    b = s.index("These add up to")
    table = "\n".join(f"| `{k}` | {v['launches']} | {v['ms']} | {v['GB/s']} |" for k, v in kern.items())
    s = s[:a] + "| kernel | launches | ms | algorithmic GB/s |\n|---|---|---|---|\n" + table + "\n\n" + s[b:]
    tot = sum(v["ms"] for v in kern.values())
    s = re.sub(r"These add up to ≈ [0-9.]+ ms; the timed step takes [0-9.]+ ms",
               f"These add up to ≈ {tot:.1f} ms; the timed step takes {d['ms_per_step']:.1f} ms", s)
    if len(sys.argv) > 1:
        launch = "\n".join(l for l in open(sys.argv[1]).read().splitlines() if l.startswith("|"))
        a = s.index("| kernel | launches | total µs | share |")
        b = s.index("### One `ncu --set full` capture")
        s = s[:a] + launch + "\n\n" + s[b:]
    open(readme, "w").write(s)


if __name__ == "__main__":
    main()
