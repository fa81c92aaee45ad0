#!/usr/bin/env python3
"""bench.py — DocScanner per-pixel pipeline throughput (input megapixels/s) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the whole hot path (warp -> gray -> illumination -> stretch -> ink mask || adaptive
threshold -> blend -> rotate -> close; DocScanner.py:310-346, CLI defaults, scale_long=1600) over one batch
of 256 synthetic 12 MP uint8 BGR pages per GPU (BASELINE.json configs[1]); pages are independent, so N GPUs
each process their own batch (weak scaling, no collective on the data path).

`value`   : device-resident inputs/outputs, CUDA-event time on the launching stream, max over ranks.
`e2e`     : the same metric through the C-ABI call with HOST (pinned) buffers, H2D of every input page and
            D2H of both result images (warped, binary) inside the timed region, >= 10 steps; beside it `e2e.pageable`
            (plain numpy buffers, same call) and `e2e.pcie` (plain pinned copies of the same byte mix on every rank
            at once: the link ceiling the leg runs against).
`parity_checked`: two pages of every rank's timed batch compared with the oracle after the run (`with_skew_estimate.angles_checked`:
            likewise the angles the device estimated for those pages).
`--total-pages T`: BASELINE config 3 as written — page ids 0..T-1 sharded over the ranks by id (strong scaling).
`roofline`: the kernel with the largest share of the step, timed live with CUDA events in a second,
            instrumented pass (docscan_profile_enable) — algorithmic bytes / average launch time vs the
            measured HBM copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the reference's own CPU path (the cv2 call chain of DocScanner.py,
            oracle/ref_cv2.py) on this box's host cores, page-parallel over all cores.
`with_skew_estimate`: the same batch with every page's deskew angle estimated on the device (Canny + HoughLines +
            median: what a plain process_document(path) does), with its own `cpu_baseline` (the cv2 chain including
            deskew()'s estimate) beside it.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "docscanner_pipeline_input_megapixels_per_s"
PAGE_H, PAGE_W = 4000, 3000
PAGE_MP = PAGE_H * PAGE_W / 1e6
SCALE_LONG = 1600
WORKLOAD = ("synthetic 12 MP uint8 BGR page batch through all DocScanner stages "
            "(perspective warp, gray, illumination, stretch, ink mask, adaptive threshold, blend, rotate, close; "
            "CLI defaults, scale_long=1600)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--pages", type=int, default=256, help="pages per GPU per step")
    ap.add_argument("--e2e-distinct", type=int, default=16, help="distinct pinned host pages the e2e leg cycles through")
    ap.add_argument("--ref-pages", type=int, default=0, help="page-jobs per step of the CPU reference arm (0 = --pages, the same step as the GPU arm)")
    ap.add_argument("--total-pages", type=int, default=0,
                    help="BASELINE config 3 as written: this many page ids per step, sharded over the ranks by page id (strong scaling); "
                         "0 = the weak-scaling default of --pages per GPU")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run check of 2 pages per rank against the oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-skew", action="store_true", help="skip the extra leg with the device-side skew estimate")
    ap.add_argument("--scale-long", type=int, default=SCALE_LONG,
                    help="long side of the rectified page; 1600 = CLI default = the headline workload, 4000 = SURVEY config 2's full-resolution variant")
    ap.add_argument("--preset", choices=["cli", "gui"], default="cli", help="process_document tunables: CLI defaults or the GUI's call")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arm
_W = {}


# what holds each kernel family below the HBM roof today, from the ncu captures under profiles/ (reported as roofline.limiter)
KERNEL_LIMITER = {
    "warp_perspective_c3": "instruction issue + L1 gather (fp64 coordinate chain, 4 scattered 8-byte loads per pixel)",
    "warp_affine": "instruction issue + L1 gather",
    "adaptive_gauss": "fp32 pipe: the parity-fixed ordered fma chain",
    "blur_gauss": "instruction issue (integer MACs on CUDA cores)",
    "tc_blur": "TMEM read-out + epilogue issue (the banded contraction itself runs on the tensor cores)",
    "tc_adaptive": "TMEM read-out + epilogue issue",
    "morph_march": "shared-memory staging + barrier latency",
    "morph_close3_fused": "L2/HBM latency at 1 load per 16 px",
    "mask_blend": "HBM",
    "pw_lut": "HBM",
}


GUI_TUNABLES = dict(illum_method="divide", illum_blur_frac=0.05, block_size=31, C=3, morph_ksize=1, morph_iters=0)   # AI_classification.py:646-663


def _cpu_worker_init(path, quads, angles, scale_long=SCALE_LONG, tunables=None):
    """Worker process: one cv2 thread, pages memory-mapped from a scratch file."""
    from oracle import ref_cv2
    if ref_cv2.HAVE_CV2:
        import cv2
        cv2.setNumThreads(1)
    _W["pages"] = np.load(path, mmap_mode="r")
    _W["quads"], _W["angles"] = quads, angles
    _W["scale_long"], _W["tun"] = scale_long, dict(tunables or {})
    _W["fn"] = ref_cv2.hot_path if ref_cv2.HAVE_CV2 else None


def _cpu_worker_job(j):
    i = j % len(_W["quads"])
    page = np.asarray(_W["pages"][i])
    if _W["fn"] is not None:
        _W["fn"](page, _W["quads"][i], _W["angles"][i], scale_long=_W["scale_long"], **_W["tun"])
    else:
        from oracle import oracle as O
        O.hot_path(page, _W["quads"][i], _W["angles"][i], scale_long=_W["scale_long"], **_W["tun"])
    return 0


class CpuReference:
    """The reference's own CPU path — the cv2 call chain of DocScanner.py (oracle/ref_cv2.py) — page-parallel over
    every host core: os.cpu_count() spawned worker processes with one cv2 thread each (SURVEY.md 8d mode B, the
    faster of the two modes the survey measured).  Never forks after cv2/CUDA have been initialised."""

    def __init__(self, pages, quads, angles, scale_long=SCALE_LONG, tunables=None):
        import multiprocessing as mp
        import tempfile
        from oracle import ref_cv2
        self.cores = os.cpu_count() or 1
        base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
        fd, self.path = tempfile.mkstemp(suffix=".npy", dir=base)
        os.close(fd)
        np.save(self.path, np.stack(pages))
        self.n_distinct = len(pages)
        self.how = (f"cv2 call chain of DocScanner.py (oracle/ref_cv2.py), {self.cores} worker processes x 1 cv2 thread"
                    if ref_cv2.HAVE_CV2 else f"C oracle (cv2 not installed), {self.cores} worker processes")
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_worker_init,
                                                 initargs=(self.path, [np.asarray(q) for q in quads], list(angles), scale_long, tunables))
        self.run(2 * self.cores)                                   # imports + first-touch, not timed

    def run(self, jobs: int):
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker_job, range(jobs), chunksize=1)
        dt = time.perf_counter() - t0
        return jobs * PAGE_MP / dt, dt

    def close(self):
        self.pool.close()
        self.pool.join()
        try:
            os.remove(self.path)
        except OSError:
            pass


def _workload(args):
    if args.scale_long == SCALE_LONG and args.preset == "cli":
        return WORKLOAD
    return WORKLOAD.replace("CLI defaults, scale_long=1600", f"{'GUI preset' if args.preset == 'gui' else 'CLI defaults'}, scale_long={args.scale_long}")


def _config(args, n_gpus):
    """The `config` object: identical in both arms for the same command line (the driver compares them)."""
    from smart_image_processing_b200 import DocScanner as DS
    from smart_image_processing_b200.synth import synth_quad
    tw, th = DS.target_size(synth_quad(0, PAGE_W, PAGE_H), "A4", args.scale_long)
    cfg = {"workload": _workload(args), "pages_per_gpu": args.pages, "page": f"{PAGE_H}x{PAGE_W}x3", "warped": f"{th}x{tw}",
           "scale_long": args.scale_long, "parallelism": f"pages sharded over {n_gpus} GPU(s) by page id, no collective",
           "l2": f"inputs larger than L2 ({args.pages * PAGE_H * PAGE_W * 3 / 1e9:.1f} GB read per step per GPU)"}
    if args.total_pages:
        cfg["total_pages"] = args.total_pages
    return cfg


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from smart_image_processing_b200.synth import synth_angle, synth_page_numpy
    distinct = 2
    pages, quads, angles = [], [], []
    for s in range(distinct):
        img, quad = synth_page_numpy(s, PAGE_W, PAGE_H)
        pages.append(img); quads.append(quad); angles.append(synth_angle(s))
    tun = GUI_TUNABLES if args.preset == "gui" else {}
    ref = CpuReference(pages, quads, angles, args.scale_long, tun)
    # a step of the reference arm is the GPU arm's step (args.pages page-jobs), bounded so that the run ends within minutes:
    # a first probe measures the rate, and steps that would take more than ~6 s each are cut to a sample of the batch
    ref_pages = args.ref_pages or args.pages
    _, dt1 = ref.run(2 * ref.cores)
    per_job = dt1 / (2 * ref.cores)
    budget_s = 150.0 / max(1, args.steps + args.warmup)
    if ref_pages * per_job > budget_s:
        ref_pages = max(ref.cores, int(budget_s / per_job))
    for _ in range(args.warmup):
        ref.run(max(ref.cores, 8))
    total = 0.0
    for _ in range(args.steps):
        total += ref.run(ref_pages)[1]
    ref.close()
    value = args.steps * ref_pages * PAGE_MP / total
    sample = f"{ref_pages} page-jobs per step over {distinct} distinct synthetic 12 MP pages; {ref.how}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.total_pages else "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": _config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "MP/s", "cores": ref.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
def _bind_to_gpu_numa_node(torch, local):
    """With several ranks on one box every rank's pinned host buffers should live on the NUMA node its GPU hangs off
    (first touch decides), otherwise the H2D / D2H copies of the e2e leg cross the socket interconnect.  Pins this process
    to the GPU's local CPUs before anything is allocated; returns the cpulist string, or None when sysfs does not say."""
    try:
        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        cpulist = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpulist
    except Exception:
        pass
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from smart_image_processing_b200 import DocScanner as DS
    from smart_image_processing_b200 import _capi
    from smart_image_processing_b200 import sharding
    from smart_image_processing_b200.synth import synth_angle

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = _bind_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL prints its version banner with printf while the communicator is
        # created (NCCL_DEBUG=VERSION on some boxes), so file descriptor 1 points at stderr until that is done
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    stream = torch.cuda.Stream(device=dev)
    ctx = _capi.Context(local, stream=stream.cuda_stream)
    tun = GUI_TUNABLES if args.preset == "gui" else {}
    params = DS.make_params(**tun)

    # ---- device-resident synthetic batch (generated on the device; not timed)
    # weak scaling (default): every rank owns args.pages pages (seeds rank*P ...).  --total-pages T (BASELINE config 3):
    # page ids 0..T-1 are sharded by id (sharding.page_ids); a rank keeps the first args.pages of its ids resident (T photos
    # would be 147 GB) and runs its other page-jobs on them again, each job writing outputs of its own.
    if args.total_pages:
        ids = sharding.page_ids(args.total_pages, rank, world)
        P = len(ids)                                   # page-jobs of this rank per step
        D = min(args.pages, P)
        seeds = list(ids[:D])
    else:
        P = D = args.pages
        seeds = list(sharding.weak_batch_seeds(P, rank))
    src = torch.empty((D, PAGE_H, PAGE_W, 3), dtype=torch.uint8, device=dev)
    quads, angles = [], []
    q8 = (C.c_float * 8)()
    for i in range(D):
        seed = seeds[i]
        im = _capi.device_image(src[i].data_ptr(), PAGE_W, PAGE_H, PAGE_W * 3, 3)
        ctx.call("docscan_synth_page", C.c_uint64(seed), C.byref(im), q8)
        quads.append(np.array(list(q8), np.float32).reshape(4, 2))
        angles.append(synth_angle(seed))
    sizes = [DS.target_size(q, "A4", args.scale_long) for q in quads]
    tw, th = sizes[0]
    assert all(s == (tw, th) for s in sizes)
    pw3, pw1 = (tw * 3 + 127) // 128 * 128, (tw + 127) // 128 * 128
    warped = torch.empty((P, th, pw3), dtype=torch.uint8, device=dev)
    binary = torch.empty((P, th, pw1), dtype=torch.uint8, device=dev)
    pages = (_capi.Page * P)()
    for i in range(P):
        d = i % D
        pages[i].src = _capi.device_image(src[d].data_ptr(), PAGE_W, PAGE_H, PAGE_W * 3, 3)
        pages[i].quad = (C.c_float * 8)(*quads[d].reshape(8).tolist())
        pages[i].angle_deg = angles[d]
        pages[i].warped = _capi.device_image(warped[i].data_ptr(), tw, th, pw3, 3)
        pages[i].binary = _capi.device_image(binary[i].data_ptr(), tw, th, pw1, 1)
    ctx.sync()
    total_jobs = args.total_pages if args.total_pages else world * P     # page-jobs of the whole job per step

    def step():
        ctx.call("docscan_process_pages", P, pages, C.byref(params))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches
    with ClockSampler(local) as clocks:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
    launches = ctx.launches - l0
    ms_total = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    value = total_jobs * PAGE_MP * args.steps / (ms_total / 1e3)

    # ---- parity of the timed batch: two pages of every rank (first and last resident page) against the oracle
    # (oracle/ is the checker here, never the thing measured); all ranks must agree for parity_checked to be true
    parity = None
    if not args.no_parity:
        from oracle import oracle as O

        def _host(t, rows, width_bytes):
            return np.ascontiguousarray(t.cpu().numpy()[:rows, :width_bytes])

        ok = 1
        for i in sorted({0, D - 1}):
            img = src[i].cpu().numpy()
            ref = O.hot_path(img, quads[i], angles[i], scale_long=args.scale_long, **tun)
            j = i if i < P else 0
            got_w = _host(warped[j], th, tw * 3).reshape(th, tw, 3)
            got_b = _host(binary[j], th, tw)
            if not (np.array_equal(got_w, ref["warped"]) and np.array_equal(got_b, ref["clean"])):
                ok = 0
        if world > 1:
            t = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = int(t.item())
        parity = bool(ok)

    # ---- per-kernel pass (instrumented; not the number reported as `value`)
    ctx.profile(True)
    step()
    prof = ctx.profile_dump()
    ctx.profile(False)
    kernel_ms = sum(v[1] for v in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1][1])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    n_l, t_ms, nbytes = top[1]
    achieved = nbytes / n_l / (t_ms / n_l * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj["dram_bytes_per_launch"].get(top[0].split("_k")[0])
        if traffic is not None:            # ncu captured launches of tj["pages_per_launch"] pages: scale to this run's launches
            traffic = traffic * (P / n_l) / tj["pages_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "share_of_step": t_ms / kernel_ms,
                "algorithmic_bytes_per_launch": nbytes / n_l, "avg_launch_ms": t_ms / n_l,
                "limiter": KERNEL_LIMITER.get(top[0].split("_k")[0], "see profiles/README.md"),
                "note": "bound = the roof this byte-oriented path is measured against (HBM copy bandwidth; no dense contraction "
                        "on it); limiter = what ncu shows holding this kernel below that roof today",
                "kernels": {k: {"launches": v[0], "ms": round(v[1], 4), "GB/s": (round(v[2] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None)}
                            for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}

    # ---- the same batch with deskew()'s own skew estimate (Canny + HoughLines median) computed on the device for every
    # page instead of a supplied angle; reported beside the headline, not as it
    skew = None
    if not args.no_skew:
        nan = float("nan")
        for i in range(P):
            pages[i].angle_deg = nan
        step(); step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ssteps = max(1, min(args.steps, 3))
        s0.record(stream)
        for _ in range(ssteps):
            step()
        s1.record(stream)
        barrier()
        sms = sharding.max_over_ranks(s0.elapsed_time(s1), dev)
        ctx.profile(True)
        step()
        sprof = ctx.profile_dump()
        ctx.profile(False)
        est = (C.c_double * P)()
        ctx.call("docscan_last_angles", est, P)
        skew = {"value": total_jobs * PAGE_MP * ssteps / (sms / 1e3), "unit": "MP/s", "ms_per_step": sms / ssteps, "steps": ssteps,
                "what": "angle_deg = NaN for every page: Canny + HoughLines(1, pi/180, 150) + median angle on the device between blend and rotate",
                "kernels_ms": {k: round(v[1], 4) for k, v in sorted(sprof.items(), key=lambda kv: -kv[1][1])
                               if k.startswith(("canny", "hough", "skew"))},
                "angles_estimated_sample": [float(est[i]) for i in range(min(P, 4))]}
        if not args.no_parity:
            # the estimated angles of the first and last resident page against the oracle's deskew() estimate
            from oracle import oracle as O
            pix = {k: v for k, v in tun.items() if k not in ("canny_low", "canny_high", "max_rotate")}
            ok = 1
            for i in sorted({0, D - 1}):
                st = O.hot_path(src[i].cpu().numpy(), quads[i], 0.0, scale_long=args.scale_long, **pix)
                want = O.estimate_skew_angle(st["weighted"], tun.get("canny_low", 50), tun.get("canny_high", 150), tun.get("max_rotate", 10.0))
                if float(est[i if i < P else 0]) != want:
                    ok = 0
            if world > 1:
                t = torch.tensor([ok], dtype=torch.int32, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                ok = int(t.item())
            skew["angles_checked"] = bool(ok)
        for i in range(P):
            pages[i].angle_deg = angles[i % D]
        step()                                  # leave the supplied-angle results in the device-resident outputs again
        barrier()

    # ---- e2e: HOST buffers through the same C-ABI call, copies inside the timed region.  Headline leg: pinned buffers
    # (docscan_host_alloc), >= 10 steps.  Beside it: the same call with PAGEABLE numpy buffers (what cv2.imread hands a caller),
    # and the box's raw concurrent pinned H2D + D2H rate measured on every rank at once (the ceiling the leg runs against).
    e2e = None
    if not args.no_e2e:
        De = min(args.e2e_distinct, D)

        def host_pages(alloc):
            h_src = [alloc((PAGE_H, PAGE_W, 3)) for _ in range(De)]
            h_w = [alloc((th, tw, 3)) for _ in range(De)]
            h_b = [alloc((th, tw)) for _ in range(De)]
            for i in range(De):
                _capi.lib().docscan_memcpy_d2h(ctx._h, h_src[i].ctypes.data, C.c_void_p(src[i].data_ptr()), h_src[i].nbytes)
            hp = (_capi.Page * P)()
            for i in range(P):
                d = i % De
                hp[i].src = _capi.image_of(h_src[d])
                hp[i].quad = (C.c_float * 8)(*quads[d].reshape(8).tolist())
                hp[i].angle_deg = angles[d]
                hp[i].warped = _capi.image_of(h_w[d])
                hp[i].binary = _capi.image_of(h_b[d])
            return hp, (h_src, h_w, h_b)

        def timed_host(hp, nsteps):
            ctx.call("docscan_process_pages", P, hp, C.byref(params))
            barrier()
            tb0 = ctx.transfer_bytes
            t0 = time.perf_counter()
            e0.record(stream)
            for _ in range(nsteps):
                ctx.call("docscan_process_pages", P, hp, C.byref(params))
            e1.record(stream)
            barrier()
            wall = time.perf_counter() - t0
            tb1 = ctx.transfer_bytes
            ms = sharding.max_over_ranks(max(e0.elapsed_time(e1), wall * 1e3), dev)
            return total_jobs * PAGE_MP * nsteps / (ms / 1e3), (tb1[0] - tb0[0]) // nsteps, (tb1[1] - tb0[1]) // nsteps

        hpages, keep_pinned = host_pages(ctx.pinned_empty)
        esteps = max(10, args.steps)
        ev, h2d, d2h = timed_host(hpages, esteps)
        # results of the host path == results of the device-resident path (same page, same bytes)
        same = bool(np.array_equal(np.asarray(keep_pinned[2][0]), binary[0].cpu().numpy()[:th, :tw]))
        e2e = {"value": ev, "unit": "MP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": esteps,
               "matches_device_resident": same,
               "host_buffers": f"pinned; {De} distinct pages cycled, every page copied every step; the library uploads only the "
                               f"rows/columns of each {PAGE_H * PAGE_W * 3} B photo under its quad (bytes counted by the library)"}
        # raw link rate, all ranks at once: plain pinned copies of the step's own byte mix (H2D and D2H concurrently, in the
        # proportion the step moves them) -- the ceiling the leg above runs against
        n_up = 1 << 30
        n_dn = max(1 << 20, int(n_up * (d2h / max(h2d, 1))) & ~0xFFFFF)
        h_up, h_dn = torch.empty(n_up, dtype=torch.uint8).pin_memory(), torch.empty(n_dn, dtype=torch.uint8).pin_memory()
        d_up, d_dn = torch.empty(n_up, dtype=torch.uint8, device=dev), torch.empty(n_dn, dtype=torch.uint8, device=dev)
        s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def link_round(reps):
            for _ in range(reps):
                with torch.cuda.stream(s_up):
                    d_up.copy_(h_up, non_blocking=True)
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
            torch.cuda.synchronize(dev)

        link_round(2)
        barrier()
        reps = 6
        t0 = time.perf_counter()
        link_round(reps)
        link_s = sharding.max_over_ranks(time.perf_counter() - t0, dev)
        link_gbs = reps * n_up / link_s / 1e9                       # H2D rate per rank with every rank copying and the D2H share alongside
        del h_up, h_dn, d_up, d_dn
        floor_s = h2d / (link_gbs * 1e9)                            # a step cannot beat its uploads at that rate
        e2e["pcie"] = {"h2d_GBps_per_rank_all_ranks_copying": round(link_gbs, 2),
                       "with_concurrent_d2h_share": round(n_dn / n_up, 3),
                       "step_floor_ms": round(floor_s * 1e3, 2),
                       "e2e_frac_of_link_floor": round((floor_s * 1e3) / (total_jobs * PAGE_MP / ev * 1e3), 3)}
        del hpages, keep_pinned
        # pageable buffers (plain numpy arrays)
        ppages, keep_pageable = host_pages(lambda shape: np.empty(shape, np.uint8))
        pv, _, _ = timed_host(ppages, 2)
        e2e["pageable"] = {"value": pv, "unit": "MP/s", "steps": 2, "host_buffers": "plain numpy arrays (pageable), same call"}
        del ppages, keep_pageable

    # ---- CPU baseline beside it (rank 0, single-GPU run only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Dc = min(4, D)
        hp = [np.empty((PAGE_H, PAGE_W, 3), np.uint8) for _ in range(Dc)]
        for i in range(Dc):
            _capi.lib().docscan_memcpy_d2h(ctx._h, hp[i].ctypes.data, C.c_void_p(src[i].data_ptr()), hp[i].nbytes)
        ref = CpuReference(hp, quads[:Dc], angles[:Dc], args.scale_long, tun)
        _, dt1 = ref.run(2 * ref.cores)
        jobs = int(max(2 * ref.cores, min(4096, 12.0 / max(dt1 / (2 * ref.cores), 1e-6))))
        v, dt = ref.run(jobs)
        ref.close()
        cpu = {"value": v, "unit": "MP/s", "cores": ref.cores, "kind": "port",
               "sample": f"{jobs} page-jobs over {Dc} distinct pages of this run's batch, {dt:.1f} s; {ref.how}"}
        if skew is not None:
            # the same chain with deskew()'s own estimate (Canny + HoughLines + median, DocScanner.py:218-231) beside the GPU leg
            # that estimates every page's angle on the device: a bounded sample, about 8 s
            ref = CpuReference(hp, quads[:Dc], [None] * Dc, args.scale_long, tun)
            _, dt1 = ref.run(2 * ref.cores)
            jobs = int(max(2 * ref.cores, min(2048, 8.0 / max(dt1 / (2 * ref.cores), 1e-6))))
            v2, dt2 = ref.run(jobs)
            ref.close()
            skew["cpu_baseline"] = {"value": v2, "unit": "MP/s", "cores": ref.cores, "kind": "port",
                                    "sample": f"{jobs} page-jobs, {dt2:.1f} s; the same cv2 chain with deskew()'s own Canny + HoughLines estimate"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if args.total_pages else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": _config(args, world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "with_skew_estimate": skew, "gpu_launches": int(launches),
            "parity_checked": parity, "page_jobs_per_step": total_jobs, "rank0_cpu_affinity": numa,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
