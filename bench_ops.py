#!/usr/bin/env python3
"""bench_ops.py — stage-level sweeps of BASELINE.json configs[3] and configs[4] on one B200.

  config 4: one 3840x2160 u8 gray; erode / dilate / MORPH_CLOSE (rect k x k) and adaptiveThreshold (gaussian, mean,
            C=10) for k = 3, 5, .., 31 (+ morph_seq's k = 2)
  config 5: one 7680x4320 u8 gray; illumination_correction(method="divide") with k in {101, 151, 217} and the
            close variant normalize(divide(g, close(g, k), 255))

Every op runs device-resident through the C ABI (CUDA events, 20 iterations after 3 warm-ups, input larger than or
comparable to L2 is not guaranteed here: these are single-image latency numbers, L2-warm), is compared bit-for-bit
with cv2 on the same image when cv2 is installed, and is timed against cv2 on the host cores.
Writes one JSON line per op and a markdown table (profiles/ops_sweep.md with --write).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def synth_gray(h, w, seed=0):
    """SURVEY 8d gray generator: illumination gradient, 25 % ink blocks (16x4 cells), N(0,3) noise."""
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:h, 0:w]
    base = 235.0 * (0.55 + 0.45 * (0.6 * xs / w + 0.4 * ys / h))
    ink = rng.random((h // 4 + 1, w // 16 + 1)) < 0.25
    ink = np.repeat(np.repeat(ink, 4, 0), 16, 1)[:h, :w]
    img = np.where(ink, base * 0.25, base) + rng.normal(0, 3, (h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import torch
    from smart_image_processing_b200 import _capi
    try:
        import cv2
    except ImportError:
        cv2 = None
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    ctx = _capi.Context(0, stream=stream.cuda_stream)
    peak = 6546.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rows = []

    def dimg(t):
        h, w = t.shape
        return _capi.device_image(t.data_ptr(), w, h, t.stride(0), 1)

    def run(name, cfg, host, call, ref, alg_bytes_per_px=2.0):
        h, w = host.shape
        pitch = (w + 127) // 128 * 128
        src = torch.zeros((h, pitch), dtype=torch.uint8, device=dev)[:, :w]
        src.copy_(torch.from_numpy(host).to(dev))
        dst = torch.zeros((h, pitch), dtype=torch.uint8, device=dev)[:, :w]
        s, d = dimg(src), dimg(dst)
        for _ in range(3):
            call(s, d)
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.iters):
            call(s, d)
        e1.record(stream)
        ctx.sync()
        ms = e0.elapsed_time(e1) / args.iters
        out = dst.cpu().numpy()
        mism, cpu_ms = None, None
        if cv2 is not None:
            r = ref(host)
            mism = int(np.count_nonzero(r != out))
            n = 3 if h * w > 2e7 else 5
            t0 = time.perf_counter()
            for _ in range(n):
                ref(host)
            cpu_ms = (time.perf_counter() - t0) / n * 1e3
        mp = h * w / 1e6
        row = {"config": cfg, "op": name, "gpu_ms": round(ms, 4), "gpu_MP/s": round(mp / ms * 1e3, 1),
               "alg_GB/s": round(alg_bytes_per_px * h * w / ms / 1e6, 1), "hbm_frac": round(alg_bytes_per_px * h * w / ms / 1e6 / peak, 4),
               "cv2_ms": None if cpu_ms is None else round(cpu_ms, 2), "cv2_MP/s": None if cpu_ms is None else round(mp / cpu_ms * 1e3, 1),
               "mismatching_px_vs_cv2": mism}
        rows.append(row)
        print(json.dumps(row), flush=True)

    lib = ctx
    g4 = synth_gray(2160, 3840, 1)
    ks = [2] + list(range(3, 32, 2))
    if args.quick:
        ks = [2, 3, 9, 31]
    for k in ks:
        se = None if cv2 is None else cv2.getStructuringElement(cv2.MORPH_RECT, (k, k))
        for opname, opcode, cvf in (("erode", _capi.MORPH_ERODE, lambda a, se=se: cv2.erode(a, se)),
                                    ("dilate", _capi.MORPH_DILATE, lambda a, se=se: cv2.dilate(a, se)),
                                    ("close", _capi.MORPH_CLOSE, lambda a, se=se: cv2.morphologyEx(a, cv2.MORPH_CLOSE, se))):
            run(f"{opname} {k}x{k}", 4, g4,
                lambda s, d, opcode=opcode, k=k: lib.call("docscan_morph_rect", opcode, C.byref(s), k, k, 1, C.byref(d)), cvf)
    for k in ([3, 11, 31] if args.quick else range(3, 32, 2)):
        run(f"adaptive gaussian k={k}", 4, g4,
            lambda s, d, k=k: lib.call("docscan_adaptive_threshold", C.byref(s), _capi.ADAPTIVE_GAUSSIAN, k, 10, 1, C.byref(d)),
            lambda a, k=k: cv2.adaptiveThreshold(a, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, k, 10))
        run(f"adaptive mean k={k}", 4, g4,
            lambda s, d, k=k: lib.call("docscan_adaptive_threshold", C.byref(s), _capi.ADAPTIVE_MEAN, k, 10, 1, C.byref(d)),
            lambda a, k=k: cv2.adaptiveThreshold(a, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, k, 10))
    # the rows added this round: Canny, the whole skew estimate, resize_long_side
    run("canny 50/150", 4, g4, lambda s, d: lib.call("docscan_canny", C.byref(s), 50.0, 150.0, C.byref(d)),
        lambda a: cv2.Canny(a, 50, 150), alg_bytes_per_px=2.0)
    ang = C.c_double()

    def skew(s, d):
        lib.call("docscan_skew_angle", C.byref(s), 50.0, 150.0, 10.0, C.byref(ang))

    def skew_ref(a):
        lines = cv2.HoughLines(cv2.Canny(a, 50, 150), 1, np.pi / 180, 150)
        return np.zeros_like(a)          # timing only; the angle is compared below

    run("skew estimate (Canny + HoughLines + median)", 4, g4, skew, skew_ref, alg_bytes_per_px=1.0)
    rows[-1]["mismatching_px_vs_cv2"] = None
    if cv2 is not None:
        from oracle import oracle as O          # the checker: numpy-float32 median of cv2's own lines
        lines = cv2.HoughLines(cv2.Canny(g4, 50, 150), 1, np.pi / 180, 150)
        per = np.zeros(180, np.int32)
        if lines is not None:
            for th in lines[:, 0, 1]:
                per[int(round(float(th) / (np.pi / 180)))] += 1
        rows[-1]["mismatching_px_vs_cv2"] = 0 if O.median_angle(per) == ang.value else 1
    g5 = synth_gray(4320, 7680, 2)
    for k in (101, 151, 217):
        run(f"illumination divide k={k}", 5, g5,
            lambda s, d, k=k: lib.call("docscan_illumination_correction", C.byref(s), 1, k, C.byref(d)),
            lambda a, k=k: cv2.normalize(cv2.divide(a, cv2.GaussianBlur(a, (k, k), 0), scale=255), None, 0, 255, cv2.NORM_MINMAX))
        h, w = g5.shape
        pitch = (w + 127) // 128 * 128
        tmp = torch.zeros((h, pitch), dtype=torch.uint8, device=dev)[:, :w]
        t = dimg(tmp)

        def close_divide(s, d, k=k, t=t):
            lib.call("docscan_morph_rect", _capi.MORPH_CLOSE, C.byref(s), k, k, 1, C.byref(t))
            lib.call("docscan_binary_op", _capi.OP_DIV255, C.byref(s), C.byref(t), C.byref(t))
            lib.call("docscan_normalize_minmax", C.byref(t), C.byref(d))

        run(f"close+divide+normalize k={k}", 5, g5, close_divide,
            lambda a, k=k: cv2.normalize(cv2.divide(a, cv2.morphologyEx(a, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (k, k))),
                                                    scale=255), None, 0, 255, cv2.NORM_MINMAX))
    # ---- config 1: public/sample.jpg (committed as tests/golden/sample_bgr.npz), whole per-pixel path, single-image latency
    # through the numpy drop-in (ordinary pageable arrays in and out), with the reference's quad / angle and with the
    # device-side skew estimate; cv2 chain of DocScanner.py on the host beside it (all cores / one core)
    lat = []
    try:
        from smart_image_processing_b200 import DocScanner as DS
        from oracle import ref_cv2
        img = np.load(os.path.join(ROOT, "tests", "golden", "sample_bgr.npz"))["bgr"]
        meta = json.load(open(os.path.join(ROOT, "tests", "golden", "sample_golden.json")))
        for preset in ("cli", "gui"):
            p = meta["presets"][preset]
            quad = np.frombuffer(bytes.fromhex(p["quad_f32_hex"]), np.float32).reshape(4, 2)
            angle = float.fromhex(p["angle_hex"])
            tun = {k: v for k, v in p["params"].items() if k not in ("canny_low", "canny_high", "max_rotate")}
            skew = {k: p["params"][k] for k in ("canny_low", "canny_high", "max_rotate")}

            def t_ms(fn, n=20):
                fn(); fn()
                t0 = time.perf_counter()
                for _ in range(n):
                    fn()
                return (time.perf_counter() - t0) / n * 1e3

            ms_given = t_ms(lambda: DS.process_pages([img], [quad], [angle], **tun))
            ms_est = t_ms(lambda: DS.process_pages([img], [quad], [None], **tun, **skew))
            w, b, used = DS.process_pages([img], [quad], [None], return_angles=True, **tun, **skew)
            row = {"config": 1, "preset": preset, "image": "public/sample.jpg 1280x963", "gpu_ms_angle_given": round(ms_given, 3),
                   "gpu_ms_angle_estimated_on_device": round(ms_est, 3), "angle_used": used[0], "angle_reference": angle}
            if cv2 is not None:
                cpu = ref_cv2.hot_path(img, quad, angle, **tun)
                row["mismatching_px_vs_cv2"] = int(np.count_nonzero(cpu[1] != DS.process_pages([img], [quad], [angle], **tun)[1][0]))
                row["cv2_ms_angle_given_all_cores"] = round(t_ms(lambda: ref_cv2.hot_path(img, quad, angle, **tun), 10), 2)

                def cv_skew():
                    st = ref_cv2.hot_path(img, quad, 0.0, keep_stages=True, **tun)
                    lines = cv2.HoughLines(cv2.Canny(st["weighted"], skew["canny_low"], skew["canny_high"]), 1, np.pi / 180, 150)
                    return lines

                row["cv2_ms_with_its_own_skew_estimate"] = round(t_ms(cv_skew, 5), 2)
            lat.append(row)
            print(json.dumps(row), flush=True)
    except Exception as e:          # the sweep tables above do not depend on this block
        print(json.dumps({"config": 1, "error": repr(e)}), flush=True)
    if args.write:
        with open(os.path.join(ROOT, "profiles", "latency_sample_jpg.json"), "w") as f:
            json.dump(lat, f, indent=1)
        path = os.path.join(ROOT, "profiles", "ops_sweep.md")
        with open(path, "w") as f:
            f.write("# Stage-level sweeps (BASELINE.json configs 4 and 5), one B200, device-resident, CUDA events\n\n")
            f.write(f"cv2 columns: same op on this box's host cores (cv2 default threads, {os.cpu_count()} cores). "
                    "`mismatching px` = full-size bit-for-bit comparison of the GPU result with cv2.\n\n")
            f.write("| cfg | op | GPU ms | GPU MP/s | alg GB/s | of 6546 GB/s | cv2 ms | cv2 MP/s | mismatching px |\n|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write(f"| {r['config']} | {r['op']} | {r['gpu_ms']} | {r['gpu_MP/s']} | {r['alg_GB/s']} | {r['hbm_frac']} | {r['cv2_ms']} | {r['cv2_MP/s']} | {r['mismatching_px_vs_cv2']} |\n")
        print("wrote", path)


if __name__ == "__main__":
    main()
