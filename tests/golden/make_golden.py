#!/usr/bin/env python3
"""Generates the fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (it needs /root/reference and cv2):

    python tests/golden/make_golden.py

What it writes (all small, all committed):

  sample_bgr.npz          decoded public/sample.jpg (BASELINE.json config 1 input)
  sample_golden.json      per preset: quad + deskew angle produced by the reference's control path
                          (localize_document / Canny+HoughLines inside deskew) and the sha256 of
                          every stage image the reference's DocScanner functions produce
  sample_<preset>_bin.npz the bilevel stage images themselves (ink mask, adaptive, blend, deskew, clean)
  crops.npz               three crops of sample.jpg pushed through every reference stage function with
                          synthetic quads/angles; all stage outputs stored in full
  kat.npz                 the reference's own committed artefacts: outputs/morphseq_01_gray.png ->
                          morphseq_02_eroded.png (KAT-1) and the constant scan_03..08 chain (KAT-2)
  gauss_kernels.npz       cv2.getGaussianKernel(k, 0, CV_32F) for every odd k <= 255
  ops.npz                 per-op known answers on small random / structured inputs (cv2 outputs)

The reference functions are imported from /root/reference/DocScanner.py unmodified; nothing is copied.
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
REF = "/root/reference"
sys.path.insert(0, REF)
import cv2  # noqa: E402
import DocScanner as DS  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

PRESETS = {
    # DocScanner.py:262-276 defaults (the CLI preset)
    "cli": dict(page="A4", scale_long=1600, canny_low=50, canny_high=150, illum_method="subtract",
                illum_blur_frac=0.02, block_size=35, C=10, thresh_method="gaussian", mask_blur_ksize=51,
                blackhat_ksize=9, blackhat_vertical_ratio=2.0, ink_dilate_iters=1, mask_thresh_offset=8,
                morph_ksize=3, morph_iters=1, max_rotate=10.0),
    # AI_classification.py:646-663 (the GUI preset)
    "gui": dict(page="A4", scale_long=1200, canny_low=30, canny_high=100, illum_method="divide",
                illum_blur_frac=0.05, block_size=31, C=3, thresh_method="gaussian", mask_blur_ksize=51,
                blackhat_ksize=9, blackhat_vertical_ratio=2.0, ink_dilate_iters=1, mask_thresh_offset=8,
                morph_ksize=1, morph_iters=0, max_rotate=10.0),
}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def deskew_angle(gray, canny_low, canny_high, max_rotate):
    """The control-path half of DS.deskew (DocScanner.py:218-231), evaluated with the same calls so
    that the fixture records the exact Python float the reference hands to getRotationMatrix2D."""
    edges = cv2.Canny(gray, canny_low, canny_high)
    lines = cv2.HoughLines(edges, 1, np.pi / 180, 150)
    angle_deg = 0.0
    if lines is not None and len(lines) > 0:
        angles = []
        for rho, theta in lines[:, 0, :]:
            ang = (theta * 180.0 / np.pi)
            ang = (ang + 90.0) % 180.0 - 90.0
            angles.append(ang)
        if angles:
            angle_deg = float(np.median(angles))
            if abs(angle_deg) > max_rotate:
                angle_deg = 0.0
    return angle_deg


def run_reference_stages(color, quad, p, angle=None):
    """Chains the reference's stage functions exactly like process_document (DocScanner.py:310-346)."""
    out = {}
    out["warped"] = DS.perspective_warp(color, quad, page=p["page"], scale_long=p["scale_long"])
    out["gray"] = cv2.cvtColor(out["warped"], cv2.COLOR_BGR2GRAY)
    out["illum"] = DS.illumination_correction(out["gray"], method=p["illum_method"], blur_frac=p["illum_blur_frac"])
    out["stretch"] = DS.contrast_stretch(out["illum"])
    out["inkmask"] = DS._compute_ink_mask(out["stretch"], mask_blur_ksize=p["mask_blur_ksize"],
                                          blackhat_ksize=p["blackhat_ksize"],
                                          blackhat_vertical_ratio=p["blackhat_vertical_ratio"],
                                          dilate_iters=p["ink_dilate_iters"],
                                          threshold_offset=p["mask_thresh_offset"])
    out["adapt"] = DS.adaptive_binarize(out["stretch"], block_size=p["block_size"], C=p["C"],
                                        method=p["thresh_method"])
    b = out["adapt"].copy()
    b[out["inkmask"] == 0] = 255
    out["weighted"] = b
    if angle is None:
        angle = deskew_angle(b, p["canny_low"], p["canny_high"], p["max_rotate"])
        out["deskew"] = DS.deskew(b, canny_low=p["canny_low"], canny_high=p["canny_high"], max_rotate=p["max_rotate"])
    else:
        h, w = b.shape[:2]
        m = cv2.getRotationMatrix2D((w / 2.0, h / 2.0), angle, 1.0)
        out["deskew"] = cv2.warpAffine(b, m, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
    out["clean"] = DS.morph_cleanup(out["deskew"], ksize=p["morph_ksize"], iterations=p["morph_iters"])
    return out, angle


def main():
    rng = np.random.default_rng(20261018)
    color = DS.load_image(os.path.join(REF, "public", "sample.jpg"))
    np.savez_compressed(os.path.join(HERE, "sample_bgr.npz"), bgr=color)

    # ---- config 1: sample.jpg through the reference, both presets
    meta = {"input_sha256": sha(color), "shape": list(color.shape), "cv2": cv2.__version__,
            "numpy": np.__version__, "presets": {}}
    for name, p in PRESETS.items():
        quad = DS.localize_document(color, canny_low=p["canny_low"], canny_high=p["canny_high"])
        stages, angle = run_reference_stages(color, quad, p)
        # process_document itself must agree with the chained functions (DocScanner.py:262-365)
        res = DS.process_document(os.path.join(REF, "public", "sample.jpg"), out_dir="/tmp/golden_dump_" + name,
                                  **{k: v for k, v in p.items()})
        assert np.array_equal(res["binary"], stages["clean"]) and np.array_equal(res["warped"], stages["warped"])
        meta["presets"][name] = {
            "params": p,
            "quad": [[float(v) for v in pt] for pt in quad],
            "quad_f32_hex": np.asarray(quad, np.float32).tobytes().hex(),
            "angle": angle, "angle_hex": float(angle).hex(),
            "shapes": {k: list(v.shape) for k, v in stages.items()},
            "sha256": {k: sha(v) for k, v in stages.items()},
        }
        np.savez_compressed(os.path.join(HERE, f"sample_{name}_bin.npz"),
                            **{k: stages[k] for k in ("inkmask", "adapt", "weighted", "deskew", "clean")})
    with open(os.path.join(HERE, "sample_golden.json"), "w") as f:
        json.dump(meta, f, indent=1)

    # ---- crops with synthetic quads (full stage outputs stored)
    crops = {}
    specs = [
        ("a", (200, 100, 360, 291), dict(PRESETS["cli"], scale_long=400), 1.5),
        ("b", (500, 300, 333, 250), dict(PRESETS["gui"], scale_long=333), -2.5),
        ("c", (60, 420, 301, 407), dict(PRESETS["cli"], scale_long=517, thresh_method="mean", illum_method="divide",
                                       block_size=21, C=7, morph_ksize=2, morph_iters=2, page="custom"), 0.0),
    ]
    for tag, (y0, x0, hh, ww), p, angle in specs:
        crop = np.ascontiguousarray(color[y0:y0 + hh, x0:x0 + ww])
        quad = (np.array([[0.08 * ww, 0.06 * hh], [0.93 * ww, 0.09 * hh], [0.95 * ww, 0.94 * hh], [0.05 * ww, 0.9 * hh]])
                + rng.uniform(-6, 6, (4, 2))).astype(np.float32)
        stages, _ = run_reference_stages(crop, quad, p, angle=angle)
        crops[f"{tag}_input"] = crop
        crops[f"{tag}_quad"] = quad
        crops[f"{tag}_angle"] = np.float64(angle)
        crops[f"{tag}_params"] = np.array(json.dumps(p))
        for k, v in stages.items():
            crops[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "crops.npz"), **crops)

    # ---- the reference's own committed artefacts (KAT-1, KAT-2)
    kat = {
        "morphseq_01_gray": cv2.imread(os.path.join(REF, "outputs", "morphseq_01_gray.png"), cv2.IMREAD_UNCHANGED),
        "morphseq_02_eroded": cv2.imread(os.path.join(REF, "outputs", "morphseq_02_eroded.png"), cv2.IMREAD_UNCHANGED),
    }
    for n in ("03_warped", "04_illum", "05_stretch", "05a_inkmask", "06_adapt", "06b_weighted", "07_deskew", "08_clean"):
        a = cv2.imread(os.path.join(REF, "outputs", f"scan_{n}.png"), cv2.IMREAD_UNCHANGED)
        vals = np.unique(a.reshape(-1, a.shape[2]) if a.ndim == 3 else a.reshape(-1, 1), axis=0)
        assert len(vals) == 1, n     # every committed scan_03..08 image is constant
        kat[f"scan_{n}_shape"] = np.array(a.shape)
        kat[f"scan_{n}_value"] = vals[0]
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **kat)

    # ---- Gaussian kernels
    np.savez_compressed(os.path.join(HERE, "gauss_kernels.npz"),
                        **{f"k{k}": cv2.getGaussianKernel(k, 0, cv2.CV_32F).ravel() for k in range(1, 256, 2)})

    # ---- per-op known answers from cv2 on small inputs
    ops = {}
    g = rng.integers(0, 256, (61, 83), dtype=np.uint8)
    page = np.clip(200 - 150 * (rng.random((75, 101)) < 0.12) + rng.normal(0, 4, (75, 101)), 0, 255).astype(np.uint8)
    page = cv2.GaussianBlur(page, (3, 3), 0)
    rgb = rng.integers(0, 256, (37, 41, 3), dtype=np.uint8)
    ops["g"] = g; ops["page"] = page; ops["rgb"] = rgb
    ops["gray_bgr"] = cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY)
    ops["gray_rgb"] = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
    for k in (3, 5, 7, 9, 15, 23, 43, 51, 101):
        ops[f"blur_{k}"] = cv2.GaussianBlur(g, (k, k), 0)
    a = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 256, 1)
    ops["div_table"] = cv2.divide(a, a.T.copy(), scale=255)
    ops["normalize_page"] = cv2.normalize(page, None, 0, 255, cv2.NORM_MINMAX)
    ops["otsu_page"] = np.float64(cv2.threshold(page, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[0])
    ops["otsu_g"] = np.float64(cv2.threshold(g, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[0])
    for kw, kh in ((2, 2), (3, 3), (9, 19), (4, 6), (31, 31), (101, 5)):
        se = cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kh))
        ops[f"erode_{kw}x{kh}"] = cv2.erode(g, se)
        ops[f"dilate_{kw}x{kh}"] = cv2.dilate(g, se)
        ops[f"close_{kw}x{kh}_it2"] = cv2.morphologyEx(g, cv2.MORPH_CLOSE, se, iterations=2)
    ops["blackhat_9x19"] = cv2.morphologyEx(page, cv2.MORPH_BLACKHAT, cv2.getStructuringElement(cv2.MORPH_RECT, (9, 19)))
    for k, c in ((3, 2), (11, 5), (31, 3), (35, 10)):
        ops[f"adapt_gauss_{k}_{c}"] = cv2.adaptiveThreshold(page, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, k, c)
        ops[f"adapt_mean_{k}_{c}"] = cv2.adaptiveThreshold(page, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, k, c)
    for ang in (0.0, 0.5, -3.0, 9.5):
        m = cv2.getRotationMatrix2D((page.shape[1] / 2.0, page.shape[0] / 2.0), ang, 1.0)
        ops[f"rot_{ang}"] = cv2.warpAffine(page, m, (page.shape[1], page.shape[0]), flags=cv2.INTER_LINEAR,
                                           borderMode=cv2.BORDER_REPLICATE)
        ops[f"rotm_{ang}"] = m
    quad = np.array([[3.5, 2.25], [38.0, 4.0], [36.5, 33.0], [1.0, 35.5]], np.float32)
    dst = np.array([[0, 0], [69, 0], [69, 98], [0, 98]], np.float32)
    m = cv2.getPerspectiveTransform(quad, dst)
    ops["persp_quad"] = quad; ops["persp_dst"] = dst; ops["persp_m"] = m
    ops["persp_out"] = cv2.warpPerspective(rgb, m, (70, 99), flags=cv2.INTER_LINEAR)
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **ops)

    for fn in sorted(os.listdir(HERE)):
        print(f"{os.path.getsize(os.path.join(HERE, fn)):>9}  {fn}")


if __name__ == "__main__":
    main()
