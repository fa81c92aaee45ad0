"""oracle/ref_cv2.py — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's per-pixel path restated with the same OpenCV calls the reference makes, in the same order
(DocScanner.py:117-259, :316, :338-339).  The reference itself cannot travel to the GPU box (/root/reference
does not exist there) and is nothing but these cv2 calls, so this port — running on the very library the
reference depends on (opencv-python, requirements.txt:1) — is what bench.py times as the CPU baseline
(`cpu_baseline.kind = "port"`) and as `--impl reference`.  tests/test_oracle_vs_cv2.py and
tests/golden/make_golden.py tie it to the C oracle and to the reference's own functions.
"""
from __future__ import annotations

import math

import numpy as np

try:
    import cv2
    HAVE_CV2 = True
except ImportError:  # pragma: no cover
    cv2 = None
    HAVE_CV2 = False


def hot_path(color, quad, angle_deg, *, page="A4", scale_long=1600, illum_method="subtract", illum_blur_frac=0.02,
             block_size=35, C=10, thresh_method="gaussian", mask_blur_ksize=51, blackhat_ksize=9,
             blackhat_vertical_ratio=2.0, ink_dilate_iters=1, mask_thresh_offset=8, morph_ksize=3, morph_iters=1,
             keep_stages=False, canny_low=50, canny_high=150, max_rotate=10.0):
    quad = np.asarray(quad, np.float32)
    # --- perspective_warp (DocScanner.py:117-144)
    tl, tr, br, bl = quad
    width = max(int(np.linalg.norm(tr - tl)), int(np.linalg.norm(br - bl)))
    height = max(int(np.linalg.norm(bl - tl)), int(np.linalg.norm(br - tr)))
    if page.upper() in ("A4", "A3", "A5", "LETTER"):
        ratio = math.sqrt(2.0) if page.upper() != "LETTER" else (11.0 / 8.5)
    else:
        ratio = height / max(width, 1)
    if height >= width:
        th, tw = scale_long, int(round(scale_long / ratio))
    else:
        tw, th = scale_long, int(round(scale_long * ratio))
    dst = np.array([[0, 0], [tw - 1, 0], [tw - 1, th - 1], [0, th - 1]], dtype=np.float32)
    warped = cv2.warpPerspective(color, cv2.getPerspectiveTransform(quad, dst), (tw, th), flags=cv2.INTER_LINEAR)
    gray = cv2.cvtColor(warped, cv2.COLOR_BGR2GRAY)                                     # :316
    # --- illumination_correction (:147-160)
    base = max(15, int(round(min(gray.shape[:2]) * illum_blur_frac)))
    base += 1 - base % 2
    bg = cv2.GaussianBlur(gray, (base, base), 0)
    tmp = cv2.divide(gray, bg, scale=255) if illum_method.lower() == "divide" else cv2.subtract(gray, bg)
    illum = cv2.normalize(tmp, None, 0, 255, cv2.NORM_MINMAX)
    stretched = cv2.normalize(illum, None, alpha=0, beta=255, norm_type=cv2.NORM_MINMAX)   # :171-172
    # --- _compute_ink_mask (:175-214)
    mk = mask_blur_ksize + 1 - mask_blur_ksize % 2
    ink_sub = cv2.normalize(cv2.subtract(cv2.GaussianBlur(stretched, (mk, mk), 0), stretched), None, 0, 255, cv2.NORM_MINMAX)
    t_sub, _ = cv2.threshold(ink_sub, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    _, mask_sub = cv2.threshold(ink_sub, max(0, int(round(t_sub - mask_thresh_offset))), 255, cv2.THRESH_BINARY)
    bk = max(3, blackhat_ksize)
    bk += 1 - bk % 2
    bh_h = max(3, int(round(bk * blackhat_vertical_ratio)))
    bh_h += 1 - bh_h % 2
    bh = cv2.morphologyEx(stretched, cv2.MORPH_BLACKHAT, cv2.getStructuringElement(cv2.MORPH_RECT, (bk, bh_h)))
    bh = cv2.normalize(bh, None, 0, 255, cv2.NORM_MINMAX)
    t_bh, _ = cv2.threshold(bh, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    _, mask_bh = cv2.threshold(bh, max(0, int(round(t_bh - mask_thresh_offset))), 255, cv2.THRESH_BINARY)
    ink = cv2.max(mask_sub, mask_bh)
    if ink_dilate_iters > 0:
        ink = cv2.dilate(ink, cv2.getStructuringElement(cv2.MORPH_RECT, (2, 2)), iterations=ink_dilate_iters)
    # --- adaptive_binarize (:163-168)
    blk = block_size + 1 - block_size % 2
    algo = cv2.ADAPTIVE_THRESH_GAUSSIAN_C if thresh_method.lower() == "gaussian" else cv2.ADAPTIVE_THRESH_MEAN_C
    adapt = cv2.adaptiveThreshold(stretched, 255, algo, cv2.THRESH_BINARY, blk, C)
    weighted = adapt.copy()                                                              # :338-339
    weighted[ink == 0] = 255
    # --- deskew (:217-236): the rotation, with the angle from the control path or, for angle_deg=None / NaN, the reference's own
    # estimate (:218-231: Canny, HoughLines(1, pi/180, 150), median of the folded line angles, 0 beyond max_rotate)
    if angle_deg is None or angle_deg != angle_deg:
        angle_deg = 0.0
        lines = cv2.HoughLines(cv2.Canny(weighted, canny_low, canny_high), 1, np.pi / 180, 150)
        if lines is not None and len(lines) > 0:
            angles = [(theta * 180.0 / np.pi + 90.0) % 180.0 - 90.0 for _, theta in lines[:, 0, :]]
            angle_deg = float(np.median(angles))
            if abs(angle_deg) > max_rotate:
                angle_deg = 0.0
    h, w = weighted.shape[:2]
    rot = cv2.warpAffine(weighted, cv2.getRotationMatrix2D((w / 2.0, h / 2.0), angle_deg, 1.0), (w, h),
                         flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
    # --- morph_cleanup (:247-259)
    clean = rot if morph_ksize <= 1 else cv2.morphologyEx(
        rot, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (morph_ksize, morph_ksize)), iterations=morph_iters)
    if keep_stages:
        return dict(warped=warped, gray=gray, illum=illum, stretch=stretched, inkmask=ink, adapt=adapt,
                    weighted=weighted, deskew=rot, clean=clean)
    return warped, clean
