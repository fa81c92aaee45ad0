"""Drop-in for the per-pixel stage functions of the reference's DocScanner.py.

Same names, positional order, defaults, numpy-uint8-in / numpy-uint8-out contract and ownership rules as
`/root/reference/DocScanner.py:117-259` and `process_document` (`:262-365`), but every pixel operation
runs as a CUDA kernel of libdocscan.so (B200, sm_100a) through the C ABI in include/docscan.h.

    import smart_image_processing_b200.DocScanner as DS     # instead of: import DocScanner as DS

What is NOT here: the reference's quad detection (localize_document), file I/O and OCR.  `process_document` calls
them through `control.py`, which uses OpenCV on the host when it is installed; pass `quad=` to skip it.  The skew
estimate of deskew() (Canny + HoughLines median) runs on the device.  There is no CPU fallback for the pixel path.
"""
from __future__ import annotations

import ctypes as _ct
import math
import os
from typing import Optional, Sequence

import numpy as np

from . import _capi, ops
from ._capi import Page, Params, image_of

__all__ = [
    "resize_long_side", "perspective_warp", "illumination_correction", "adaptive_binarize", "contrast_stretch", "_compute_ink_mask",
    "deskew", "rotate", "morph_cleanup", "process_document", "process_pages", "hot_path", "target_size",
]


def _ctx(ctx=None):
    return ctx if ctx is not None else _capi.default_context()


def _a_series_ratio() -> float:
    return math.sqrt(2.0)


def target_size(quad: np.ndarray, page: str = "A4", scale_long: int = 1600):
    """Size of the rectified page: DocScanner.py:120-139 (same numpy calls, so the same roundings)."""
    (tl, tr, br, bl) = quad
    w_top = np.linalg.norm(tr - tl)
    w_bottom = np.linalg.norm(br - bl)
    h_left = np.linalg.norm(bl - tl)
    h_right = np.linalg.norm(br - tr)
    width = max(int(w_top), int(w_bottom))
    height = max(int(h_left), int(h_right))
    portrait = height >= width
    if page.upper() in ("A4", "A3", "A5", "LETTER"):
        ratio = _a_series_ratio() if page.upper() != "LETTER" else (11.0 / 8.5)
    else:
        ratio = height / max(width, 1)
    if portrait:
        target_h = scale_long
        target_w = int(round(target_h / ratio))
    else:
        target_w = scale_long
        target_h = int(round(target_w * ratio))
    return target_w, target_h


def resize_long_side(img: np.ndarray, scale_long: int) -> np.ndarray:
    """DocScanner.py:27-36: the whole-photo fallback (INTER_AREA when the long side shrinks, INTER_CUBIC otherwise;
    returns its argument when scale_long <= 0, like the reference)."""
    h, w = img.shape[:2]
    if scale_long <= 0:
        return img
    long = max(h, w)
    sf = scale_long / float(long)
    new_w = int(round(w * sf))
    new_h = int(round(h * sf))
    return ops.resize(img, (new_w, new_h), _capi.INTER_AREA if sf < 1.0 else _capi.INTER_CUBIC)


def perspective_warp(img: np.ndarray, quad: np.ndarray, page: str = "A4", scale_long: int = 1600) -> np.ndarray:
    """DocScanner.py:117-144."""
    quad = np.asarray(quad)
    target_w, target_h = target_size(quad, page, scale_long)
    dst = np.array([[0, 0], [target_w - 1, 0], [target_w - 1, target_h - 1], [0, target_h - 1]], dtype=np.float32)
    m = ops.get_perspective_transform(quad.astype(np.float32), dst)
    return ops.warp_perspective(img, m, (target_w, target_h))


def _illum_ksize(h: int, w: int, blur_frac: float) -> int:
    base = max(15, int(round(min(h, w) * blur_frac)))        # DocScanner.py:150-152
    if base % 2 == 0:
        base += 1
    return base


def illumination_correction(gray: np.ndarray, method: str = "subtract", blur_frac: float = 0.02) -> np.ndarray:
    """DocScanner.py:147-160: large Gaussian background, subtract or divide, MINMAX normalise — one fused
    blur+epilogue+min/max kernel, one LUT pass."""
    gray = np.ascontiguousarray(gray)
    h, w = gray.shape[:2]
    out = np.empty_like(gray)
    s, d = image_of(gray), image_of(out)
    _ctx().call("docscan_illumination_correction", _ct.byref(s), 1 if method.lower() == "divide" else 0,
                _illum_ksize(h, w, blur_frac), _ct.byref(d))
    return out


def adaptive_binarize(gray: np.ndarray, block_size: int = 35, C: int = 10, method: str = "gaussian") -> np.ndarray:
    """DocScanner.py:163-168."""
    if block_size % 2 == 0:
        block_size += 1
    return ops.adaptive_threshold(gray, "gaussian" if method.lower() == "gaussian" else "mean", block_size, C)


def contrast_stretch(gray: np.ndarray) -> np.ndarray:
    """DocScanner.py:171-172."""
    return ops.normalize_minmax(gray)


def _compute_ink_mask(gray: np.ndarray, mask_blur_ksize: int = 61, blackhat_ksize: int = 9,
                      blackhat_vertical_ratio: float = 2.0, dilate_iters: int = 1,
                      threshold_offset: int = 8) -> np.ndarray:
    """DocScanner.py:175-214 as one fused sequence: blur+subtract+histogram, black-hat+histogram, device-side
    normalise/Otsu scalars, combine + 2x2 dilate."""
    if mask_blur_ksize % 2 == 0:
        mask_blur_ksize += 1
    if blackhat_ksize < 3:
        blackhat_ksize = 3
    if blackhat_ksize % 2 == 0:
        blackhat_ksize += 1
    bh_h = max(3, int(round(blackhat_ksize * blackhat_vertical_ratio)))
    if bh_h % 2 == 0:
        bh_h += 1
    gray = np.ascontiguousarray(gray)
    out = np.empty_like(gray)
    s, d = image_of(gray), image_of(out)
    _ctx().call("docscan_ink_mask", _ct.byref(s), int(mask_blur_ksize), int(blackhat_ksize), int(bh_h),
                int(dilate_iters), int(threshold_offset), _ct.byref(d))
    return out


def rotate(gray: np.ndarray, angle_deg: float) -> np.ndarray:
    """The pixel half of deskew (DocScanner.py:233-236): rotate about the centre, bilinear, replicate border."""
    h, w = gray.shape[:2]
    m = ops.get_rotation_matrix((w / 2.0, h / 2.0), angle_deg)
    return ops.warp_affine(gray, m, (w, h))


def deskew(gray: np.ndarray, canny_low: int = 50, canny_high: int = 150, max_rotate: float = 10.0,
           angle: Optional[float] = None) -> np.ndarray:
    """DocScanner.py:217-236: Canny + HoughLines median skew estimate and the rotation, both on the device;
    pass `angle=` to supply the angle instead."""
    if angle is None:
        angle = ops.skew_angle(gray, canny_low, canny_high, max_rotate)
    return rotate(gray, angle)


def morph_cleanup(bin_img: np.ndarray, ksize: int = 3, iterations: int = 1) -> np.ndarray:
    """DocScanner.py:247-259 (returns its argument when ksize <= 1, like the reference)."""
    if ksize <= 1:
        return bin_img
    return ops.morph_close(bin_img, ksize, ksize, iterations)


# ------------------------------------------------------------------------------------------------- batched path

_PIXEL_KEYS = ("illum_method", "illum_blur_frac", "block_size", "C", "thresh_method", "mask_blur_ksize",
               "blackhat_ksize", "blackhat_vertical_ratio", "ink_dilate_iters", "mask_thresh_offset",
               "morph_ksize", "morph_iters")


def make_params(illum_method="subtract", illum_blur_frac=0.02, block_size=35, C=10, thresh_method="gaussian",
                mask_blur_ksize=51, blackhat_ksize=9, blackhat_vertical_ratio=2.0, ink_dilate_iters=1,
                mask_thresh_offset=8, morph_ksize=3, morph_iters=1, cv_tail_compat=True,
                canny_low=50, canny_high=150, max_rotate=10.0) -> Params:
    """process_document's pixel tunables (DocScanner.py:268-273) as a docscan_params struct."""
    p = Params()
    p.illum_method = 1 if str(illum_method).lower() == "divide" else 0
    p.illum_blur_frac = float(illum_blur_frac)
    p.block_size = int(block_size)
    p.C = int(math.ceil(C))
    p.thresh_method = _capi.ADAPTIVE_GAUSSIAN if str(thresh_method).lower() == "gaussian" else _capi.ADAPTIVE_MEAN
    p.mask_blur_ksize = int(mask_blur_ksize)
    p.blackhat_ksize = int(blackhat_ksize)
    p.blackhat_vertical_ratio = float(blackhat_vertical_ratio)
    p.ink_dilate_iters = int(ink_dilate_iters)
    p.mask_thresh_offset = int(mask_thresh_offset)
    p.morph_ksize = int(morph_ksize)
    p.morph_iters = int(morph_iters)
    p.cv_tail_compat = int(bool(cv_tail_compat))
    p.canny_low, p.canny_high, p.max_rotate = float(canny_low), float(canny_high), float(max_rotate)
    return p


def process_pages(images: Sequence[np.ndarray], quads: Sequence[np.ndarray], angles: Sequence[float], *,
                  page: str = "A4", scale_long: int = 1600, ctx=None, out_warped=None, out_binary=None,
                  return_angles: bool = False, **tunables):
    """The per-pixel part of process_document (DocScanner.py:310-346) for a batch of independent pages in one
    C-ABI call: warp -> gray -> illumination -> stretch -> ink mask || adaptive threshold -> blend -> rotate ->
    close.  `images` are HxWx3 uint8 BGR numpy arrays (host); returns (warped list, binary list).
    A quad of None sends that page through the whole-photo fallback (resize_long_side) instead of the warp; an
    angle of None makes the library estimate the skew on the device exactly like deskew() (canny_low, canny_high,
    max_rotate tunables).  return_angles=True appends the list of angles the pages were rotated by.
    `out_warped` / `out_binary` may hold preallocated (e.g. pinned) arrays of the right shapes."""
    ctx = _ctx(ctx)
    n = len(images)
    params = make_params(**tunables)
    pages = (Page * n)()
    warped, binary, keep = [], [], []
    for i in range(n):
        dev_img = images[i] if isinstance(images[i], _capi.DeviceBuffer) else None      # photo already on the device
        img = images[i] if dev_img is not None else np.ascontiguousarray(images[i])
        if dev_img is None and (img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3):
            raise TypeError("process_pages: images must be HxWx3 uint8 (BGR)")
        if dev_img is not None and (len(img.shape) != 3 or img.shape[2] != 3):
            raise TypeError("process_pages: device photos must be HxWx3 (BGR)")
        whole = quads[i] is None                       # no usable quad: resize_long_side (DocScanner.py:313)
        if whole:
            if scale_long <= 0:
                raise ValueError("process_pages: scale_long must be positive for whole-photo pages")
            sf = scale_long / float(max(img.shape[:2]))
            tw, th = int(round(img.shape[1] * sf)), int(round(img.shape[0] * sf))
            q = np.zeros((4, 2), np.float32)
        else:
            q = np.asarray(quads[i], np.float32).reshape(4, 2)
            tw, th = target_size(q, page, scale_long)
        w_arr = out_warped[i] if out_warped is not None else np.empty((th, tw, 3), np.uint8)
        b_arr = out_binary[i] if out_binary is not None else np.empty((th, tw), np.uint8)
        if w_arr.shape != (th, tw, 3) or b_arr.shape != (th, tw):
            raise ValueError("process_pages: preallocated outputs have the wrong shape")
        pages[i].src = dev_img.image() if dev_img is not None else image_of(img)
        pages[i].quad = (_ct.c_float * 8)(*q.reshape(8).tolist())
        pages[i].angle_deg = float("nan") if angles[i] is None else float(angles[i])
        pages[i].warped = image_of(w_arr)
        pages[i].binary = image_of(b_arr)
        pages[i].use_whole = int(whole)
        keep.append(img)
        warped.append(w_arr)
        binary.append(b_arr)
    ctx.call("docscan_process_pages", n, pages, _ct.byref(params))
    if return_angles:
        out = (_ct.c_double * max(n, 1))()
        ctx.call("docscan_last_angles", out, n)
        return warped, binary, [float(out[i]) for i in range(n)]
    return warped, binary


def hot_path(color: np.ndarray, quad: np.ndarray, angle_deg: float, *, page="A4", scale_long=1600, **tunables):
    """Stage-by-stage evaluation through the stage functions above (each a separate C-ABI call); returns every
    intermediate image keyed like the reference's PNG dumps.  Used by the parity tests."""
    t = dict(illum_method="subtract", illum_blur_frac=0.02, block_size=35, C=10, thresh_method="gaussian",
             mask_blur_ksize=51, blackhat_ksize=9, blackhat_vertical_ratio=2.0, ink_dilate_iters=1,
             mask_thresh_offset=8, morph_ksize=3, morph_iters=1)
    t.update(tunables)
    out = {}
    out["warped"] = (perspective_warp(color, quad, page=page, scale_long=scale_long) if quad is not None
                     else resize_long_side(color, scale_long))
    out["gray"] = ops.bgr2gray(out["warped"])
    out["illum"] = illumination_correction(out["gray"], method=t["illum_method"], blur_frac=t["illum_blur_frac"])
    out["stretch"] = contrast_stretch(out["illum"])
    out["inkmask"] = _compute_ink_mask(out["stretch"], mask_blur_ksize=t["mask_blur_ksize"],
                                       blackhat_ksize=t["blackhat_ksize"],
                                       blackhat_vertical_ratio=t["blackhat_vertical_ratio"],
                                       dilate_iters=t["ink_dilate_iters"], threshold_offset=t["mask_thresh_offset"])
    out["adapt"] = adaptive_binarize(out["stretch"], block_size=t["block_size"], C=t["C"], method=t["thresh_method"])
    out["weighted"] = ops.mask_select(out["adapt"], out["inkmask"])
    out["deskew"] = rotate(out["weighted"], angle_deg)
    out["clean"] = morph_cleanup(out["deskew"], ksize=t["morph_ksize"], iterations=t["morph_iters"])
    return out


def process_document(input_path: str, out_dir: str = "outputs", page: str = "A4", scale_long: int = 1600,
                     do_ocr: bool = False,
                     bilateral_d: int = 9, bilateral_sigmaColor: float = 75, bilateral_sigmaSpace: float = 75,
                     gaussian_ksize: int = 0,
                     canny_low: int = 50, canny_high: int = 150,
                     min_area_ratio: float = 0.2, max_area_ratio: float = 0.98,
                     illum_method: str = "subtract", illum_blur_frac: float = 0.02,
                     block_size: int = 35, C: int = 10, thresh_method: str = "gaussian",
                     mask_blur_ksize: int = 51, blackhat_ksize: int = 9,
                     blackhat_vertical_ratio: float = 2.0, ink_dilate_iters: int = 1,
                     mask_thresh_offset: int = 8,
                     morph_ksize: int = 3, morph_iters: int = 1,
                     max_rotate: float = 10.0,
                     fallback_use_whole: bool = True,
                     min_quad_area_ratio: float = 0.15,
                     *, quad: Optional[np.ndarray] = None, angle: Optional[float] = None,
                     save_stages: bool = True, decode: str = "host") -> dict:
    """DocScanner.process_document (DocScanner.py:262-365): same 28 parameters, same result dict
    {"quad", "warped", "binary"}, same twelve PNG dumps in `out_dir` (scan_01_pre ... scan_08_clean).
    The control path (load, bilateral `preprocess` dump, quad localisation, overlay dump) runs on the host through
    control.py unless `quad` is supplied; the per-pixel path and deskew()'s angle estimate run on the GPU.
    Keyword-only extras: `quad=` / `angle=` skip the localisation / the skew estimate; `save_stages=False` drops the
    dumps and takes the single fused C-ABI call (the reference always dumps, hence the default).
    `decode="device"` (needs `quad=` and `save_stages=False`): the JPEG is decoded by nvJPEG straight into device memory, so only
    the file crosses PCIe instead of the raw photo — NOT bit-identical with cv2.imread's libjpeg decode, hence opt-in; "warped"
    and "binary" then differ from the reference's by the decoders' +-1..2 grey levels."""
    from . import control
    if out_dir:                                                # ensure_dir(out_dir), DocScanner.py:277
        os.makedirs(out_dir, exist_ok=True)
    if decode == "device":
        if quad is None or save_stages:
            raise ValueError("decode='device' keeps the photo on the GPU: pass quad= and save_stages=False")
        photo = control.load_image_device(input_path)
        quad = np.asarray(quad, np.float32)
        w, b = process_pages([photo], [quad], [angle], page=page, scale_long=scale_long, canny_low=canny_low, canny_high=canny_high,
                             max_rotate=max_rotate, illum_method=illum_method, illum_blur_frac=illum_blur_frac, block_size=block_size, C=C,
                             thresh_method=thresh_method, mask_blur_ksize=mask_blur_ksize, blackhat_ksize=blackhat_ksize,
                             blackhat_vertical_ratio=blackhat_vertical_ratio, ink_dilate_iters=ink_dilate_iters,
                             mask_thresh_offset=mask_thresh_offset, morph_ksize=morph_ksize, morph_iters=morph_iters)
        return {"quad": quad, "warped": w[0], "binary": b[0]}
    color = control.load_image(input_path)
    if save_stages:
        # DocScanner.py:280-282: the denoised image is only ever dumped, nothing downstream reads it
        control.save_image(os.path.join(out_dir, "scan_01_pre.png"),
                           control.preprocess(color, bilateral_d, bilateral_sigmaColor, bilateral_sigmaSpace, gaussian_ksize))
    use_whole = False
    if quad is None:
        quad = control.localize_document(color, canny_low=canny_low, canny_high=canny_high,
                                         min_area_ratio=min_area_ratio, max_area_ratio=max_area_ratio)
    if quad is None:
        use_whole = True
    else:
        ratio = control.quad_area(quad) / max(color.shape[0] * color.shape[1], 1)
        if ratio < min_quad_area_ratio:
            use_whole = True
    if use_whole and not fallback_use_whole:
        raise RuntimeError("Quad too small or missing, and fallback disabled.")
    if save_stages:
        control.save_image(os.path.join(out_dir, "scan_02_quad.png"), control.quad_overlay(color, None if use_whole else quad))
    tun = dict(illum_method=illum_method, illum_blur_frac=illum_blur_frac, block_size=block_size, C=C,
               thresh_method=thresh_method, mask_blur_ksize=mask_blur_ksize, blackhat_ksize=blackhat_ksize,
               blackhat_vertical_ratio=blackhat_vertical_ratio, ink_dilate_iters=ink_dilate_iters,
               mask_thresh_offset=mask_thresh_offset, morph_ksize=morph_ksize, morph_iters=morph_iters)
    quad = np.asarray(quad, np.float32) if quad is not None else None
    warp_quad = None if use_whole else quad                    # DocScanner.py:310-313
    # resize_long_side hands the photo on unchanged for scale_long <= 0 (DocScanner.py:29-30); the fused call has no
    # such page kind, so that case takes the stage-by-stage route as well
    if save_stages or (use_whole and scale_long <= 0):
        st = hot_path(color, warp_quad, 0.0, page=page, scale_long=scale_long, **tun)
        if angle is None:
            angle = ops.skew_angle(st["weighted"], canny_low, canny_high, max_rotate)     # DocScanner.py:342
        st["deskew"] = rotate(st["weighted"], angle)
        st["clean"] = morph_cleanup(st["deskew"], ksize=morph_ksize, iterations=morph_iters)
        warped, clean = st["warped"], st["clean"]
        if save_stages:
            control.save_stage_dumps(out_dir, st)
    else:
        # one C-ABI call; without a supplied angle the skew estimate runs on the device between blend and rotate
        w, b = process_pages([color], [warp_quad], [angle], page=page, scale_long=scale_long, canny_low=canny_low,
                             canny_high=canny_high, max_rotate=max_rotate, **tun)
        warped, clean = w[0], b[0]
    result = {"quad": quad, "warped": warped, "binary": clean}
    if do_ocr:
        try:
            import pytesseract
            ocr_text = pytesseract.image_to_string(clean, config="--psm 6")
            os.makedirs(out_dir, exist_ok=True)
            with open(os.path.join(out_dir, "scan_ocr.txt"), "w", encoding="utf-8") as f:
                f.write(ocr_text)
            result["ocr_text"] = ocr_text
        except Exception as e:  # same contract as the reference: OCR errors are reported, not raised
            result["ocr_error"] = str(e)
    return result
