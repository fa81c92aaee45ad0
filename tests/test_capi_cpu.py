"""CPU: the C-ABI library loads, exports every symbol include/docscan.h declares, fails loudly without a GPU,
and its host-side parameter preparation agrees with the oracle bit for bit (no kernel is launched here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_npz
from oracle import oracle as O
from smart_image_processing_b200 import _capi, ops
from smart_image_processing_b200 import DocScanner as DS


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "docscan.h")).read()
    declared = set(re.findall(r"^DOCSCAN_API\s+[\w\s\*]+?\b(docscan_\w+)\(", header, flags=re.M))
    assert len(declared) >= 30
    lib = _capi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    assert lib.docscan_version() >= 100


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_capi.DocscanError):
        _capi.Context(0)
    with pytest.raises(_capi.DocscanError):
        ops.gaussian_blur(np.zeros((8, 8), np.uint8), 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "smart_image_processing_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "from oracle" not in text and "import oracle" not in text and "libdocscan_oracle" not in text, fn


def test_gaussian_kernels_match_golden_and_oracle():
    kern = load_npz("gauss_kernels.npz")
    lib = _capi.lib()
    for k in range(1, 256, 2):
        g = np.zeros(k, np.float32)
        q = np.zeros(k, np.int32)
        assert lib.docscan_gaussian_kernel_f32(k, g.ctypes.data_as(C.POINTER(C.c_float))) == 0
        assert lib.docscan_gaussian_kernel_q8(k, q.ctypes.data_as(C.POINTER(C.c_int32))) == 0
        assert np.array_equal(g, kern[f"k{k}"]), k
        assert np.array_equal(q, O.gaussian_kernel_q8(k)), k


def test_matrices_match_oracle_bitwise():
    rng = np.random.default_rng(3)
    for _ in range(300):
        W, H = rng.integers(200, 5000, 2)
        quad = (np.array([[0.1 * W, 0.07 * H], [0.9 * W, 0.09 * H], [0.93 * W, 0.93 * H], [0.07 * W, 0.91 * H]])
                + rng.uniform(-60, 60, (4, 2))).astype(np.float32)
        tw, th = rng.integers(50, 4000, 2)
        dst = np.array([[0, 0], [tw - 1, 0], [tw - 1, th - 1], [0, th - 1]], np.float32)
        assert np.array_equal(ops.get_perspective_transform(quad, dst), O.get_perspective_transform(quad, dst))
        ang = float(rng.integers(-20, 21)) * 0.5
        assert np.array_equal(ops.get_rotation_matrix((W / 2.0, H / 2.0), ang), O.rotation_matrix((W / 2.0, H / 2.0), ang))
        assert DS.target_size(quad, "A4", 1600) == O.target_size(quad, "A4", 1600)
        assert DS.target_size(quad, "custom", 1234) == O.target_size(quad, "custom", 1234)
    ops_npz = load_npz("ops.npz")
    assert np.array_equal(ops.get_perspective_transform(ops_npz["persp_quad"], ops_npz["persp_dst"]), ops_npz["persp_m"])


def test_otsu_from_hist_matches_oracle():
    rng = np.random.default_rng(4)
    lib = _capi.lib()
    for _ in range(50):
        im = np.clip(rng.normal(rng.uniform(30, 220), rng.uniform(2, 70), (64, 64)), 0, 255).astype(np.uint8)
        hist = np.bincount(im.ravel(), minlength=256).astype(np.int32)
        t = C.c_double()
        assert lib.docscan_otsu_from_hist(hist.ctypes.data_as(C.POINTER(C.c_int32)), im.size, C.byref(t)) == 0
        assert t.value == O.otsu_threshold(im)


def test_default_params_are_the_cli_defaults():
    p = _capi.Params()
    _capi.lib().docscan_default_params(C.byref(p))
    assert (p.illum_method, p.block_size, p.C, p.thresh_method, p.mask_blur_ksize, p.blackhat_ksize) == (0, 35, 10, 1, 51, 9)
    assert (p.ink_dilate_iters, p.mask_thresh_offset, p.morph_ksize, p.morph_iters) == (1, 8, 3, 1)
    assert p.illum_blur_frac == 0.02 and p.blackhat_vertical_ratio == 2.0
    q = DS.make_params()
    for name, _ in _capi.Params._fields_:
        assert getattr(p, name) == getattr(q, name), name


def test_signatures_mirror_the_reference():
    import inspect
    sig = {n: list(inspect.signature(getattr(DS, n)).parameters) for n in
           ("perspective_warp", "illumination_correction", "adaptive_binarize", "contrast_stretch", "_compute_ink_mask",
            "morph_cleanup")}
    assert sig["perspective_warp"] == ["img", "quad", "page", "scale_long"]
    assert sig["illumination_correction"] == ["gray", "method", "blur_frac"]
    assert sig["adaptive_binarize"] == ["gray", "block_size", "C", "method"]
    assert sig["_compute_ink_mask"] == ["gray", "mask_blur_ksize", "blackhat_ksize", "blackhat_vertical_ratio", "dilate_iters", "threshold_offset"]
    assert sig["morph_cleanup"] == ["bin_img", "ksize", "iterations"]
    pd = inspect.signature(DS.process_document).parameters
    names = list(pd)[:28]
    assert names[:5] == ["input_path", "out_dir", "page", "scale_long", "do_ocr"] and names[-1] == "min_quad_area_ratio"
    assert pd["mask_blur_ksize"].default == 51 and inspect.signature(DS._compute_ink_mask).parameters["mask_blur_ksize"].default == 61
    assert list(inspect.signature(DS.deskew).parameters)[:4] == ["gray", "canny_low", "canny_high", "max_rotate"]


def test_warp_footprint_bounds_every_source_pixel_the_warp_reads():
    """The host pipeline uploads only docscan_warp_footprint's box of each photo.  For random quads (inside, touching and
    partly outside the photo, strong perspective) every bilinear tap the warp takes — computed here with the oracle's
    matrices in float64 — must lie inside the box, with the margin the kernel's 16-byte load window needs."""
    import ctypes as C
    from oracle import oracle as O
    rng = np.random.default_rng(99)
    lib = _capi.lib()
    for t in range(300):
        W, H = int(rng.integers(40, 900)), int(rng.integers(40, 900))
        spread = [0.05, 0.2, 0.6][t % 3] * min(W, H)
        quad = (np.array([[0.1 * W, 0.1 * H], [0.9 * W, 0.1 * H], [0.9 * W, 0.9 * H], [0.1 * W, 0.9 * H]])
                + rng.uniform(-spread, spread, (4, 2))).astype(np.float32)
        tw, th = int(rng.integers(2, 400)), int(rng.integers(2, 400))
        page = _capi.Page()
        page.src = _capi.Image(1, W, H, W * 3, 3, _capi.HOST)
        page.warped = _capi.Image(1, tw, th, tw * 3, 3, _capi.HOST)
        page.binary = _capi.Image(1, tw, th, tw, 1, _capi.HOST)
        page.quad = (C.c_float * 8)(*quad.reshape(8).tolist())
        reg = (C.c_int32 * 4)()
        assert lib.docscan_warp_footprint(C.byref(page), reg) == 0
        x0, y0, x1, y1 = list(reg)
        assert 0 <= x0 < x1 <= W and 0 <= y0 < y1 <= H
        dst = np.array([[0, 0], [tw - 1, 0], [tw - 1, th - 1], [0, th - 1]], np.float32)
        m = O.get_perspective_transform(quad, dst)
        inv = np.linalg.inv(m)
        xs, ys = np.meshgrid(np.arange(tw, dtype=np.float64), np.arange(th, dtype=np.float64))
        wv = inv[2, 0] * xs + inv[2, 1] * ys + inv[2, 2]
        if (x0, y0, x1, y1) == (0, 0, W, H):
            continue                                         # whole photo: nothing to prove
        assert (wv > 0).all() or (wv < 0).all()
        sx = np.floor((inv[0, 0] * xs + inv[0, 1] * ys + inv[0, 2]) / wv + 1.0 / 64).astype(np.int64)
        sy = np.floor((inv[1, 0] * xs + inv[1, 1] * ys + inv[1, 2]) / wv + 1.0 / 64).astype(np.int64)
        # taps (sx, sx + 1) x (sy, sy + 1) that fall inside the photo must be resident
        for dx in (0, 1):
            for dy in (0, 1):
                tx, ty = sx + dx, sy + dy
                inside = (tx >= 0) & (tx < W) & (ty >= 0) & (ty < H)
                assert ((tx[inside] >= x0) & (tx[inside] < x1) & (ty[inside] >= y0) & (ty[inside] < y1)).all(), (t, quad.tolist())
        # interior pixels keep the kernel's fast path: 3 px of slack on the left, 5 on the right, unless the photo ends there
        inside = (sx >= 0) & (sx + 1 < W) & (sy >= 0) & (sy + 1 < H)
        if inside.any():
            assert x0 == 0 or sx[inside].min() - x0 >= 3
            assert x1 == W or x1 - sx[inside].max() >= 6


def test_header_is_plain_c_and_library_links_from_c(tmp_path):
    """include/docscan.h must compile as C99 and libdocscan.so must be usable without Python or C++."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "smart_image_processing_b200")
    exe = str(tmp_path / "abi_check")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "c_abi", "abi_check.c"), "-o", exe, "-L", libdir, "-l:libdocscan.so",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert "version 100" in out and "block 35 canny 50/150" in out
    import torch
    if not torch.cuda.is_available():
        assert "create rc -1" in out          # no device: the library refuses, it has no CPU fallback
