"""GPU (-m gpu), second batch: the host-buffer pipeline with pinned memory over many launch groups, the control path
against the quads the reference itself produced, morph_seq end to end, BASELINE configs 4 and 5 at their real sizes,
a second device, and the drop-in's file outputs."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_npz
from oracle import oracle as O
from smart_image_processing_b200 import DocScanner as DS
from smart_image_processing_b200 import _capi, ops
from smart_image_processing_b200 import morph_seq as MS
from test_gpu_parity import eq, page_like, sha

pytestmark = pytest.mark.gpu


def _photo(rng, H, W):
    base = page_like(rng, H, W)
    return np.stack([np.clip(base * s, 0, 255).astype(np.uint8) for s in (0.97, 1.0, 1.02)], -1)


def test_pinned_host_pipeline_many_groups_mixed_pages():
    """26 pages in PINNED host buffers (the asynchronous three-stream pipeline: 8 pages per group, two staging sets) with
    different photo sizes, quads (hence different upload footprints and output sizes) and whole-photo pages: every page
    must equal the result of processing it alone from pageable memory.  Group g reuses the staging set of group g-2, whose
    results may still be on their way back to the host when its uploads start (staging race)."""
    ctx = _capi.Context(0)
    rng = np.random.default_rng(77)
    n = 26
    imgs, quads, angles, outs_w, outs_b = [], [], [], [], []
    for i in range(n):
        H, W = int(rng.integers(300, 900)), int(rng.integers(260, 800))
        if i % 8 in (1, 6):                 # a few much bigger photos so that src slots dwarf the neighbours' outputs
            H, W = H * 2, W * 2
        img = _photo(rng, H, W)
        pin = ctx.pinned_empty(img.shape)
        pin[...] = img
        imgs.append(pin)
        if i % 7 == 3:
            quads.append(None)
        else:
            fx0, fy0, fx1, fy1 = rng.uniform(0.02, 0.3), rng.uniform(0.02, 0.3), rng.uniform(0.7, 0.98), rng.uniform(0.7, 0.98)
            quads.append(np.array([[fx0 * W, fy0 * H], [fx1 * W, fy0 * H + 5], [fx1 * W - 4, fy1 * H], [fx0 * W + 6, fy1 * H - 3]], np.float32))
        angles.append(float(rng.integers(-6, 7)) * 0.5)
    sl = 420
    for i in range(n):
        if quads[i] is None:
            sf = sl / float(max(imgs[i].shape[:2]))
            tw, th = int(round(imgs[i].shape[1] * sf)), int(round(imgs[i].shape[0] * sf))
        else:
            tw, th = DS.target_size(quads[i], "A4", sl)
        outs_w.append(ctx.pinned_empty((th, tw, 3)))
        outs_b.append(ctx.pinned_empty((th, tw)))
    for rep in range(2):                    # twice: the second call starts with both staging sets dirty
        for a in outs_w + outs_b:
            a[...] = 0x5A
        w, b = DS.process_pages(imgs, quads, angles, scale_long=sl, ctx=ctx, out_warped=outs_w, out_binary=outs_b)
        for i in range(n):
            w1, b1 = DS.process_pages([np.array(imgs[i])], [quads[i]], [angles[i]], scale_long=sl, ctx=ctx)
            eq(w[i], w1[0], f"pinned batch rep {rep} page {i} warped")
            eq(b[i], b1[0], f"pinned batch rep {rep} page {i} binary")
    # the same batch from PAGEABLE arrays (what cv2.imread hands a caller): pinned mirror + host copy threads inside the library
    wp, bp = DS.process_pages([np.array(x) for x in imgs], quads, angles, scale_long=sl, ctx=ctx)
    for i in range(n):
        eq(wp[i], np.array(w[i]), f"pageable batch page {i} warped")
        eq(bp[i], np.array(b[i]), f"pageable batch page {i} binary")
    # and one of them against the oracle, so that "equal to itself" cannot hide a common error
    ref = O.hot_path(np.array(imgs[0]), quads[0], angles[0], scale_long=sl)
    eq(np.array(outs_b[0]), ref["clean"], "pinned batch page 0 vs oracle")
    del imgs, outs_w, outs_b, w, b
    ctx.close()


@pytest.mark.parametrize("preset", ["cli", "gui"])
def test_localize_document_matches_reference_quads(preset):
    """control.localize_document (gray + Canny on the device, HoughLinesP / contours / polygon fit with cv2 on the host)
    against the quad the REFERENCE's localize_document returned for public/sample.jpg (tests/golden/make_golden.py),
    bit for bit, and the area gate against cv2.contourArea."""
    cv2 = pytest.importorskip("cv2")
    from smart_image_processing_b200 import control
    gold = json.load(open(os.path.join(GOLDEN, "sample_golden.json")))["presets"][preset]
    img = load_npz("sample_bgr.npz")["bgr"]
    want = np.frombuffer(bytes.fromhex(gold["quad_f32_hex"]), np.float32).reshape(4, 2)
    p = gold["params"]
    quad = control.localize_document(img, canny_low=p["canny_low"], canny_high=p["canny_high"])
    assert quad is not None and quad.dtype == np.float32
    assert quad.tobytes() == want.tobytes(), f"{preset}: {quad.tolist()} vs {want.tolist()}"
    assert control.quad_area(quad) == float(cv2.contourArea(quad.astype(np.float32).reshape(-1, 1, 2)))


def test_process_morph_seq_end_to_end(tmp_path):
    """morph_seq.process_morph_seq: file in, five keyed images out, four PNG dumps; every step against the oracle."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    img = _photo(rng, 312, 406)
    path = str(tmp_path / "in.png")
    cv2.imwrite(path, img)
    out = MS.process_morph_seq(path, out_dir=str(tmp_path / "o"))
    assert list(out) == ["original", "step1_gray", "step2_eroded", "step3_otsu", "step4_closed"]
    rgb = img[:, :, ::-1]
    eq(out["original"], np.ascontiguousarray(rgb), "morph_seq original (RGB)")
    gray = O.to_grayscale(np.ascontiguousarray(rgb))
    eq(out["step1_gray"], gray, "morph_seq gray")
    eq(out["step2_eroded"], O.grayscale_erosion(gray), "morph_seq eroded")
    eq(out["step3_otsu"], O.otsu_binarize(O.grayscale_erosion(gray)), "morph_seq otsu")
    eq(out["step4_closed"], O.binary_closing(O.otsu_binarize(O.grayscale_erosion(gray))), "morph_seq closed")
    for name, key in (("morphseq_01_gray.png", "step1_gray"), ("morphseq_02_eroded.png", "step2_eroded"),
                      ("morphseq_03_otsu.png", "step3_otsu"), ("morphseq_04_closed.png", "step4_closed")):
        eq(cv2.imread(str(tmp_path / "o" / name), cv2.IMREAD_GRAYSCALE), out[key], name)


def _gray_scan(seed, h, w):
    """SURVEY 8(d) gray-only generator: illumination gradient, 25 % ink blocks of 16x4 px, N(0,3) noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = 235.0 * (0.55 + 0.45 * (0.6 * xx / w + 0.4 * yy / h))
    cells = rng.random((h // 4 + 1, w // 16 + 1)) < 0.25
    ink = np.kron(cells, np.ones((4, 16), bool))[:h, :w]
    img = np.where(ink, img * 0.25, img) + rng.normal(0, 3, (h, w)).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def test_config4_full_size_4k():
    """BASELINE config 4 at its real size (3840x2160 gray): erode / dilate / close and both adaptive thresholds against
    cv2 itself (the reference's arithmetic; the C oracle would take minutes here) for the smallest, a middle and the
    largest kernel of the sweep, plus morph_seq's 2x2."""
    cv2 = pytest.importorskip("cv2")
    g = _gray_scan(4, 2160, 3840)
    for k in (2, 3, 17, 31):
        se = cv2.getStructuringElement(cv2.MORPH_RECT, (k, k))
        eq(ops.erode(g, k, k, 1), cv2.erode(g, se), f"4K erode {k}")
        eq(ops.dilate(g, k, k, 1), cv2.dilate(g, se), f"4K dilate {k}")
        eq(ops.morph_close(g, k, k, 1), cv2.morphologyEx(g, cv2.MORPH_CLOSE, se), f"4K close {k}")
    for k in (3, 17, 31):
        eq(ops.adaptive_threshold(g, "gaussian", k, 10), cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, k, 10), f"4K adaptive gauss {k}")
        eq(ops.adaptive_threshold(g, "mean", k, 10), cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, k, 10), f"4K adaptive mean {k}")


@pytest.mark.parametrize("k", [101, 217])
def test_config5_full_size_8k(k):
    """BASELINE config 5 at its real size (7680x4320 gray): illumination_correction(divide) with a >= 101 px Gaussian and the
    close+divide+normalize variant, against the cv2 composition (60 strips, several vertical segments per strip)."""
    cv2 = pytest.importorskip("cv2")
    g = _gray_scan(5, 4320, 7680)
    frac = (k - 0.4) / 4320.0
    assert DS._illum_ksize(4320, 7680, frac) == k
    bg = cv2.GaussianBlur(g, (k, k), 0)
    want = cv2.normalize(cv2.divide(g, bg, scale=255), None, 0, 255, cv2.NORM_MINMAX)
    eq(DS.illumination_correction(g, "divide", frac), want, f"8K illumination divide k={k}")
    if k == 101:
        se = cv2.getStructuringElement(cv2.MORPH_RECT, (k, k))
        bgc = cv2.morphologyEx(g, cv2.MORPH_CLOSE, se)
        want = cv2.normalize(cv2.divide(g, bgc, scale=255), None, 0, 255, cv2.NORM_MINMAX)
        got = ops.normalize_minmax(ops.divide255(g, ops.morph_close(g, k, k, 1)))
        eq(got, want, f"8K close+divide+normalize k={k}")


def _device_count():
    import torch
    return torch.cuda.device_count()


def test_same_pages_on_a_second_device():
    """SURVEY 4.3 / BASELINE config 3: page i must give the same bytes on GPU k as on GPU 0.  Needs 2 GPUs (skipped on
    the one-GPU test box; run with `gpurun --gpus 2`)."""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from test_gpu_parity import _device_batch, _d2h
    hashes = []
    for dev in (0, 1):
        ctx = _capi.Context(dev)
        n = 6
        pages, bufs, quads, angles = _device_batch(ctx, n, 1600, list(range(500, 500 + n)))
        params = DS.make_params()
        ctx.call("docscan_process_pages", n, pages, C.byref(params))
        ctx.sync()
        hashes.append([(sha(_d2h(ctx, dw, th, pw3, tw * 3)), sha(_d2h(ctx, db, th, pw1, tw))) for (_, dw, db, tw, th, pw3, pw1) in bufs])
        for src, dw, db, *_ in bufs:
            ctx.device_free(src); ctx.device_free(dw); ctx.device_free(db)
        ctx.close()
    assert hashes[0] == hashes[1]


def test_process_document_writes_the_reference_dumps(tmp_path):
    """process_document's default behaviour is the reference's: out_dir is created and the twelve stage files are written
    (DocScanner.py:277-346); scale_long <= 0 on a whole-photo page hands the photo through (DocScanner.py:29-30)."""
    cv2 = pytest.importorskip("cv2")
    from smart_image_processing_b200.synth import synth_page_numpy
    img, _ = synth_page_numpy(9, 600, 800)
    path = str(tmp_path / "page.png")
    cv2.imwrite(path, img)
    out = tmp_path / "dumps" / "nested"
    res = DS.process_document(path, out_dir=str(out), scale_long=500)
    names = ["scan_01_pre", "scan_02_quad", "scan_03_warped", "scan_04_illum", "scan_05_stretch", "scan_05a_inkmask",
             "scan_06_adapt", "scan_06b_weighted", "scan_07_deskew", "scan_08_clean"]
    for nm in names:
        assert os.path.exists(out / (nm + ".png")), nm
    eq(cv2.imread(str(out / "scan_08_clean.png"), cv2.IMREAD_GRAYSCALE), res["binary"], "scan_08_clean.png")
    eq(cv2.imread(str(out / "scan_03_warped.png"), cv2.IMREAD_COLOR), res["warped"], "scan_03_warped.png")
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    eq(cv2.imread(str(out / "scan_01_pre.png"), cv2.IMREAD_GRAYSCALE), cv2.bilateralFilter(gray, 9, 75, 75), "scan_01_pre.png")
    # whole-photo page with scale_long <= 0: the reference passes the photo through unchanged
    small = img[:300, :400].copy()
    cv2.imwrite(path, small)
    res = DS.process_document(path, out_dir=str(tmp_path / "o2"), scale_long=0, min_quad_area_ratio=2.0, angle=0.0)
    eq(res["warped"], small, "scale_long <= 0 passes the photo through")
    eq(res["binary"], O.hot_path(small, None, 0.0, scale_long=0)["clean"], "whole-photo page, scale_long 0")
    # out_dir is created even without dumps
    DS.process_document(path, out_dir=str(tmp_path / "o3"), scale_long=300, save_stages=False)
    assert os.path.isdir(tmp_path / "o3") and not os.listdir(tmp_path / "o3")


def test_tensor_core_adaptive_threshold_opt_in(monkeypatch):
    """The opt-in tensor-core GAUSSIAN_C (DOCSCAN_TC_ADAPTIVE=1: fixed-point mean on tcgen05, guard band, exact re-evaluation of
    the listed pixels) must give cv2's bytes exactly like the default kernel: pages, noise, ragged shapes, both presets' sizes."""
    monkeypatch.setenv("DOCSCAN_TC_ADAPTIVE", "1")
    rng = np.random.default_rng(35)
    for (h, w) in [(300, 260), (97, 131), (1600, 1131), (257, 1031), (33, 70)]:
        for k, c in [(35, 10), (31, 3), (11, 10), (3, 0), (61, -2)]:
            for kind in ("page", "noise"):
                g = page_like(rng, max(h, 16), max(w, 16))[:h, :w] if kind == "page" else rng.integers(0, 256, (h, w), dtype=np.uint8)
                eq(ops.adaptive_threshold(g, "gaussian", k, c), O.adaptive_threshold(g, "gaussian", k, c), f"tc adaptive {h}x{w} k={k} C={c} {kind}")


def test_packed_pair_adaptive_threshold_opt_in(monkeypatch):
    """The opt-in packed-pair GAUSSIAN_C (DOCSCAN_ADAPT_PACKED=1: FFMA2 / FADD2 on (f[u], f[u + 2]) pairs, two columns per thread
    in the column pass) must give cv2's bytes exactly like the default kernel: every unrolled radius, ragged and 1-pixel shapes,
    odd pitches (numpy views), the cv2 tail columns."""
    monkeypatch.setenv("DOCSCAN_ADAPT_PACKED", "1")
    rng = np.random.default_rng(36)
    for (h, w) in [(300, 260), (97, 131), (1600, 1131), (257, 1031), (33, 70), (1, 40), (40, 1), (64, 129), (50, 255)]:
        for k, c in [(35, 10), (31, 3), (27, 0), (19, 5), (11, 10), (3, 0), (9, -2)]:
            for kind in ("page", "noise"):
                g = page_like(rng, max(h, 16), max(w, 16))[:h, :w] if kind == "page" else rng.integers(0, 256, (h, w), dtype=np.uint8)
                eq(ops.adaptive_threshold(g, "gaussian", k, c), O.adaptive_threshold(g, "gaussian", k, c), f"packed adaptive {h}x{w} k={k} C={c} {kind}")


@pytest.mark.parametrize("grid", ["1", "2", "5"])
def test_tc_blur_long_tile_runs_per_cta(monkeypatch, grid):
    """The tensor-core blur deals every CTA a contiguous run of tiles and keeps a 64-entry ring of decoded tiles, refilled 32 at
    a time, with the epilogue lagging one tile behind the drain: force runs far longer than the ring (few CTAs, hundreds of tiles
    each, several pages and page sizes in one launch) and compare every page with the oracle — blur, both illumination
    epilogues with their min-max, and the ink-mask branch with its histogram."""
    monkeypatch.setenv("DOCSCAN_TC_GRID", grid)
    rng = np.random.default_rng(int(grid))
    imgs = [page_like(rng, 1600, 1131), page_like(rng, 700, 333), page_like(rng, 1600, 1131)]
    for k in (23, 51):
        for g in imgs[:2]:
            eq(ops.gaussian_blur(g, k), O.gaussian_blur_u8(g, k), f"tc blur grid={grid} k={k} {g.shape}")
    h, w = imgs[0].shape
    frac = (23 - 0.4) / min(h, w)
    for method in ("subtract", "divide"):
        eq(DS.illumination_correction(imgs[0], method, frac), O.illumination_correction(imgs[0], method, frac), f"illum {method} grid={grid}")
    eq(DS._compute_ink_mask(imgs[2], mask_blur_ksize=51), O._compute_ink_mask(imgs[2], mask_blur_ksize=51), f"ink mask grid={grid}")
    # several pages of two sizes in one launch through the fused pipeline
    photos = []
    for g in imgs:
        photos.append(np.stack([np.clip(g * s, 0, 255).astype(np.uint8) for s in (0.97, 1.0, 1.02)], -1))
    quads = [np.array([[30, 20], [p.shape[1] - 25, 28], [p.shape[1] - 20, p.shape[0] - 30], [22, p.shape[0] - 26]], np.float32) for p in photos]
    sl = 900
    wd, bn = DS.process_pages(photos, quads, [0.5, -1.0, 0.0], scale_long=sl)
    for i in range(3):
        ref = O.hot_path(photos[i], quads[i], [0.5, -1.0, 0.0][i], scale_long=sl)
        eq(bn[i], ref["clean"], f"pipeline page {i} grid={grid}")


def test_device_jpeg_decode_opt_in(tmp_path):
    """process_document(decode="device"): nvJPEG puts the photo straight into device memory.  Not bit-identical with cv2.imread
    (different IDCT / upsampling), so the checks are: the decoded photo is within a few grey levels of cv2's, and the page
    that comes out is the page the host-decoded photo gives except for a small fraction of threshold-edge pixels."""
    cv2 = pytest.importorskip("cv2")
    from smart_image_processing_b200 import control
    from smart_image_processing_b200.synth import synth_page_numpy
    img, quad = synth_page_numpy(21, 900, 1200)
    path = str(tmp_path / "page.jpg")
    cv2.imwrite(path, img, [cv2.IMWRITE_JPEG_QUALITY, 92])
    host = cv2.imread(path, cv2.IMREAD_COLOR)
    try:
        dev = control.load_image_device(path)
    except RuntimeError as e:
        pytest.skip(str(e))
    got = dev.to_numpy()
    assert got.shape == host.shape
    d = np.abs(got.astype(np.int16) - host.astype(np.int16))
    assert d.max() <= 6 and d.mean() < 0.6, (int(d.max()), float(d.mean()))
    res = DS.process_document(path, out_dir=str(tmp_path / "o"), scale_long=700, quad=quad, angle=0.5, save_stages=False, decode="device")
    ref = DS.process_document(path, out_dir=str(tmp_path / "o"), scale_long=700, quad=quad, angle=0.5, save_stages=False)
    assert res["binary"].shape == ref["binary"].shape
    assert np.mean(res["binary"] != ref["binary"]) < 0.10        # the rotated page is grey-level: every decoder-induced flip shows
    with pytest.raises(ValueError):
        DS.process_document(path, out_dir=str(tmp_path / "o"), scale_long=700, decode="device")
