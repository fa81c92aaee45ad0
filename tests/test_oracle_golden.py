"""CPU: the oracle (oracle/) against the fixtures produced by running the reference itself
(tests/golden/make_golden.py) and against the reference's own committed artefacts (KAT-1/KAT-2)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_npz
from oracle import oracle as O


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ops():
    return load_npz("ops.npz")


def test_gaussian_kernels_all_odd_k():
    kern = load_npz("gauss_kernels.npz")
    for k in range(1, 256, 2):
        assert np.array_equal(kern[f"k{k}"], O.gaussian_kernel_f32(k)), k
        q = O.gaussian_kernel_q8(k)
        assert q.sum() == 256 and (q >= 0).all() and np.array_equal(q, q[::-1])


def test_bgr2gray(ops):
    assert np.array_equal(O.bgr2gray(ops["rgb"]), ops["gray_bgr"])
    assert np.array_equal(O.bgr2gray(ops["rgb"], swap_rb=True), ops["gray_rgb"])


@pytest.mark.parametrize("k", [3, 5, 7, 9, 15, 23, 43, 51, 101])
def test_gaussian_blur_u8(ops, k):
    assert np.array_equal(O.gaussian_blur_u8(ops["g"], k), ops[f"blur_{k}"])


def test_divide_exhaustive(ops):
    a = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 256, 1)
    assert np.array_equal(O.divide255(a, a.T.copy()), ops["div_table"])


def test_normalize_and_otsu(ops):
    assert np.array_equal(O.normalize_minmax(ops["page"]), ops["normalize_page"])
    assert O.otsu_threshold(ops["page"]) == float(ops["otsu_page"])
    assert O.otsu_threshold(ops["g"]) == float(ops["otsu_g"])
    const = np.full((5, 7), 93, np.uint8)
    assert (O.normalize_minmax(const) == 0).all() and O.otsu_threshold(const) == 0.0


@pytest.mark.parametrize("kw,kh", [(2, 2), (3, 3), (9, 19), (4, 6), (31, 31), (101, 5)])
def test_morphology(ops, kw, kh):
    assert np.array_equal(O.erode(ops["g"], kw, kh), ops[f"erode_{kw}x{kh}"])
    assert np.array_equal(O.dilate(ops["g"], kw, kh), ops[f"dilate_{kw}x{kh}"])
    assert np.array_equal(O.morph_close(ops["g"], kw, kh, 2), ops[f"close_{kw}x{kh}_it2"])


def test_blackhat(ops):
    assert np.array_equal(O.blackhat(ops["page"], 9, 19), ops["blackhat_9x19"])


@pytest.mark.parametrize("k,c", [(3, 2), (11, 5), (31, 3), (35, 10)])
def test_adaptive_threshold(ops, k, c):
    assert np.array_equal(O.adaptive_threshold(ops["page"], "gaussian", k, c), ops[f"adapt_gauss_{k}_{c}"])
    assert np.array_equal(O.adaptive_threshold(ops["page"], "mean", k, c), ops[f"adapt_mean_{k}_{c}"])


@pytest.mark.parametrize("ang", [0.0, 0.5, -3.0, 9.5])
def test_rotation(ops, ang):
    page = ops["page"]
    m = O.rotation_matrix((page.shape[1] / 2.0, page.shape[0] / 2.0), ang)
    assert np.array_equal(m, ops[f"rotm_{ang}"])
    assert np.array_equal(O.warp_affine(page, m, (page.shape[1], page.shape[0])), ops[f"rot_{ang}"])
    if ang == 0.0:
        assert np.array_equal(ops[f"rot_{ang}"], page)


def test_perspective(ops):
    m = O.get_perspective_transform(ops["persp_quad"], ops["persp_dst"])
    assert np.array_equal(m, ops["persp_m"])
    assert np.array_equal(O.warp_perspective(ops["rgb"], m, (70, 99)), ops["persp_out"])


def test_kat1_morphseq_erode():
    """Reference artefact outputs/morphseq_01_gray.png -> morphseq_02_eroded.png (erode rect 2x2, 1 iter)."""
    kat = load_npz("kat.npz")
    assert np.array_equal(O.grayscale_erosion(kat["morphseq_01_gray"]), kat["morphseq_02_eroded"])
    assert not np.array_equal(O.erode(kat["morphseq_01_gray"], 3, 3), kat["morphseq_02_eroded"])


def test_kat2_constant_chain():
    """Reference artefacts outputs/scan_03_warped.png -> scan_04..08 (GUI preset, constant image)."""
    kat = load_npz("kat.npz")
    shape = tuple(int(v) for v in kat["scan_03_warped_shape"])
    warped = np.empty(shape, np.uint8)
    warped[:] = kat["scan_03_warped_value"]
    gray = O.bgr2gray(warped)
    illum = O.illumination_correction(gray, method="divide", blur_frac=0.05)
    stretch = O.contrast_stretch(illum)
    ink = O._compute_ink_mask(stretch, mask_blur_ksize=51)
    adapt = O.adaptive_binarize(stretch, block_size=31, C=3)
    weighted = O.mask_select(adapt, ink)
    desk = O.rotate(weighted, 0.0)
    clean = O.morph_cleanup(desk, ksize=1, iterations=0)
    assert clean is desk
    for name, img in (("04_illum", illum), ("05_stretch", stretch), ("05a_inkmask", ink), ("06_adapt", adapt),
                      ("06b_weighted", weighted), ("07_deskew", desk), ("08_clean", clean)):
        assert img.shape == tuple(int(v) for v in kat[f"scan_{name}_shape"])
        assert (img == kat[f"scan_{name}_value"][0]).all(), name


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_crops_every_stage(tag):
    z = load_npz("crops.npz")
    p = json.loads(str(z[f"{tag}_params"]))
    for drop in ("canny_low", "canny_high", "max_rotate"):
        p.pop(drop)
    out = O.hot_path(z[f"{tag}_input"], z[f"{tag}_quad"], float(z[f"{tag}_angle"]), **p)
    for k, v in out.items():
        assert np.array_equal(v, z[f"{tag}_{k}"]), (tag, k)


@pytest.mark.parametrize("preset", ["cli", "gui"])
def test_sample_jpg_every_stage(preset):
    """BASELINE.json config 1: public/sample.jpg through the whole path, both presets."""
    meta = json.load(open(os.path.join(GOLDEN, "sample_golden.json")))
    color = load_npz("sample_bgr.npz")["bgr"]
    assert sha(color) == meta["input_sha256"]
    g = meta["presets"][preset]
    p = dict(g["params"])
    for drop in ("canny_low", "canny_high", "max_rotate"):
        p.pop(drop)
    quad = np.frombuffer(bytes.fromhex(g["quad_f32_hex"]), np.float32).reshape(4, 2)
    out = O.hot_path(color, quad, float.fromhex(g["angle_hex"]), **p)
    for k, v in out.items():
        assert list(v.shape) == g["shapes"][k]
        assert sha(v) == g["sha256"][k], (preset, k)
    binz = load_npz(f"sample_{preset}_bin.npz")
    for k in binz.files:
        assert np.array_equal(out[k], binz[k])


def test_degenerate_shapes():
    rng = np.random.default_rng(0)
    for shape in ((1, 40), (40, 1), (1, 1), (2, 3)):
        g = rng.integers(0, 256, shape, dtype=np.uint8)
        for k in (3, 15):
            assert O.gaussian_blur_u8(g, k).shape == shape
            assert O.adaptive_threshold(g, "gaussian", k, 5).shape == shape
            assert O.adaptive_threshold(g, "mean", k, 5).shape == shape
        assert np.array_equal(O.rotate(g, 0.0), g)
        assert O.morph_close(g, 3).shape == shape
