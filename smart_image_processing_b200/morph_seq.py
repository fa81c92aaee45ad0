"""Drop-in for the reference's morph_seq module (shipped only as __pycache__/morph_seq.cpython-310.pyc; function
names, defaults and step order recovered from its bytecode): gray -> erode 2x2 -> Otsu ->
threshold 127 + close 2x2.  Every pixel step is a CUDA kernel of libdocscan.so."""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import ops

KSIZE = 2
ITERATIONS = 1


def to_grayscale(rgb: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(rgb, COLOR_RGB2GRAY)."""
    return ops.bgr2gray(rgb, swap_rb=True)


def grayscale_erosion(gray: np.ndarray, ksize: int = KSIZE, iterations: int = ITERATIONS) -> np.ndarray:
    """cv2.erode with a ksize x ksize rectangle."""
    return ops.erode(gray, ksize, ksize, iterations)


def otsu_binarize(gray: np.ndarray) -> np.ndarray:
    """cv2.threshold(gray, 0, 255, THRESH_BINARY + THRESH_OTSU)[1].  (The reference computes this image and
    then forgets to return it; returning it is the evident intent.)"""
    _, binary = ops.otsu_threshold(gray, return_image=True)
    return binary


def binary_closing(binary: np.ndarray, ksize: int = KSIZE, iterations: int = ITERATIONS) -> np.ndarray:
    """threshold at 127, then MORPH_CLOSE with a ksize x ksize rectangle."""
    return ops.morph_close(ops.threshold_binary(binary, 127), ksize, ksize, iterations)


def process_morph_seq(input_path: str, out_dir: str = "outputs", save_intermediate: bool = True) -> Dict[str, np.ndarray]:
    from . import control
    cv2 = control._cv2()
    rgb = cv2.cvtColor(control.load_image(input_path), cv2.COLOR_BGR2RGB)
    gray = to_grayscale(rgb)
    eroded = grayscale_erosion(gray)
    otsu = otsu_binarize(eroded)
    closed = binary_closing(otsu)
    out = {"original": rgb, "step1_gray": gray, "step2_eroded": eroded, "step3_otsu": otsu, "step4_closed": closed}
    if save_intermediate:
        import os
        os.makedirs(out_dir, exist_ok=True)
        for name, key in (("morphseq_01_gray.png", "step1_gray"), ("morphseq_02_eroded.png", "step2_eroded"),
                          ("morphseq_03_otsu.png", "step3_otsu"), ("morphseq_04_closed.png", "step4_closed")):
            cv2.imwrite(os.path.join(out_dir, name), out[key])
    return out
