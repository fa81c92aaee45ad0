#!/usr/bin/env python3
"""Fixtures for the whole-photo fallback (resize_long_side, DocScanner.py:27-36, taken at :313), made by RUNNING THE
REFERENCE's own function.  Build container only (needs /root/reference and cv2):

    python tests/golden/make_golden_resize.py        ->  tests/golden/resize.npz, resize_golden.json

INTER_AREA results do not depend on IPP.  For INTER_CUBIC the pip wheels of cv2 route the call to IPP's closed-source
float cubic, which differs from OpenCV's own code by +-1 in a few percent of the pixels; the fixtures pin OpenCV's own
code (cv2.ipp.setUseIPP(False) while generating) and record how far the IPP flavour is from it.
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


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out, meta = {}, {"cv2": cv2.__version__, "cases": {}, "full": {}}
    s3 = DS.load_image(os.path.join(REF, "public", "sample3.jpg"))
    d2 = DS.load_image(os.path.join(REF, "public", "DIP test", "document2.png"))
    crops = {"s3a": s3[300:420, 200:360], "s3b": s3[700:833, 400:571], "d2a": d2[100:231, 50:199], "d2g": d2[300:400, 300:450, 1]}
    cases = [("s3a", 200), ("s3a", 100), ("s3a", 80), ("s3a", 53), ("s3a", 160), ("s3b", 300), ("s3b", 57), ("s3b", 171),
             ("d2a", 400), ("d2a", 75), ("d2a", 149), ("d2g", 333), ("d2g", 50), ("d2g", 75)]
    for tag, crop in crops.items():
        out[f"in_{tag}"] = np.ascontiguousarray(crop)
    for tag, sl in cases:
        crop = np.ascontiguousarray(crops[tag])
        cv2.ipp.setUseIPP(False)
        ref = DS.resize_long_side(crop, sl)
        cv2.ipp.setUseIPP(True)
        ipp = DS.resize_long_side(crop, sl)
        d = np.abs(ref.astype(int) - ipp.astype(int))
        out[f"out_{tag}_{sl}"] = ref
        meta["cases"][f"{tag}_{sl}"] = {"shape": list(ref.shape), "kind": "area" if sl < max(crop.shape[:2]) else "cubic",
                                        "ipp_differs": int((d > 0).sum()), "ipp_max_abs_diff": int(d.max())}
    # full images: hashes only
    for name, img, sl in (("sample3_1600", s3, 1600), ("sample3_1200", s3, 1200), ("document2_1600", d2, 1600), ("document2_1200", d2, 1200)):
        cv2.ipp.setUseIPP(False)
        ref = DS.resize_long_side(img, sl)
        cv2.ipp.setUseIPP(True)
        ipp = DS.resize_long_side(img, sl)
        d = np.abs(ref.astype(int) - ipp.astype(int))
        meta["full"][name] = {"input_sha256": sha(img), "input_shape": list(img.shape), "shape": list(ref.shape), "sha256": sha(ref),
                              "ipp_differs": int((d > 0).sum()), "ipp_max_abs_diff": int(d.max())}
    out["sample3_bgr"] = s3
    np.savez_compressed(os.path.join(HERE, "resize.npz"), **out)
    with open(os.path.join(HERE, "resize_golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta["full"], indent=1))


if __name__ == "__main__":
    main()
