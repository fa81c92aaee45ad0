"""How many pixels of real pipeline pages fall into the adaptive threshold's guard band (developer tool, B200)."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from smart_image_processing_b200 import DocScanner as DS
from smart_image_processing_b200.synth import synth_page_numpy, synth_angle
os.environ["DOCSCAN_TC_DEBUG"] = tempfile.NamedTemporaryFile(suffix=".bin", delete=False).name
for seed in (0, 1):
    img, quad = synth_page_numpy(seed, 3000, 4000)
    st = DS.hot_path(img, quad, synth_angle(seed), scale_long=1600)
    s = st["stretch"]
    print("seed", seed, "stretch stats: mean %.1f std %.1f" % (s.mean(), s.std()), "hist low bins", np.bincount(s.ravel(), minlength=256)[:12].tolist())
