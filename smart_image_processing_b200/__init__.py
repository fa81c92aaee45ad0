"""b200-docscan: the per-pixel document-scan path of Brianlov/Smart-Image-Processing (DocScanner.py, morph_seq)
as hand-written sm_100a CUDA kernels behind a C ABI (include/docscan.h).

    from smart_image_processing_b200 import DocScanner as DS      # drop-in for the reference module
    from smart_image_processing_b200 import ops                   # one function per cv2 call on the path

Importing the package does not touch CUDA; the first call loads libdocscan.so and raises if it (or a GPU)
is missing.  There is no CPU fallback.
"""
__version__ = "0.1.0"
