"""ctypes binding of libdocscan.so (include/docscan.h).  No CPU fallback: if the library or a CUDA device
is missing, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdocscan.so")

HOST, DEVICE = 0, 1
OP_SUB, OP_DIV255, OP_MAX, OP_MASK_SELECT = 0, 1, 2, 3
MORPH_ERODE, MORPH_DILATE, MORPH_CLOSE, MORPH_BLACKHAT, MORPH_OPEN = 0, 1, 2, 3, 4
ADAPTIVE_MEAN, ADAPTIVE_GAUSSIAN = 0, 1
INTER_CUBIC, INTER_AREA = 2, 3


class DocscanError(RuntimeError):
    pass


class Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("pitch", C.c_int32),
                ("channels", C.c_int32), ("space", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("illum_method", C.c_int32), ("illum_blur_frac", C.c_double), ("block_size", C.c_int32),
                ("C", C.c_int32), ("thresh_method", C.c_int32), ("mask_blur_ksize", C.c_int32),
                ("blackhat_ksize", C.c_int32), ("blackhat_vertical_ratio", C.c_double),
                ("ink_dilate_iters", C.c_int32), ("mask_thresh_offset", C.c_int32), ("morph_ksize", C.c_int32),
                ("morph_iters", C.c_int32), ("cv_tail_compat", C.c_int32),
                ("canny_low", C.c_double), ("canny_high", C.c_double), ("max_rotate", C.c_double)]


class Page(C.Structure):
    _fields_ = [("src", Image), ("quad", C.c_float * 8), ("angle_deg", C.c_double), ("warped", Image), ("binary", Image),
                ("use_whole", C.c_int32)]


_lib = None
_lib_lock = threading.Lock()

_SIGS = {
    "docscan_version": (C.c_int, []),
    "docscan_strerror": (C.c_char_p, [C.c_int]),
    "docscan_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "docscan_destroy": (C.c_int, [C.c_void_p]),
    "docscan_sync": (C.c_int, [C.c_void_p]),
    "docscan_get_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "docscan_last_error": (C.c_char_p, [C.c_void_p]),
    "docscan_launch_count": (C.c_int64, [C.c_void_p]),
    "docscan_transfer_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "docscan_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "docscan_profile_dump": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "docscan_host_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "docscan_host_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "docscan_device_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "docscan_device_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "docscan_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "docscan_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "docscan_get_perspective_transform": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_double)]),
    "docscan_get_rotation_matrix": (C.c_int, [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]),
    "docscan_gaussian_kernel_f32": (C.c_int, [C.c_int, C.POINTER(C.c_float)]),
    "docscan_gaussian_kernel_q8": (C.c_int, [C.c_int, C.POINTER(C.c_int32)]),
    "docscan_otsu_from_hist": (C.c_int, [C.POINTER(C.c_int32), C.c_int64, C.POINTER(C.c_double)]),
    "docscan_warp_perspective": (C.c_int, [C.c_void_p, C.POINTER(Image), C.POINTER(C.c_double), C.POINTER(Image), C.POINTER(Image)]),
    "docscan_bgr2gray": (C.c_int, [C.c_void_p, C.POINTER(Image), C.POINTER(Image), C.c_int]),
    "docscan_gaussian_blur": (C.c_int, [C.c_void_p, C.POINTER(Image), C.c_int, C.POINTER(Image)]),
    "docscan_binary_op": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Image), C.POINTER(Image), C.POINTER(Image)]),
    "docscan_minmax": (C.c_int, [C.c_void_p, C.POINTER(Image), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "docscan_hist256": (C.c_int, [C.c_void_p, C.POINTER(Image), C.POINTER(C.c_int32)]),
    "docscan_normalize_minmax": (C.c_int, [C.c_void_p, C.POINTER(Image), C.POINTER(Image)]),
    "docscan_otsu_threshold": (C.c_int, [C.c_void_p, C.POINTER(Image), C.POINTER(C.c_double), C.POINTER(Image)]),
    "docscan_threshold_binary": (C.c_int, [C.c_void_p, C.POINTER(Image), C.c_int, C.POINTER(Image)]),
    "docscan_morph_rect": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Image), C.c_int, C.c_int, C.c_int, C.POINTER(Image)]),
    "docscan_adaptive_threshold": (C.c_int, [C.c_void_p, C.POINTER(Image), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Image)]),
    "docscan_warp_affine": (C.c_int, [C.c_void_p, C.POINTER(Image), C.POINTER(C.c_double), C.POINTER(Image)]),
    "docscan_canny": (C.c_int, [C.c_void_p, C.POINTER(Image), C.c_double, C.c_double, C.POINTER(Image)]),
    "docscan_hough_lines": (C.c_int, [C.c_void_p, C.POINTER(Image), C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32)]),
    "docscan_median_angle": (C.c_int, [C.POINTER(C.c_int32), C.c_double, C.POINTER(C.c_double)]),
    "docscan_skew_angle": (C.c_int, [C.c_void_p, C.POINTER(Image), C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]),
    "docscan_last_angles": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int]),
    "docscan_resize": (C.c_int, [C.c_void_p, C.POINTER(Image), C.POINTER(Image), C.c_int, C.c_int]),
    "docscan_illumination_correction": (C.c_int, [C.c_void_p, C.POINTER(Image), C.c_int, C.c_int, C.POINTER(Image)]),
    "docscan_ink_mask": (C.c_int, [C.c_void_p, C.POINTER(Image), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Image)]),
    "docscan_default_params": (None, [C.POINTER(Params)]),
    "docscan_target_size": (C.c_int, [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "docscan_warp_footprint": (C.c_int, [C.POINTER(Page), C.POINTER(C.c_int32)]),
    "docscan_process_pages": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Page), C.POINTER(Params)]),
    "docscan_synth_page": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(Image), C.POINTER(C.c_float)]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


def lib():
    """Loads libdocscan.so.  Raises if it has not been built: there is no fallback implementation."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise DocscanError(f"{LIB_PATH} is missing: build it with `python -m smart_image_processing_b200.build` "
                                   "(nvcc, sm_100a). There is no CPU fallback.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGS.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def image_of(a: np.ndarray) -> Image:
    """docscan_image view of a C-contiguous-rows uint8 numpy array (HxW or HxWx3)."""
    if a.dtype != np.uint8:
        raise TypeError(f"expected uint8 image, got {a.dtype}")
    if a.ndim == 2:
        h, w = a.shape
        ch = 1
    elif a.ndim == 3 and a.shape[2] in (1, 3):
        h, w, ch = a.shape
    else:
        raise ValueError(f"unsupported image shape {a.shape}")
    if a.size and (a.strides[-1] != 1 or (a.ndim == 3 and a.strides[1] != ch)):
        raise ValueError("image rows must be densely packed")
    return Image(a.ctypes.data, w, h, a.strides[0] if h > 1 else w * ch, ch, HOST)


def device_image(ptr: int, width: int, height: int, pitch: int, channels: int) -> Image:
    return Image(ptr, width, height, pitch, channels, DEVICE)


class _PinnedBlock:
    """Owner of one docscan_host_alloc block; numpy arrays made from it keep it (and its context) alive through `.base`."""

    def __init__(self, ctx: "Context", ptr: int, nbytes: int):
        self._ctx, self._ptr = ctx, ptr
        self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}

    def __del__(self):
        try:
            ctx, ptr = self._ctx, self._ptr
            self._ptr = None
            if ptr and getattr(ctx, "_h", None) and ctx._h.value:
                ctx._lib.docscan_host_free(ctx._h, C.c_void_p(ptr))
        except Exception:
            pass


class Context:
    """One docscan_ctx: a CUDA device + stream + scratch arena.  Not thread-safe; use one per thread."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        self._lib = lib()
        rc = self._lib.docscan_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self._h))
        if rc != 0:
            raise DocscanError(f"docscan_create(device={device}) failed: {self._lib.docscan_strerror(rc).decode()}")
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.docscan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self._lib.docscan_last_error(self._h).decode() or self._lib.docscan_strerror(rc).decode()
            raise DocscanError(f"{what}: {msg}")

    def call(self, name: str, *args):
        self.check(getattr(self._lib, name)(self._h, *args), name)

    def sync(self):
        self.call("docscan_sync")

    @property
    def launches(self) -> int:
        return int(self._lib.docscan_launch_count(self._h))

    @property
    def transfer_bytes(self):
        """(host->device, device->host) bytes copied for caller HOST images so far."""
        a, b = C.c_int64(), C.c_int64()
        self.call("docscan_transfer_bytes", C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def profile(self, on: bool):
        self.call("docscan_profile_enable", int(on))

    def profile_dump(self) -> dict:
        """{kernel name: (launches, total_ms, algorithmic_bytes)} since the last dump."""
        buf = C.create_string_buffer(1 << 16)
        n = self._lib.docscan_profile_dump(self._h, buf, len(buf))
        if n < 0:
            self.check(n, "docscan_profile_dump")
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms, nbytes = line.split()
            out[name] = (int(cnt), float(ms), float(nbytes))
        return out

    # pinned numpy arrays for the fast host path
    def pinned_empty(self, shape, dtype=np.uint8) -> np.ndarray:
        """Page-locked host array (docscan_host_alloc).  The block is owned by the array: it is released with
        docscan_host_free when the array and every view of it are gone (the block keeps this context alive until then)."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self.call("docscan_host_alloc", nbytes, C.byref(p))
        block = _PinnedBlock(self, p.value, max(nbytes, 1))
        return np.asarray(block)[:nbytes].view(dtype).reshape(shape)

    @property
    def stream(self) -> int:
        """The context's cudaStream_t as an integer handle."""
        p = C.c_void_p()
        self.call("docscan_get_stream", C.byref(p))
        return p.value or 0

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self.call("docscan_device_alloc", int(nbytes), C.byref(p))
        return p.value

    def device_free(self, ptr: int):
        self.call("docscan_device_free", C.c_void_p(ptr))


class DeviceBuffer:
    """A device-resident uint8 image owned by Python (docscan_device_alloc / docscan_device_free): `process_pages` accepts it
    in place of a numpy photo, e.g. for photos decoded on the device."""

    def __init__(self, ctx: Context, height: int, width: int, channels: int):
        self.ctx, self.shape, self.pitch = ctx, (height, width, channels), width * channels
        self.ptr = ctx.device_alloc(self.pitch * height)

    def image(self) -> Image:
        h, w, ch = self.shape
        return device_image(self.ptr, w, h, self.pitch, ch)

    def to_numpy(self) -> np.ndarray:
        out = np.empty(self.shape, np.uint8)
        self.ctx.check(lib().docscan_memcpy_d2h(self.ctx._h, out.ctypes.data, C.c_void_p(self.ptr), out.nbytes), "docscan_memcpy_d2h")
        return out

    def __del__(self):
        try:
            if self.ptr and getattr(self.ctx, "_h", None) and self.ctx._h.value:
                self.ctx.device_free(self.ptr)
            self.ptr = 0
        except Exception:
            pass


_tls = threading.local()


def default_context() -> Context:
    """Per-thread context on the current CUDA device (device 0, or LOCAL_RANK under torchrun)."""
    ctx = getattr(_tls, "ctx", None)
    if ctx is None:
        dev = int(os.environ.get("DOCSCAN_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        ctx = Context(dev)
        _tls.ctx = ctx
    return ctx
