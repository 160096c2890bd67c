"""Python mirror of the reference's am::morph interface (reference morph.h:11-76), bound to the flat
C wrappers of include/amx_morph.h.  Same member names, argument meaning and error behaviour, so
tests read like code written against the reference.  All work happens in libatomorph_b200.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from .engine import PARAM, pack_rgba, Engine  # noqa: F401  (re-exported)

SIZE_MAX = 2 ** 64 - 1

_bound = False


def _bind():
    global _bound
    L = _lib.lib()
    if _bound:
        return L
    vp, u64, u32, u16, i32, f64 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint16, C.c_int32, C.c_double

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("amx_morph_create", vp)
    sig("amx_morph_destroy", None, vp)
    sig("amx_morph_clear", None, vp)
    sig("amx_morph_last_error", C.c_char_p, vp)
    sig("amx_morph_device_context", vp, vp)
    sig("amx_morph_set", None, vp, i32, f64)
    sig("amx_morph_add_pixel", i32, vp, u64, u16, u16, u32)
    sig("amx_morph_add_pixels", i32, vp, u64, u64, vp, vp, vp)
    sig("amx_morph_add_frame", i32, vp, u64)
    sig("amx_morph_set_resolution", None, vp, u16, u16)
    sig("amx_morph_get_width", u16, vp)
    sig("amx_morph_get_height", u16, vp)
    sig("amx_morph_get_frame_count", u64, vp)
    sig("amx_morph_get_pixel_count", u64, vp, u64)
    sig("amx_morph_compute", None, vp)
    sig("amx_morph_compute_seconds", None, vp, f64)
    sig("amx_morph_iterate", None, vp, u64)
    sig("amx_morph_suspend", None, vp)
    sig("amx_morph_suspend_timeout", i32, vp, f64)
    sig("amx_morph_is_busy", i32, vp)
    sig("amx_morph_synchronize", i32, vp)
    sig("amx_morph_next_state", None, vp)
    sig("amx_morph_get_state", C.c_uint, vp)
    sig("amx_morph_get_energy", f64, vp)
    sig("amx_morph_get_frame_key", u64, vp, f64)
    sig("amx_morph_get_time", f64, vp, u64, u64)
    sig("amx_morph_normalize_time", f64, vp, f64)
    sig("amx_morph_get_pixels", i32, vp, f64, vp)
    sig("amx_morph_get_pixels_blob", C.c_int64, vp, u64, f64, u64, vp, vp, C.POINTER(u64))
    sig("amx_morph_get_pixel", u32, vp, u64, u64)
    sig("amx_morph_get_average_pixel", None, vp, u64, vp, vp)
    sig("amx_morph_get_average_pixel_blob", None, vp, u64, u64, vp, vp)
    sig("amx_morph_get_background", u32, vp, u16, u16, f64)
    sig("amx_morph_get_blob_count", u64, vp, u64)
    sig("amx_morph_get_blob_count_all", u64, vp)
    sig("amx_morph_get_blob", i32, vp, u64, u64, vp, vp)
    sig("amx_morph_get_blob_surface", i32, vp, u64, u64, vp)
    sig("amx_morph_blob2pixel", u32, vp, u64, u64, vp)
    sig("amx_morph_interpolate_point", u64, vp, u64, u64, f64)
    sig("amx_morph_interpolate_color", u32, vp, u32, u32, f64)
    sig("amx_morph_interpolate_color_eased", u32, vp, u32, u32, f64, f64, f64)
    _bound = True
    return L


MORPH_SYMBOLS = [
    "amx_morph_create", "amx_morph_destroy", "amx_morph_clear", "amx_morph_last_error", "amx_morph_device_context",
    "amx_morph_set", "amx_morph_add_pixel", "amx_morph_add_pixels", "amx_morph_add_frame", "amx_morph_set_resolution",
    "amx_morph_get_width", "amx_morph_get_height", "amx_morph_get_frame_count", "amx_morph_get_pixel_count",
    "amx_morph_compute", "amx_morph_compute_seconds", "amx_morph_iterate", "amx_morph_suspend",
    "amx_morph_suspend_timeout", "amx_morph_is_busy", "amx_morph_synchronize", "amx_morph_next_state",
    "amx_morph_get_state", "amx_morph_get_energy", "amx_morph_get_frame_key", "amx_morph_get_time",
    "amx_morph_normalize_time", "amx_morph_get_pixels", "amx_morph_get_pixels_blob", "amx_morph_get_pixel",
    "amx_morph_get_average_pixel", "amx_morph_get_average_pixel_blob", "amx_morph_get_background",
    "amx_morph_get_blob_count", "amx_morph_get_blob_count_all", "amx_morph_get_blob", "amx_morph_get_blob_surface",
    "amx_morph_blob2pixel", "amx_morph_interpolate_point", "amx_morph_interpolate_color",
    "amx_morph_interpolate_color_eased",
]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Morph:
    """am::morph.  Setters are `set_<name>(value)` exactly as in morph.h:52-76."""

    def __init__(self):
        self.L = _bind()
        self.h = self.L.amx_morph_create()

    def close(self):
        if getattr(self, "h", None):
            self.L.amx_morph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getattr__(self, name):
        # set_blob_threshold(...), set_motion(...), set_seed(...) ... one per reference setter
        if name.startswith("set_") and name[4:] in PARAM:
            pid = PARAM[name[4:]]
            return lambda v: self.L.amx_morph_set(self.h, pid, float(v))
        raise AttributeError(name)

    def last_error(self):
        return self.L.amx_morph_last_error(self.h).decode()

    def device_context(self):
        return self.L.amx_morph_device_context(self.h)

    def clear(self):
        self.L.amx_morph_clear(self.h)

    # ---- ingest
    def add_pixel(self, frame, x, y, rgba):
        return bool(self.L.amx_morph_add_pixel(self.h, frame, x, y, rgba))

    def add_frame(self, frame):
        return bool(self.L.amx_morph_add_frame(self.h, frame))

    def add_image(self, frame, rgba, present=None):
        """Bulk add_pixel for every pixel with alpha != 0 (what demo/main.cpp:96-127 does per PNG)."""
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        if present is None:
            present = rgba[..., 3] != 0
        ys, xs = np.nonzero(present)
        if len(xs) == 0:
            self.add_frame(frame)
            return
        x = xs.astype(np.uint16)
        y = ys.astype(np.uint16)
        c = np.ascontiguousarray(pack_rgba(rgba)[ys, xs].astype(np.uint32))
        self.L.amx_morph_add_pixels(self.h, frame, len(x), _p(x), _p(y), _p(c))

    def set_resolution(self, w, h):
        self.L.amx_morph_set_resolution(self.h, w, h)

    def get_width(self):
        return self.L.amx_morph_get_width(self.h)

    def get_height(self):
        return self.L.amx_morph_get_height(self.h)

    def get_frame_count(self):
        return self.L.amx_morph_get_frame_count(self.h)

    def get_pixel_count(self, frame):
        return self.L.amx_morph_get_pixel_count(self.h, frame)

    # ---- run control
    def compute(self, seconds=None):
        if seconds is None:
            self.L.amx_morph_compute(self.h)
        else:
            self.L.amx_morph_compute_seconds(self.h, float(seconds))

    def iterate(self, n):
        self.L.amx_morph_iterate(self.h, int(n))

    def suspend(self, timeout=None):
        if timeout is None:
            self.L.amx_morph_suspend(self.h)
            return True
        return bool(self.L.amx_morph_suspend_timeout(self.h, float(timeout)))

    def is_busy(self):
        return bool(self.L.amx_morph_is_busy(self.h))

    def synchronize(self):
        return bool(self.L.amx_morph_synchronize(self.h))

    def next_state(self):
        self.L.amx_morph_next_state(self.h)

    def get_state(self):
        return self.L.amx_morph_get_state(self.h)

    def get_energy(self):
        return self.L.amx_morph_get_energy(self.h)

    def wait(self):
        """Block until an iterate(n) finished (the reference's callers poll is_busy)."""
        import time
        while self.is_busy():
            time.sleep(0.0005)

    # ---- time
    def get_frame_key(self, t):
        return self.L.amx_morph_get_frame_key(self.h, float(t))

    def get_time(self, f, total):
        return self.L.amx_morph_get_time(self.h, f, total)

    def normalize_time(self, t):
        return self.L.amx_morph_normalize_time(self.h, float(t))

    # ---- fetch
    def get_pixels(self, t):
        w, h = self.get_width(), self.get_height()
        out = np.zeros((h, w), dtype=np.uint32)
        self.L.amx_morph_get_pixels(self.h, float(t), _p(out))
        return out

    def get_pixels_blob(self, blob, t):
        cap = 1 << 22
        xy = np.zeros((cap, 2), dtype=np.uint16)
        c = np.zeros(cap, dtype=np.uint32)
        g = C.c_uint64(0)
        n = self.L.amx_morph_get_pixels_blob(self.h, blob, float(t), cap, _p(xy), _p(c), C.byref(g))
        if n < 0:
            return None
        return dict(group=int(g.value), xy=xy[:n].copy(), rgba=c[:n].copy())

    def get_pixel(self, frame, pos):
        return self.L.amx_morph_get_pixel(self.h, frame, pos)

    def get_average_pixel(self, frame, blob=None):
        xy = np.zeros(2, dtype=np.uint16)
        c = np.zeros(1, dtype=np.uint32)
        if blob is None:
            self.L.amx_morph_get_average_pixel(self.h, frame, _p(xy), _p(c))
        else:
            self.L.amx_morph_get_average_pixel_blob(self.h, frame, blob, _p(xy), _p(c))
        return int(xy[0]), int(xy[1]), int(c[0])

    def get_background(self, x, y, t):
        return self.L.amx_morph_get_background(self.h, x, y, float(t))

    def get_blob_count(self, frame=None):
        if frame is None:
            return self.L.amx_morph_get_blob_count_all(self.h)
        return self.L.amx_morph_get_blob_count(self.h, frame)

    def get_blob(self, frame, blob):
        stats = np.zeros(6)
        meta = np.zeros(2, dtype=np.uint64)
        if not self.L.amx_morph_get_blob(self.h, frame, blob, _p(stats), _p(meta)):
            return None
        surf = np.zeros(int(meta[1]), dtype=np.uint64)
        if len(surf):
            self.L.amx_morph_get_blob_surface(self.h, frame, blob, _p(surf))
        return dict(stats=stats, group=int(meta[0]), surface=surf)

    def interpolate_point(self, p1, p2, w):
        return self.L.amx_morph_interpolate_point(self.h, int(p1), int(p2), float(w))

    def interpolate_color(self, c1, c2, w, lag=None, slope=None):
        if lag is None:
            return self.L.amx_morph_interpolate_color(self.h, int(c1), int(c2), float(w))
        return self.L.amx_morph_interpolate_color_eased(self.h, int(c1), int(c2), float(lag), float(slope), float(w))
