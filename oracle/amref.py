"""ctypes binding of oracle/_ref/libamref.so -- TEST INFRASTRUCTURE ONLY.

The shared object is the UNMODIFIED reference (1Hyena/atomorph) compiled by
oracle/Makefile plus oracle/ref_harness.cpp.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libamref.so")

# am:: constants (reference atomorph.h:235-246, 304-306)
RGB, HSP, NONE, LINEAR, SPLINE, COSINE, PERLIN = 0, 1, 2, 3, 4, 5, 6
STATE_BLOB_DETECTION, STATE_BLOB_UNIFICATION, STATE_BLOB_MATCHING, STATE_ATOM_MORPHING, STATE_DONE = 0, 1, 2, 3, 4
TEXTURE, AVERAGE, DISTINCT = 0, 1, 2
HAS_PIXEL, HAS_FLUID = 1, 2

# include/amx_params.h
P = dict(blob_delimiter=0, blob_threshold=1, blob_max_size=2, blob_min_size=3, blob_box_grip=4,
         blob_box_samples=5, blob_number=6, blob_rgba_weight=7, blob_size_weight=8, blob_xy_weight=9,
         degeneration=10, density=11, motion=12, fading=13, threads=14, cycle_length=15, feather=16,
         keep_background=17, finite=18, show_blobs=19, fluid=20, seed=21)

FP_STRIDE = 24  # fluid particle record, see ref_harness.cpp

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError("oracle/_ref/libamref.so missing: run `make -C oracle ref` where /root/reference exists")
    L = C.CDLL(LIB_PATH)
    vp, u64, f64, u32, u16, i32 = C.c_void_p, C.c_uint64, C.c_double, C.c_uint32, C.c_uint16, C.c_int

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("amref_create", vp)
    sig("amref_destroy", None, vp)
    sig("amref_set", None, vp, i32, f64)
    sig("amref_add_pixels", None, vp, u64, u64, vp, vp, vp)
    sig("amref_add_frame", i32, vp, u64)
    sig("amref_set_resolution", None, vp, u16, u16)
    sig("amref_sync", None, vp)
    sig("amref_state", C.c_uint, vp)
    sig("amref_iterate", None, vp, u64)
    sig("amref_next_state", None, vp)
    sig("amref_energy", f64, vp)
    sig("amref_best_blob_energy", f64, vp)
    sig("amref_worker_values", None, vp, vp)
    sig("amref_e1_state", u64, vp)
    sig("amref_run_until", C.c_uint, vp, C.c_uint, u64, u64)
    sig("amref_true_cost", f64, vp)
    sig("amref_frame_count", u64, vp)
    sig("amref_frame_keys", None, vp, vp)
    sig("amref_pixel_count", u64, vp, u64)
    sig("amref_get_pixel", u32, vp, u64, u64)
    sig("amref_stored_pixel", u32, vp, u64, u64, vp)
    sig("amref_average_pixel", None, vp, u64, vp, vp)
    sig("amref_frame_means", None, vp, u64, vp)
    sig("amref_bbox", None, vp, vp)
    sig("amref_perlin", None, vp, i32, vp)
    sig("amref_blob_count", u64, vp, u64)
    sig("amref_blob_info", i32, vp, u64, u64, vp, vp)
    sig("amref_blob_surface", None, vp, u64, u64, vp)
    sig("amref_blob_labels", None, vp, u64, u32, u32, vp)
    sig("amref_chain_count", u64, vp)
    sig("amref_chain_info", None, vp, u64, vp)
    sig("amref_chain_points", None, vp, u64, vp)
    sig("amref_import_blobs", None, vp, u64, u64, vp, vp, vp, vp)
    sig("amref_import_chain", i32, vp, u64, u64, u64, u64, vp)
    sig("amref_finish_import", None, vp)
    sig("amref_render", None, vp, f64, vp)
    sig("amref_render_blob", C.c_int64, vp, u64, f64, u64, vp, vp, vp)
    sig("amref_get_time", f64, vp, u64, u64)
    sig("amref_get_frame_key", u64, vp, f64)
    sig("amref_get_background", u32, vp, u16, u16, f64)
    sig("amref_interpolate_point", u64, vp, u64, u64, f64)
    sig("amref_interpolate_color", u32, vp, u32, u32, f64, f64, f64, i32)
    sig("amref_rgb_to_hsp", u32, u32)
    sig("amref_hsp_to_rgb", u32, u32)
    sig("amref_color_distance", f64, u32, u32)
    sig("amref_octave_noise", f64, C.c_uint, f64, f64, i32)
    sig("amref_spline_point", None, u64, vp, vp, f64, vp)
    sig("amref_point_distance", u64, u64, u64)
    sig("amref_time_morph_steps", f64, vp, u64)
    sig("amref_time_render", f64, vp, f64, vp)
    sig("amref_hardware_concurrency", C.c_uint)
    sig("amref_fluid_create", vp, C.c_uint, C.c_uint, C.c_uint)
    sig("amref_fluid_destroy", None, vp)
    sig("amref_fluid_set_particles", None, vp, u64, vp)
    sig("amref_fluid_get_particles", None, vp, u64, vp)
    sig("amref_fluid_step", None, vp, u64, f64, f64)
    sig("amref_fluid_get_nodes", None, vp, vp)
    sig("amref_morph_fluid_sanitize", None, vp)
    sig("amref_morph_fluid_count", u64, vp)
    sig("amref_morph_fluid_get", None, vp, u64, vp)
    sig("amref_morph_fluid_dims", None, vp, vp)
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pack_rgba(rgba_u8):
    """(..., 4) uint8 -> (...) uint32 little-endian r | g<<8 | b<<16 | a<<24."""
    a = np.ascontiguousarray(rgba_u8, dtype=np.uint8)
    return a.view(np.uint32).reshape(a.shape[:-1])


def unpack_rgba(u32):
    a = np.ascontiguousarray(u32, dtype=np.uint32)
    return a.view(np.uint8).reshape(a.shape + (4,))


class RefMorph:
    """The reference's am::morph, driven deterministically (threads=0, iterate)."""

    def __init__(self, **params):
        self.L = lib()
        self.h = self.L.amref_create()
        self.width = self.height = 0
        self.set(threads=0, cycle_length=0)
        self.set(**params)

    def close(self):
        if self.h:
            self.L.amref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, **params):
        for k, v in params.items():
            self.L.amref_set(self.h, P[k], float(v))

    # ---- ingest
    def add_image(self, frame, rgba, present=None):
        """rgba: (H, W, 4) uint8.  Adds every pixel with alpha != 0 (demo/main.cpp:96-127) unless
        `present` (H, W) bool is given, in which case exactly those pixels are added."""
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        H, W = rgba.shape[:2]
        if present is None:
            present = rgba[..., 3] != 0
        ys, xs = np.nonzero(present)
        if len(xs) == 0:
            self.L.amref_add_frame(self.h, frame)
            return
        x = xs.astype(np.uint16)
        y = ys.astype(np.uint16)
        c = pack_rgba(rgba)[ys, xs].astype(np.uint32)
        self.L.amref_add_pixels(self.h, frame, len(x), _p(x), _p(y), _p(c))

    def set_resolution(self, w, h):
        self.width, self.height = int(w), int(h)
        self.L.amref_set_resolution(self.h, w, h)

    # ---- run control
    def sync(self):
        self.L.amref_sync(self.h)

    def state(self):
        return self.L.amref_state(self.h)

    def iterate(self, n):
        self.L.amref_iterate(self.h, int(n))

    def run_until(self, target=STATE_ATOM_MORPHING, chunk=20000, match_steps=0):
        return self.L.amref_run_until(self.h, target, chunk, match_steps)

    def true_cost(self):
        return self.L.amref_true_cost(self.h)

    def energy(self):
        return self.L.amref_energy(self.h)

    def worker_values(self):
        out = np.zeros(10)
        self.L.amref_worker_values(self.h, _p(out))
        return dict(zip(("blob_map_e", "best_e", "best_blob_map_e", "bbox_d", "blob_map_w", "blob_map_h", "counter",
                         "w_rgba", "w_size", "w_xy"), out))

    def e1_state(self):
        return int(self.L.amref_e1_state(self.h))

    def best_blob_energy(self):
        return self.L.amref_best_blob_energy(self.h)

    # ---- dumps
    def frame_keys(self):
        n = self.L.amref_frame_count(self.h)
        out = np.zeros(n, dtype=np.uint64)
        self.L.amref_frame_keys(self.h, _p(out))
        return [int(k) for k in out]

    def bbox(self):
        out = np.zeros(4, dtype=np.uint16)
        self.L.amref_bbox(self.h, _p(out))
        return tuple(int(v) for v in out)

    def perlin_tables(self):
        lag = np.zeros(512, dtype=np.int32)
        slope = np.zeros(512, dtype=np.int32)
        self.L.amref_perlin(self.h, 0, _p(lag))
        self.L.amref_perlin(self.h, 1, _p(slope))
        return lag, slope

    def fetch_image(self, frame):
        """get_pixel(frame, pos) for every position of the WxH canvas (RGB after the HSP round trip)."""
        out = np.zeros((self.height, self.width), dtype=np.uint32)
        for y in range(self.height):
            for x in range(self.width):
                out[y, x] = self.L.amref_get_pixel(self.h, frame, y * 65536 + x)
        return out

    def stored_image(self, frame):
        out = np.zeros((self.height, self.width), dtype=np.uint32)
        pres = np.zeros((self.height, self.width), dtype=bool)
        flag = C.c_int(0)
        for y in range(self.height):
            for x in range(self.width):
                out[y, x] = self.L.amref_stored_pixel(self.h, frame, y * 65536 + x, C.byref(flag))
                pres[y, x] = bool(flag.value)
        return out, pres

    def frame_means(self, frame):
        out = np.zeros(6)
        self.L.amref_frame_means(self.h, frame, _p(out))
        return out

    def average_pixel(self, frame):
        xy = np.zeros(2, dtype=np.uint16)
        c = np.zeros(1, dtype=np.uint32)
        self.L.amref_average_pixel(self.h, frame, _p(xy), _p(c))
        return int(xy[0]), int(xy[1]), int(c[0])

    def blobs(self, frame):
        """-> list of dict(stats[6], group, surface (sorted uint64 positions))."""
        out = []
        n = self.L.amref_blob_count(self.h, frame)
        for b in range(n):
            stats = np.zeros(6)
            meta = np.zeros(2, dtype=np.uint64)
            if not self.L.amref_blob_info(self.h, frame, b, _p(stats), _p(meta)):
                out.append(None)
                continue
            surf = np.zeros(int(meta[1]), dtype=np.uint64)
            if len(surf):
                self.L.amref_blob_surface(self.h, frame, b, _p(surf))
            out.append(dict(stats=stats, group=int(meta[0]), surface=surf))
        return out

    def blob_labels(self, frame):
        out = np.zeros((self.height, self.width), dtype=np.int32)
        self.L.amref_blob_labels(self.h, frame, self.width, self.height, _p(out))
        return out

    def chains(self):
        """-> list of dict(key, width, height, max_surface, words (height, width) uint64 column-major)."""
        out = []
        for i in range(self.L.amref_chain_count(self.h)):
            info = np.zeros(4, dtype=np.uint64)
            self.L.amref_chain_info(self.h, i, _p(info))
            key, w, h, ms = (int(v) for v in info)
            words = np.zeros((h, w), dtype=np.uint64)
            if w * h:
                self.L.amref_chain_points(self.h, i, _p(words))
            out.append(dict(key=key, width=w, height=h, max_surface=ms, words=words))
        return out

    # ---- imports
    def import_blobs(self, frame, blobs):
        n = len(blobs)
        group = np.array([b["group"] for b in blobs], dtype=np.uint64)
        stats = np.ascontiguousarray(np.array([b["stats"] for b in blobs], dtype=np.float64).reshape(n, 6))
        offs = np.zeros(n + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(b["surface"]) for b in blobs])
        pos = (np.concatenate([np.asarray(b["surface"], dtype=np.uint64) for b in blobs])
               if n else np.zeros(0, dtype=np.uint64))
        pos = np.ascontiguousarray(pos)
        self.L.amref_import_blobs(self.h, frame, n, _p(group), _p(stats), _p(offs), _p(pos))

    def import_chain(self, key, words, max_surface):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        h, w = words.shape
        if not self.L.amref_import_chain(self.h, key, w, h, int(max_surface), _p(words)):
            raise MemoryError("renew_chain failed")

    def finish_import(self):
        self.L.amref_finish_import(self.h)

    # ---- render
    def get_time(self, f, total):
        return self.L.amref_get_time(self.h, f, total)

    def render(self, t):
        out = np.zeros((self.height, self.width), dtype=np.uint32)
        self.L.amref_render(self.h, float(t), _p(out))
        return out

    def render_blob(self, b, t):
        cap = self.width * self.height * 2 + 16
        xy = np.zeros((cap, 2), dtype=np.uint16)
        c = np.zeros(cap, dtype=np.uint32)
        g = C.c_uint64(0)
        n = self.L.amref_render_blob(self.h, b, float(t), cap, _p(xy), _p(c), C.byref(g))
        if n < 0:
            return None
        return dict(group=int(g.value), xy=xy[:n].copy(), rgba=c[:n].copy())

    def get_background(self, x, y, t):
        return self.L.amref_get_background(self.h, x, y, float(t))

    def time_morph_steps(self, n):
        return self.L.amref_time_morph_steps(self.h, int(n))

    def time_render(self, t):
        out = np.zeros((self.height, self.width), dtype=np.uint32)
        return self.L.amref_time_render(self.h, float(t), _p(out)), out

    def fluid_sanitize(self):
        self.L.amref_morph_fluid_sanitize(self.h)

    def fluid_particles(self):
        n = self.L.amref_morph_fluid_count(self.h)
        rec = np.zeros((n, FP_STRIDE))
        if n:
            self.L.amref_morph_fluid_get(self.h, n, _p(rec))
        return rec


class RefFluid:
    """Stand-alone reference FluidModel (fluidmodel.cpp) for single-step parity."""

    def __init__(self, gx, gy, n):
        self.L = lib()
        self.gx, self.gy, self.n = gx, gy, n
        self.h = self.L.amref_fluid_create(gx, gy, n)

    def close(self):
        if self.h:
            self.L.amref_fluid_destroy(self.h)
            self.h = None

    def set_particles(self, rec):
        rec = np.ascontiguousarray(rec, dtype=np.float64)
        assert rec.shape == (self.n, FP_STRIDE)
        self.L.amref_fluid_set_particles(self.h, self.n, _p(rec))

    def get_particles(self):
        rec = np.zeros((self.n, FP_STRIDE))
        self.L.amref_fluid_get_particles(self.h, self.n, _p(rec))
        return rec

    def step(self, steps_left, freedom_radius, t):
        self.L.amref_fluid_step(self.h, int(steps_left), float(freedom_radius), float(t))

    def nodes(self):
        out = np.zeros((self.gy, self.gx, 13))
        self.L.amref_fluid_get_nodes(self.h, _p(out))
        return out


# ---- pure functions
def rgb_to_hsp(c):
    return lib().amref_rgb_to_hsp(int(c))


def hsp_to_rgb(c):
    return lib().amref_hsp_to_rgb(int(c))


def color_distance(a, b):
    return lib().amref_color_distance(int(a), int(b))


def octave_noise(seed, x, y, octaves=8):
    return lib().amref_octave_noise(int(seed), float(x), float(y), int(octaves))


def spline_point(xs, ys, t):
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    ys = np.ascontiguousarray(ys, dtype=np.float64)
    out = np.zeros(2)
    lib().amref_spline_point(len(xs), _p(xs), _p(ys), float(t), _p(out))
    return out


def point_distance(w1, w2):
    return lib().amref_point_distance(int(w1), int(w2))


def hardware_concurrency():
    return lib().amref_hardware_concurrency()
