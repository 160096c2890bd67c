"""ctypes binding of oracle/libamoracle.so (oracle/am_oracle.cpp) -- TEST INFRASTRUCTURE ONLY.

The CPU restatement of the reference's hot-path algorithms.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libamoracle.so")

_lib = None


def available():
    return os.path.exists(LIB_PATH)


class Scene(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("cw", C.c_uint32), ("ch", C.c_uint32),
        ("nframes", C.c_uint32), ("nchains", C.c_uint32),
        ("bbox", C.c_uint32 * 4),
        ("frame_keys", C.c_void_p),
        ("fetch", C.c_void_p), ("has", C.c_void_p),
        ("nblobs", C.c_void_p), ("blob_group", C.c_void_p), ("blob_stats", C.c_void_p),
        ("chain_key", C.c_void_p), ("chain_width", C.c_void_p), ("chain_words", C.c_void_p),
        ("motion", C.c_uint32), ("fading", C.c_uint32), ("density", C.c_uint32), ("feather", C.c_uint32),
        ("show_blobs", C.c_uint32), ("keep_background", C.c_uint32), ("blob_delimiter", C.c_uint32), ("seed", C.c_uint32),
    ]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError("oracle/libamoracle.so missing: run `make -C oracle port`")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, f64, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_double, C.c_int

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("amo_rgb_to_hsp", u32, u32)
    sig("amo_hsp_to_rgb", u32, u32)
    sig("amo_color_distance", f64, u32, u32)
    sig("amo_point_distance", u64, u64, u64)
    sig("amo_octave_noise", f64, C.c_uint, f64, f64, i32)
    sig("amo_perlin_table", None, C.c_uint, vp)
    sig("amo_spline_point", None, u64, vp, vp, f64, vp)
    sig("amo_interpolate_color", u32, u32, u32, f64, f64, f64, i32)
    sig("amo_interpolate_point", u64, u64, u64, f64)
    sig("amo_cost", f64, u64, vp, u64, vp)
    sig("amo_morph_steps", None, vp, u64, u64, u64, u64, vp, vp)
    sig("amo_blobify", u64, u32, u32, vp, vp, vp, vp, u64)
    sig("amo_blob_distance", f64, f64, vp, f64, vp, f64, f64, f64, u32)
    sig("amo_render", None, C.POINTER(Scene), f64, vp)
    sig("amo_background", u32, C.POINTER(Scene), i32, i32, f64)
    sig("amo_fluid_step", None, u32, u32, u32, vp, vp, u64, f64)
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ptr_array(arrays):
    arr = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    return arr


# ---- pure functions
def rgb_to_hsp(c): return lib().amo_rgb_to_hsp(int(c))
def hsp_to_rgb(c): return lib().amo_hsp_to_rgb(int(c))
def color_distance(a, b): return lib().amo_color_distance(int(a), int(b))
def point_distance(a, b): return lib().amo_point_distance(int(a), int(b))
def octave_noise(seed, x, y, octaves=8): return lib().amo_octave_noise(int(seed), float(x), float(y), int(octaves))


def perlin_table(seed):
    out = np.zeros(512, dtype=np.int32)
    lib().amo_perlin_table(int(seed), _p(out))
    return out


def spline_point(xs, ys, t):
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    ys = np.ascontiguousarray(ys, dtype=np.float64)
    out = np.zeros(2)
    lib().amo_spline_point(len(xs), _p(xs), _p(ys), float(t), _p(out))
    return out


def interpolate_color(c1, c2, w, lag=None, slope=None):
    if lag is None:
        return lib().amo_interpolate_color(int(c1), int(c2), 0.0, 0.0, float(w), 0)
    return lib().amo_interpolate_color(int(c1), int(c2), float(lag), float(slope), float(w), 1)


def interpolate_point(p1, p2, w): return lib().amo_interpolate_point(int(p1), int(p2), float(w))


# ---- matcher
def cost(chains):
    widths = np.array([c["words"].shape[1] for c in chains], dtype=np.uint64)
    h = chains[0]["words"].shape[0]
    words = np.ascontiguousarray(np.concatenate([np.ascontiguousarray(c["words"], dtype=np.uint64).reshape(-1) for c in chains]))
    return lib().amo_cost(len(chains), _p(widths), h, _p(words))


def morph_steps(words, steps, cycle_length, e1_state):
    """Serial reference matcher on one chain (threads = 0).  -> (new words, new e1 state, gain)."""
    w = np.ascontiguousarray(words, dtype=np.uint64).copy()
    st = C.c_uint64(int(e1_state))
    gain = C.c_double(0)
    lib().amo_morph_steps(_p(w), w.shape[1], w.shape[0], int(steps), int(cycle_length), C.byref(st), C.byref(gain))
    return w, int(st.value), float(gain.value)


# ---- blobs
def blobify(present, stored):
    present = np.ascontiguousarray(present, dtype=np.uint8)
    stored = np.ascontiguousarray(stored, dtype=np.uint32)
    h, w = present.shape
    labels = np.zeros((h, w), dtype=np.int64)
    cap = int(present.sum()) + 1
    stats = np.zeros((cap, 7))
    n = lib().amo_blobify(w, h, _p(present), _p(stored), _p(labels), _p(stats), cap)
    return labels, stats[:n]


def blob_distance(sz1, s1, sz2, s2, w_xy, w_rgba, w_size, bbox_d):
    s1 = np.ascontiguousarray(s1, dtype=np.float64)
    s2 = np.ascontiguousarray(s2, dtype=np.float64)
    return lib().amo_blob_distance(float(sz1), _p(s1), float(sz2), _p(s2), float(w_xy), float(w_rgba), float(w_size), int(bbox_d))


# ---- renderer
class RenderScene:
    """Plain-array scene for amo_render.  fetch/has: lists of (ch, cw) arrays; blobs: per frame list of
    dict(group, stats[6]); chains: list of dict(key, words (h, w))."""

    def __init__(self, width, height, bbox, frame_keys, fetch, has, blobs, chains, motion, fading, density=1, feather=0,
                 show_blobs=0, keep_background=0, blob_delimiter=1, seed=0):
        self._keep = []
        ch, cw = fetch[0].shape
        S = Scene()
        S.width, S.height, S.cw, S.ch = width, height, cw, ch
        S.nframes, S.nchains = len(fetch), len(chains)
        for i in range(4):
            S.bbox[i] = int(bbox[i])

        def keep(a):
            self._keep.append(a)
            return a

        keys = keep(np.ascontiguousarray(frame_keys, dtype=np.uint64))
        S.frame_keys = keys.ctypes.data
        f_arrays = [keep(np.ascontiguousarray(f, dtype=np.uint32)) for f in fetch]
        h_arrays = [keep(np.ascontiguousarray(h, dtype=np.uint8)) for h in has]
        S.fetch = C.cast(keep(_ptr_array(f_arrays)), C.c_void_p)
        S.has = C.cast(keep(_ptr_array(h_arrays)), C.c_void_p)
        nb = keep(np.array([len(b) for b in blobs], dtype=np.uint32))
        S.nblobs = nb.ctypes.data
        g_arrays = [keep(np.array([x["group"] for x in b] + [0], dtype=np.uint64)) for b in blobs]
        s_arrays = [keep(np.ascontiguousarray(np.array([x["stats"] for x in b] + [np.zeros(6)], dtype=np.float64).reshape(-1))) for b in blobs]
        S.blob_group = C.cast(keep(_ptr_array(g_arrays)), C.c_void_p)
        S.blob_stats = C.cast(keep(_ptr_array(s_arrays)), C.c_void_p)
        ck = keep(np.array([c["key"] for c in chains] + [0], dtype=np.uint64))
        cwid = keep(np.array([c["words"].shape[1] for c in chains] + [0], dtype=np.uint64))
        w_arrays = [keep(np.ascontiguousarray(c["words"], dtype=np.uint64)) for c in chains]
        S.chain_key, S.chain_width = ck.ctypes.data, cwid.ctypes.data
        S.chain_words = C.cast(keep(_ptr_array(w_arrays)), C.c_void_p) if chains else None
        S.motion, S.fading, S.density, S.feather = int(motion), int(fading), int(density), int(feather)
        S.show_blobs, S.keep_background, S.blob_delimiter, S.seed = int(show_blobs), int(keep_background), int(blob_delimiter), int(seed)
        self.S = S
        self.width, self.height = width, height

    def render(self, t):
        out = np.zeros((self.height, self.width), dtype=np.uint32)
        lib().amo_render(C.byref(self.S), float(t), _p(out))
        return out

    def background(self, x, y, t):
        return lib().amo_background(C.byref(self.S), int(x), int(y), float(t))


# ---- fluid
def fluid_step(gx, gy, rec, steps_left, freedom_radius):
    rec = np.ascontiguousarray(rec, dtype=np.float64).copy()
    nodes = np.zeros((gy, gx, 13))
    lib().amo_fluid_step(gx, gy, rec.shape[0], _p(rec), _p(nodes), int(steps_left), float(freedom_radius))
    return rec, nodes
