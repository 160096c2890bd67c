"""Thin Python view of one device engine (include/amx.h).  numpy in, numpy out.

This is plumbing for tests, bench.py and the multi-GPU driver -- all compute happens in
libatomorph_b200.so.  Method names follow the C-ABI one to one.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import AmxError

# am:: constants (reference atomorph.h:235-246, 304-306)
RGB, HSP, NONE, LINEAR, SPLINE, COSINE, PERLIN = 0, 1, 2, 3, 4, 5, 6
STATE_BLOB_DETECTION, STATE_BLOB_UNIFICATION, STATE_BLOB_MATCHING, STATE_ATOM_MORPHING, STATE_DONE = 0, 1, 2, 3, 4
TEXTURE, AVERAGE, DISTINCT = 0, 1, 2
HAS_PIXEL, HAS_FLUID = 1, 2
FP_STRIDE = 24

PARAM = dict(blob_delimiter=0, blob_threshold=1, blob_max_size=2, blob_min_size=3, blob_box_grip=4,
             blob_box_samples=5, blob_number=6, blob_rgba_weight=7, blob_size_weight=8, blob_xy_weight=9,
             degeneration=10, density=11, motion=12, fading=13, threads=14, cycle_length=15, feather=16,
             keep_background=17, finite=18, show_blobs=19, fluid=20, seed=21)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pack_rgba(rgba_u8):
    a = np.ascontiguousarray(rgba_u8, dtype=np.uint8)
    return a.view(np.uint32).reshape(a.shape[:-1])


def unpack_rgba(u32):
    a = np.ascontiguousarray(u32, dtype=np.uint32)
    return a.view(np.uint8).reshape(a.shape + (4,))


def frame_means(rgba, present, hsp=True):
    """Running means x,y,r,g,b,a of a frame exactly as morph::add_pixel accumulates them
    (reference morph.cpp:318-334): pixels visited in row-major order, weight 1/n."""
    # The engine only needs these for volatile-blob padding; a plain mean is equivalent to ~1e-13.
    ys, xs = np.nonzero(present)
    if len(xs) == 0:
        return np.zeros(6)
    c = rgba[ys, xs].astype(np.float64) / 255.0
    return np.array([xs.mean(), ys.mean(), c[:, 0].mean(), c[:, 1].mean(), c[:, 2].mean(), c[:, 3].mean()])


class Engine:
    def __init__(self, device=0, **params):
        self.L = _lib.lib()
        h = C.c_void_p()
        rc = self.L.amx_create(C.byref(h), int(device))
        if rc != 0:
            raise AmxError("amx_create failed (%s): a CUDA device is required, there is no CPU fallback"
                           % _lib.STATUS.get(rc, rc))
        self.h = h
        self.width = self.height = self.cw = self.ch = 0
        self.nframes = 0
        if params:
            self.set(**params)

    # ---- plumbing
    def _ck(self, rc, what):
        if rc != 0:
            msg = self.L.amx_last_error(self.h)
            raise AmxError("%s: %s (%s)" % (what, _lib.STATUS.get(rc, rc), msg.decode() if msg else ""))

    def close(self):
        if getattr(self, "h", None):
            self.L.amx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.amx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "set_stream")

    def get_stream(self):
        return self.L.amx_get_stream(self.h) or 0

    def sync(self):
        self._ck(self.L.amx_device_sync(self.h), "sync")

    def set(self, **params):
        for k, v in params.items():
            self._ck(self.L.amx_set_param(self.h, PARAM[k], float(v)), "set_param " + k)

    def get(self, name):
        return self.L.amx_get_param(self.h, PARAM[name])

    def launch_count(self):
        return int(self.L.amx_launch_count(self.h))

    def timer_start(self):
        self._ck(self.L.amx_timer_start(self.h), "timer_start")

    def timer_stop(self):
        ms = C.c_float(0)
        self._ck(self.L.amx_timer_stop(self.h, C.byref(ms)), "timer_stop")
        return float(ms.value)

    # ---- ingest
    def load_images(self, images, width=None, height=None, keys=None, presents=None):
        """images: list of (H, W, 4) uint8.  A pixel is present iff alpha != 0 (demo/main.cpp:96-127)
        unless `presents` gives explicit masks."""
        H, W = images[0].shape[:2]
        width = W if width is None else width
        height = H if height is None else height
        if presents is None:
            presents = [im[..., 3] != 0 for im in images]
        anyp = np.zeros((H, W), dtype=bool)
        for p in presents:
            anyp |= p
        ys, xs = np.nonzero(anyp)
        if len(xs):
            bbox = np.array([xs.min(), ys.min(), xs.max(), ys.max()], dtype=np.uint16)
        else:
            bbox = np.array([65535, 65535, 0, 0], dtype=np.uint16)
        self._ck(self.L.amx_reset(self.h), "reset")
        self._ck(self.L.amx_set_canvas(self.h, width, height, W, H, _p(bbox)), "set_canvas")
        self.width, self.height, self.cw, self.ch = width, height, max(W, width), max(H, height)
        if (self.cw, self.ch) != (W, H):
            raise AmxError("images must cover the canvas")
        keys = np.arange(len(images), dtype=np.uint64) if keys is None else np.asarray(keys, dtype=np.uint64)
        self._ck(self.L.amx_set_frame_count(self.h, len(images), _p(keys)), "set_frame_count")
        self.nframes = len(images)
        self.bbox = tuple(int(v) for v in bbox)
        for i, (im, pr) in enumerate(zip(images, presents)):
            rgba = np.ascontiguousarray(pack_rgba(im))
            pres = np.ascontiguousarray(pr.astype(np.uint8))
            means = frame_means(im, pr)
            if not pr.any() and len(xs):
                means[0] = (int(bbox[0]) + int(bbox[2])) / 2.0   # morph.cpp:276-283 empty key frame
                means[1] = (int(bbox[1]) + int(bbox[3])) / 2.0
            self._ck(self.L.amx_upload_frame(self.h, i, _p(rgba), _p(pres), _p(means)), "upload_frame")

    def fetch_image(self, index):
        out = np.zeros((self.ch, self.cw), dtype=np.uint32)
        self._ck(self.L.amx_download_fetch(self.h, index, _p(out)), "download_fetch")
        return out

    def stored_image(self, index):
        out = np.zeros((self.ch, self.cw), dtype=np.uint32)
        self._ck(self.L.amx_download_stored(self.h, index, _p(out)), "download_stored")
        return out

    # ---- pipeline
    def step(self, n=1):
        self._ck(self.L.amx_step(self.h, int(n)), "step")

    def next_state(self):
        self.L.amx_next_state(self.h)

    def state(self):
        return int(self.L.amx_get_state(self.h))

    def energy(self):
        return float(self.L.amx_get_energy(self.h))

    # ---- blobs
    def blobify(self):
        self._ck(self.L.amx_blobify(self.h), "blobify")

    def blob_count(self, index):
        n = C.c_uint32(0)
        self._ck(self.L.amx_blob_count(self.h, index, C.byref(n)), "blob_count")
        return int(n.value)

    def export_blobs(self, index):
        n = self.blob_count(index)
        labels = np.zeros((self.ch, self.cw), dtype=np.int32)
        stats = np.zeros((n, 6))
        meta = np.zeros((n, 2), dtype=np.uint64)
        self._ck(self.L.amx_export_blobs(self.h, index, _p(labels), _p(stats), _p(meta)), "export_blobs")
        return labels, stats, meta

    def import_blobs(self, index, labels, stats, groups):
        stats = np.ascontiguousarray(stats, dtype=np.float64).reshape(-1, 6)
        groups = np.ascontiguousarray(groups, dtype=np.uint64)
        lab = None if labels is None else np.ascontiguousarray(labels, dtype=np.int32)
        self._ck(self.L.amx_import_blobs(self.h, index, len(groups), None if lab is None else _p(lab), _p(stats), _p(groups)),
                 "import_blobs")

    def match_init(self):
        self._ck(self.L.amx_match_init(self.h), "match_init")

    def match_rounds(self, n):
        self._ck(self.L.amx_match_rounds(self.h, int(n)), "match_rounds")

    def match_energy(self):
        e = C.c_double(0)
        self._ck(self.L.amx_match_energy(self.h, C.byref(e)), "match_energy")
        return float(e.value)

    # ---- chains
    def init_chains(self):
        self._ck(self.L.amx_init_chains(self.h), "init_chains")

    def chain_count(self):
        n = C.c_uint32(0)
        self._ck(self.L.amx_chain_count(self.h, C.byref(n)), "chain_count")
        return int(n.value)

    def chains(self):
        out = []
        for c in range(self.chain_count()):
            info = np.zeros(4, dtype=np.uint64)
            self._ck(self.L.amx_chain_info(self.h, c, _p(info)), "chain_info")
            key, w, h, ms = (int(v) for v in info)
            words = np.zeros((h, w), dtype=np.uint64)
            if w * h:
                self._ck(self.L.amx_export_chain(self.h, c, _p(words)), "export_chain")
            out.append(dict(key=key, width=w, height=h, max_surface=ms, words=words))
        return out

    def import_chains(self, chains):
        """chains: list of dict(key, words (h, w) uint64, max_surface)."""
        n = len(chains)
        keys = np.array([c["key"] for c in chains], dtype=np.uint64)
        widths = np.array([c["words"].shape[1] for c in chains], dtype=np.uint64)
        ms = np.array([c.get("max_surface", c["words"].shape[1]) for c in chains], dtype=np.uint64)
        h = chains[0]["words"].shape[0] if n else 0
        if n == 1 and chains[0]["words"].dtype == np.uint64 and chains[0]["words"].flags["C_CONTIGUOUS"]:
            words = chains[0]["words"].reshape(-1)               # no host copy (the caller's buffer may be pinned)
        else:
            words = (np.concatenate([np.ascontiguousarray(c["words"], dtype=np.uint64).reshape(-1) for c in chains])
                     if n else np.zeros(0, dtype=np.uint64))
            words = np.ascontiguousarray(words)
        self._ck(self.L.amx_import_chains(self.h, n, _p(keys), _p(widths), _p(ms), h, _p(words)), "import_chains")

    def table_device_ptr(self, column):
        p = C.c_void_p()
        n = C.c_uint64(0)
        self._ck(self.L.amx_table_device_ptr(self.h, column, C.byref(p), C.byref(n)), "table_device_ptr")
        return p.value, int(n.value)

    # ---- K1
    def swap_rounds(self, rounds, chain=-1, column=-1, want_stats=True):
        st = np.zeros(3, dtype=np.uint64)
        self._ck(self.L.amx_swap_rounds(self.h, chain, column, int(rounds), _p(st) if want_stats else None), "swap_rounds")
        return st

    def swap_rounds_sharded(self, rounds, sel_mask, sel_val, chain=0, column=-1):
        self._ck(self.L.amx_swap_rounds_sharded(self.h, chain, column, int(rounds), int(sel_mask), int(sel_val)), "swap_rounds_sharded")

    def pack_owned(self, column, sel_mask, sel_val, d_out_ptr, chain=0):
        n = C.c_uint64(0)
        self._ck(self.L.amx_pack_owned(self.h, chain, column, int(sel_mask), int(sel_val), C.c_void_p(d_out_ptr), C.byref(n)), "pack_owned")
        return int(n.value)

    def unpack_owned(self, column, sel_mask, nranks, d_in_ptr, chain=0):
        self._ck(self.L.amx_unpack_owned(self.h, chain, column, int(sel_mask), int(nranks), C.c_void_p(d_in_ptr)), "unpack_owned")

    def swap_tiled_epoch(self, epoch, rounds, column, rank=0, nranks=1, chain=0):
        self._ck(self.L.amx_swap_tiled_epoch(self.h, chain, column, int(epoch), int(rounds), int(rank), int(nranks)), "swap_tiled_epoch")

    def swap_local_epoch(self, epoch, rounds, column, chain=0):
        """One locality epoch: tiles of 1024 spatial neighbours (Morton order of the column's current positions)."""
        self._ck(self.L.amx_swap_local_epoch(self.h, chain, column, int(epoch), int(rounds)), "swap_local_epoch")

    def set_swap_locality(self, every):
        """Every `every`-th epoch of swap_rounds / step pairs spatial neighbours (0 = never, the default)."""
        self._ck(self.L.amx_set_swap_locality(self.h, int(every)), "set_swap_locality")

    def pack_tiled(self, epoch, column, rank, nranks, d_out_ptr, chain=0):
        n = C.c_uint64(0)
        self._ck(self.L.amx_pack_tiled(self.h, chain, column, int(epoch), int(rank), int(nranks), C.c_void_p(d_out_ptr), C.byref(n)), "pack_tiled")
        return int(n.value)

    def unpack_tiled(self, epoch, column, d_in_ptr, chain=0):
        self._ck(self.L.amx_unpack_tiled(self.h, chain, column, int(epoch), C.c_void_p(d_in_ptr)), "unpack_tiled")

    # ---- multi-GPU matcher (amx_dist.cu): every collective is issued by the library on the engine's stream
    @staticmethod
    def comm_unique_id():
        ident = np.zeros(128, dtype=np.uint8)
        rc = _lib.lib().amx_comm_unique_id(_p(ident))
        if rc != 0:
            raise AmxError("amx_comm_unique_id: %s (is libnccl.so.2 loadable?)" % _lib.STATUS.get(rc, rc))
        return ident

    def comm_init(self, ident, rank, nranks):
        ident = np.ascontiguousarray(ident, dtype=np.uint8)
        assert ident.size == 128
        self._ck(self.L.amx_comm_init(self.h, _p(ident), int(rank), int(nranks)), "comm_init")

    def comm_destroy(self):
        self._ck(self.L.amx_comm_destroy(self.h), "comm_destroy")

    def comm_enable_p2p(self):
        """Collective.  True when every rank mapped every replica (write-through exchange), False when the box does not
        allow peer mapping (the NCCL exchange stays in use)."""
        rc = self.L.amx_comm_enable_p2p(self.h)
        if rc == 3:      # AMX_ERR_STATE
            return False
        self._ck(rc, "comm_enable_p2p")
        return True

    def comm_disable_p2p(self):
        self._ck(self.L.amx_comm_disable_p2p(self.h), "comm_disable_p2p")

    def comm_info(self):
        out = np.zeros(3, dtype=np.uint32)
        self._ck(self.L.amx_comm_info(self.h, _p(out)), "comm_info")
        return dict(rank=int(out[0]), nranks=int(out[1]), p2p=bool(out[2]))

    def comm_check(self):
        self._ck(self.L.amx_comm_check(self.h), "comm_check")

    def table_broadcast(self, root=0):
        self._ck(self.L.amx_table_broadcast(self.h, int(root)), "table_broadcast")

    def swap_part_step(self, step, sub_epochs, rounds, column=1, chain=0):
        self._ck(self.L.amx_swap_part_step(self.h, int(chain), int(column), int(step), int(sub_epochs), int(rounds)), "swap_part_step")

    def swap_columns_step(self, phase, step, epochs, rounds, chain=-1):
        self._ck(self.L.amx_swap_columns_step(self.h, int(chain), int(phase), int(step), int(epochs), int(rounds)), "swap_columns_step")

    def swap_phase_count(self):
        return int(self.L.amx_swap_phase_count(self.h))

    def column_hash(self, column):
        """(position-dependent hash, multiset hash) of one key-frame column of the table."""
        out = np.zeros(2, dtype=np.uint64)
        self._ck(self.L.amx_column_hash(self.h, int(column), _p(out)), "column_hash")
        return int(out[0]), int(out[1])

    def swap_stats(self):
        st = np.zeros(3, dtype=np.uint64)
        self._ck(self.L.amx_swap_stats(self.h, _p(st)), "swap_stats")
        return st

    def cost(self):
        c = C.c_double(0)
        self._ck(self.L.amx_cost(self.h, C.byref(c)), "cost")
        return float(c.value)

    # ---- K6
    def render_prepare(self):
        self._ck(self.L.amx_render_prepare(self.h), "render_prepare")

    def render(self, times):
        times = np.ascontiguousarray(np.atleast_1d(times), dtype=np.float64)
        out = np.zeros((len(times), self.height, self.width), dtype=np.uint32)
        self._ck(self.L.amx_render(self.h, _p(times), len(times), _p(out), 0), "render")
        return out

    def render_into(self, times, out_ptr, is_device):
        times = np.ascontiguousarray(np.atleast_1d(times), dtype=np.float64)
        self._ck(self.L.amx_render(self.h, _p(times), len(times), C.c_void_p(out_ptr), 1 if is_device else 0), "render")

    def render_pixels(self, t):
        """One frame as width*height am::pixel records (u16 x, u16 y, r, g, b, a), what morph::get_pixels(t) hands out."""
        out = np.zeros((self.height, self.width), dtype=np.uint64)
        self._ck(self.L.amx_render_pixels(self.h, float(t), _p(out)), "render_pixels")
        return out

    def render_pixels_into(self, t, out_ptr):
        self._ck(self.L.amx_render_pixels(self.h, float(t), C.c_void_p(out_ptr)), "render_pixels")

    def set_lookahead(self, enable):
        self._ck(self.L.amx_set_lookahead(self.h, 1 if enable else 0), "set_lookahead")

    def lookahead_stats(self):
        st = np.zeros(2, dtype=np.uint64)
        self._ck(self.L.amx_lookahead_stats(self.h, _p(st)), "lookahead_stats")
        return dict(hits=int(st[0]), misses=int(st[1]))

    def render_stats(self):
        st = np.zeros(3, dtype=np.uint64)
        self._ck(self.L.amx_render_stats(self.h, _p(st)), "render_stats")
        return dict(generic=int(st[0]), ties=int(st[1]), overflow=int(st[2]))

    def render_path_frames(self):
        """Frames rendered by the tiled path / by the general A-buffer path so far."""
        st = np.zeros(2, dtype=np.uint64)
        self._ck(self.L.amx_render_path_frames(self.h, _p(st)), "render_path_frames")
        return dict(tiled=int(st[0]), general=int(st[1]))

    def kernel_times(self, enable):
        """Per-kernel device times of the render batches since the last call (ms, launches, frames per kernel class:
        0 = k_bin2 / k_scatter, 1 = k_acc (k_tile) / k_gather_pixel); then switches the event recording on or off."""
        ms = np.zeros(2, dtype=np.float64)
        ln = np.zeros(2, dtype=np.uint64)
        fr = np.zeros(2, dtype=np.uint64)
        self._ck(self.L.amx_kernel_times(self.h, 1 if enable else 0, _p(ms), _p(ln), _p(fr)), "kernel_times")
        return [dict(ms=float(ms[k]), launches=int(ln[k]), frames=int(fr[k])) for k in range(2)]

    def render_tiled_stats(self):
        st = np.zeros(8, dtype=np.uint64)
        self._ck(self.L.amx_render_tiled_stats(self.h, _p(st)), "render_tiled_stats")
        return dict(max_bin=[int(v) for v in st[:4]], max_tile=int(st[4]), max_overflow=int(st[5]), fallbacks=int(st[6]), blocked=bool(st[7]))

    def render_blob(self, b, t):
        cap = self.cw * self.ch + 16
        xy = np.zeros((cap, 2), dtype=np.uint16)
        c = np.zeros(cap, dtype=np.uint32)
        n = C.c_int64(0)
        g = C.c_uint64(0)
        self._ck(self.L.amx_render_blob(self.h, b, float(t), cap, _p(xy), _p(c), C.byref(n), C.byref(g)), "render_blob")
        if n.value < 0:
            return None
        return dict(group=int(g.value), xy=xy[:n.value].copy(), rgba=c[:n.value].copy())

    def background(self, t):
        out = np.zeros((self.height, self.width), dtype=np.uint32)
        self._ck(self.L.amx_background(self.h, float(t), _p(out), 0), "background")
        return out

    # ---- K5
    def fluid_create(self, gx, gy, n):
        self._ck(self.L.amx_fluid_create(self.h, gx, gy, n), "fluid_create")
        self._fluid = (gx, gy, n)

    def fluid_set_particles(self, rec):
        rec = np.ascontiguousarray(rec, dtype=np.float64)
        self._ck(self.L.amx_fluid_set_particles(self.h, rec.shape[0], _p(rec)), "fluid_set_particles")

    def fluid_get_particles(self, n=None):
        """Particle records; after a fluid frame was rendered the model belongs to the engine and n = total atoms."""
        rec = np.zeros((self._fluid[2] if n is None else int(n), FP_STRIDE))
        self._ck(self.L.amx_fluid_get_particles(self.h, rec.shape[0], _p(rec)), "fluid_get_particles")
        return rec

    def fluid_step(self, steps_left, freedom_radius, t=0.0):
        self._ck(self.L.amx_fluid_step(self.h, int(steps_left), float(freedom_radius), float(t)), "fluid_step")

    def fluid_nodes(self):
        gx, gy, _ = self._fluid
        out = np.zeros((gy, gx, 13))
        self._ck(self.L.amx_fluid_get_nodes(self.h, _p(out)), "fluid_get_nodes")
        return out
