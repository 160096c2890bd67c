"""Shared test plumbing: run the reference on a scene, dump its state, feed it to the engine."""
import numpy as np

from atomorph_b200 import engine as eng


def build_ref(amref, images, width=None, height=None, target=None, match_steps=0, **params):
    """Reference morph driven deterministically to `target` state (default: ATOM_MORPHING)."""
    H, W = images[0].shape[:2]
    m = amref.RefMorph(**params)
    for k, im in enumerate(images):
        m.add_image(k, im)
    m.set_resolution(W if width is None else width, H if height is None else height)
    m.run_until(amref.STATE_ATOM_MORPHING if target is None else target, match_steps=match_steps)
    return m


def ref_blob_tables(m, key):
    """labels (H, W) int32, stats (n, 6), groups (n,) of one reference frame."""
    blobs = m.blobs(key)
    labels = m.blob_labels(key)
    stats = np.array([b["stats"] for b in blobs]).reshape(-1, 6)
    groups = np.array([b["group"] for b in blobs], dtype=np.uint64)
    return labels, stats, groups


def engine_from_ref(m, images, device=0, **params):
    """Engine holding the reference's own blobs + chain tables (render / cost parity set-up)."""
    e = eng.Engine(device, **params)
    e.load_images(images, width=m.width, height=m.height)
    for i, key in enumerate(m.frame_keys()):
        labels, stats, groups = ref_blob_tables(m, key)
        e.import_blobs(i, labels, stats, groups)
    e.import_chains(m.chains())
    return e


def diff_stats(a, b):
    """a, b: uint32 packed images.  -> (#pixels differing, max abs channel diff)."""
    ua = eng.unpack_rgba(a).astype(np.int32)
    ub = eng.unpack_rgba(b).astype(np.int32)
    d = np.abs(ua - ub)
    return int((d.max(axis=-1) > 0).sum()), int(d.max())
