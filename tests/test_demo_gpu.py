"""BASELINE.json config 1 through the reference's own CLI: demo/main.cpp + lodepng, compiled UNCHANGED, once against
the reference library (oracle/_ref/demo_ref, the CPU arm) and once against include/atomorph/*.h + libatomorph_b200.so
(oracle/_ref/demo_b200, the drop-in).  Both binaries are built in the build container by `make -C oracle demo`
(atomorph_b200/build.py) and travel to the GPU box as built artefacts.

The CLI drives the morph by wall clock (-M / -O seconds), so the atom matching differs between two runs of ANY
implementation; what is comparable: the first output frame (t = 0: every atom sits on its own source pixel -- the key
frame with MORE pixels comes first and the density is 1, so there are no duplicates at t = 0) must be bit-identical, and
the later frames must be morphs of the same shapes (coverage and mean colour close to the reference run's)."""
import os
import subprocess

import numpy as np
import pytest

from atomorph_b200 import scenes

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO_B200 = os.path.join(ROOT, "oracle", "_ref", "demo_b200")
DEMO_REF = os.path.join(ROOT, "oracle", "_ref", "demo_ref")
FRAMES = 8


def _run_cli(exe, indir, outdir, prefix, names, extra):
    cmd = [exe, "-f", prefix, "-i", str(indir), "-o", str(outdir)] + names + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    from PIL import Image
    out = []
    for f in range(FRAMES):
        p = os.path.join(str(outdir), "%s_%04d.png" % (prefix, f + 1))
        assert os.path.exists(p), (p, r.stdout[-1500:], r.stderr[-1500:])
        out.append(np.array(Image.open(p).convert("RGBA")))
    return out


@pytest.mark.skipif(not (os.path.exists(DEMO_B200) and os.path.exists(DEMO_REF)), reason="demo binaries not built (make -C oracle demo)")
@pytest.mark.parametrize("flags", [["--motion-linear", "--fading-linear"], ["--motion-spline", "--fading-cosine"],
                                   ["--motion-linear", "--fading-linear", "-b", "6", "-p", "3", "-c", "2", "-z", "1"]],
                         ids=["linear", "spline_cosine", "six_blobs"])
def test_reference_cli_runs_unchanged_on_the_device(tmp_path, flags):
    from PIL import Image
    images = scenes.ellipses(64, 2, seed=3)
    if (images[0][..., 3] > 0).sum() < (images[1][..., 3] > 0).sum():
        images = images[::-1]                     # more pixels first: no duplicate atoms at t = 0
    names = []
    for k, im in enumerate(images):
        names.append("key_%d.png" % k)
        Image.fromarray(im, "RGBA").save(os.path.join(str(tmp_path), names[-1]))
    extra = ["-F", str(FRAMES), "-D", "1", "-M", "1", "-O", "2", "-s", "1", "-T", "4", "--finite"] + flags
    ours = _run_cli(DEMO_B200, tmp_path, tmp_path, "b200", names, extra)
    ref = _run_cli(DEMO_REF, tmp_path, tmp_path, "ref", names, extra)
    assert ours[0].shape == ref[0].shape == (64, 64, 4)
    several_blobs = "-b" in flags
    if not several_blobs:
        # t = 0: independent of the matching -> bit-identical, and it is the first key frame after the 8-bit HSP round trip
        assert np.array_equal(ours[0], ref[0])
        assert np.array_equal(ours[0][..., 3] > 0, images[0][..., 3] > 0)
    # (with several blobs per key frame the blob partition itself is RNG-dependent in the reference, and a blob that is
    # smaller than its partner carries duplicate atoms at t = 0: all frames are compared statistically)
    for f in range(0 if several_blobs else 1, FRAMES):
        a, b = ours[f].astype(np.float64), ref[f].astype(np.float64)
        ca, cb = (a[..., 3] > 0).sum(), (b[..., 3] > 0).sum()
        # (different blob partitions move the parts of the shape differently: looser bounds with several blobs)
        assert cb > 0 and abs(ca - cb) <= (0.35 if several_blobs else 0.15) * cb, (f, ca, cb)
        ma, mb = a[a[..., 3] > 0][:, :3].mean(axis=0), b[b[..., 3] > 0][:, :3].mean(axis=0)
        assert np.abs(ma - mb).max() <= (25.0 if several_blobs else 12.0), (f, ma, mb)
