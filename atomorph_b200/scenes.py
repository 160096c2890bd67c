"""Seeded synthetic key-frame generators.  numpy only.

Every generator returns a list of (H, W, 4) uint8 RGBA images, one per key frame.  A pixel is
"present" iff alpha != 0, exactly as the reference CLI ingests PNGs (demo/main.cpp:96-127).
The same arrays feed the reference harness (oracle side) and the CUDA library (product side).
"""
import numpy as np


def _texture(h, w, rng, phase=0.0):
    """Smooth gradients + uniform noise, opaque."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    r = 255.0 * xx / max(w - 1, 1)
    g = 255.0 * yy / max(h - 1, 1)
    b = 127.5 * (1.0 + np.sin(0.02 * (xx + yy) + phase))
    img = np.stack([r, g, b], axis=-1) * 0.8 + rng.uniform(0.0, 51.0, size=(h, w, 3))
    out = np.zeros((h, w, 4), dtype=np.uint8)
    out[..., :3] = np.clip(np.round(img), 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out


def square_to_disc(n=1024, seed=1234):
    """BASELINE.json config 2: frame 0 = full n x n square, frame 1 = centred disc of radius n/2.
    W = n*n atoms with duplicates in the disc column (823 471 px at n = 1024)."""
    rng = np.random.default_rng(seed)
    f0 = _texture(n, n, rng, 0.0)
    f1 = _texture(n, n, rng, 1.0)
    yy, xx = np.mgrid[0:n, 0:n]
    c = (n - 1) / 2.0
    inside = (xx - c) ** 2 + (yy - c) ** 2 <= (n / 2.0) ** 2
    f1[~inside] = 0
    return [f0, f1]


def ellipses(n=64, frames=2, seed=7, margin=None, alpha_noise=False):
    """Rotating / scaling ellipse per key frame (cyclic), kept away from the canvas edge so that
    Catmull-Rom overshoot stays inside the image (SURVEY.md section 9 note 4)."""
    rng = np.random.default_rng(seed)
    if margin is None:
        margin = max(4, n // 6)
    out = []
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float64)
    c = (n - 1) / 2.0
    for k in range(frames):
        ang = np.pi * k / max(frames, 1)
        a = (n / 2.0 - margin) * (0.75 + 0.25 * np.cos(2 * np.pi * k / frames))
        b = (n / 2.0 - margin) * (0.55 + 0.2 * np.sin(2 * np.pi * k / frames + 0.3))
        xr = (xx - c) * np.cos(ang) + (yy - c) * np.sin(ang)
        yr = -(xx - c) * np.sin(ang) + (yy - c) * np.cos(ang)
        inside = (xr / a) ** 2 + (yr / b) ** 2 <= 1.0
        img = _texture(n, n, rng, 0.7 * k)
        if alpha_noise:
            img[..., 3] = rng.integers(40, 256, size=(n, n)).astype(np.uint8)
        img[~inside] = 0
        out.append(img)
    return out


def rect_blobs(n=512, count=2500, frames=2, seed=11, min_side=2, max_side=20):
    """BASELINE.json config 4: `count` disjoint flat-coloured rectangles per key frame with a
    >= 1 px transparent gap, a different layout per frame.  With blob_threshold = 1.0 the
    reference's partition equals the 4-connected components."""
    out = []
    for k in range(frames):
        rng = np.random.default_rng(seed + 1000 * k)
        img = np.zeros((n, n, 4), dtype=np.uint8)
        occ = np.zeros((n + 2, n + 2), dtype=bool)
        placed = 0
        tries = 0
        while placed < count and tries < count * 200:
            tries += 1
            w = int(rng.integers(min_side, max_side + 1))
            h = int(rng.integers(min_side, max_side + 1))
            x = int(rng.integers(1, n - w - 1))
            y = int(rng.integers(1, n - h - 1))
            if occ[y:y + h + 2, x:x + w + 2].any():
                continue
            occ[y + 1:y + h + 1, x + 1:x + w + 1] = True
            col = rng.integers(0, 256, size=3)
            img[y:y + h, x:x + w, :3] = col
            img[y:y + h, x:x + w, 3] = 255
            placed += 1
        out.append(img)
    return out


def random_cloud(n=48, frames=3, fill=0.6, seed=3, margin=6):
    """Sparse random semi-transparent pixel clouds of unequal size: exercises duplicates,
    one-sided atoms and alpha handling in the renderer."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(frames):
        img = np.zeros((n, n, 4), dtype=np.uint8)
        m = rng.uniform(size=(n, n)) < (fill * (0.6 + 0.4 * rng.uniform()))
        m[:margin] = m[-margin:] = False
        m[:, :margin] = m[:, -margin:] = False
        img[..., :3] = rng.integers(0, 256, size=(n, n, 3))
        img[..., 3] = rng.integers(1, 256, size=(n, n))
        img[~m] = 0
        out.append(img)
    return out


def rotating_shapes(n=4096, frames=8, seed=21, coverage=0.85):
    """BASELINE.json config 5: `frames` cyclic key frames of one shape (a rounded super-ellipse) that rotates and
    scales per key frame and covers >= 80 % of the n x n canvas; textured RGBA."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float32)
    c = (n - 1) / 2.0
    out = []
    for k in range(frames):
        ang = np.pi * k / frames
        s = 1.0 - 0.06 * (1 + np.cos(2 * np.pi * k / frames)) / 2.0
        xr = ((xx - c) * np.cos(ang) + (yy - c) * np.sin(ang)) / (c * s)
        yr = (-(xx - c) * np.sin(ang) + (yy - c) * np.cos(ang)) / (c * s)
        inside = (np.abs(xr) ** 6 + np.abs(yr) ** 6) <= 1.0          # rounded square: ~ 93 % of its bounding box
        img = _texture(n, n, rng, 0.5 * k)
        img[~inside] = 0
        out.append(img)
    return out
