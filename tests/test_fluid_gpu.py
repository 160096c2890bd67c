"""K5 parity: one FluidModel::step from identical state, fields within 1e-5 relative
(SURVEY.md row a-F, hard part 5)."""
import numpy as np
import pytest

from atomorph_b200 import engine as eng

pytestmark = pytest.mark.gpu


def _particles(n, gx, gy, seed, immature_frac=0.2, inactive_frac=0.1):
    rng = np.random.default_rng(seed)
    rec = np.zeros((n, eng.FP_STRIDE))
    rec[:, 0] = rng.uniform(12, gx - 12, n)
    rec[:, 1] = rng.uniform(12, gy - 12, n)
    rec[:, 2] = rng.normal(0, 0.2, n)
    rec[:, 3] = rng.normal(0, 0.2, n)
    rec[:, 4] = np.clip(rec[:, 0] + rng.normal(0, 3, n), 1, gx - 2)
    rec[:, 5] = np.clip(rec[:, 1] + rng.normal(0, 3, n), 1, gy - 2)
    rec[:, 6] = 1.0
    rec[:, 7] = rng.uniform(size=n) > inactive_frac
    rec[:, 8] = rng.uniform(size=n) > immature_frac
    rec[:, 9:13] = rng.uniform(size=(n, 4))
    rec[:, 13:17] = rng.uniform(size=(n, 4))
    rec[:, 17] = rng.choice([1.0, 0.1], size=n, p=[0.8, 0.2])
    return rec


@pytest.mark.parametrize("n,gx,gy,seed", [(500, 52, 52, 1), (4000, 84, 70, 2)])
def test_single_step_matches_reference(reflib, n, gx, gy, seed):
    rec = _particles(n, gx, gy, seed)
    rf = reflib.RefFluid(gx, gy, n)
    rf.set_particles(rec)
    e = eng.Engine(0)
    e.fluid_create(gx, gy, n)
    e.fluid_set_particles(rec)
    for step, (steps_left, radius) in enumerate([(3, 12.0), (2, 8.0), (0, 1.5)]):
        rf.step(steps_left, radius, 0.3)
        e.fluid_step(steps_left, radius, 0.3)
        a, b = rf.get_particles(), e.fluid_get_particles()
        act = a[:, 7] != 0
        for col in (0, 1, 2, 3, 13, 14, 15, 16):
            ref, got = a[act, col], b[act, col]
            scale = np.maximum(np.abs(ref), 1e-3 if col in (2, 3) else 1.0)
            assert np.max(np.abs(ref - got) / scale) < 1e-5, (step, col)
        # untouched fields
        assert np.array_equal(a[~act, :2], b[~act, :2])
        na, nb = rf.nodes(), e.fluid_nodes()
        for k in range(13):
            scale = max(np.abs(na[..., k]).max(), 1e-9)
            assert np.max(np.abs(na[..., k] - nb[..., k])) / scale < 1e-5, (step, "node", k)
    rf.close()
