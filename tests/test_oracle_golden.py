"""The CPU oracle (oracle/am_oracle.cpp) against the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py).  Runs anywhere -- no GPU, no reference tree."""
import numpy as np
import pytest

from oracle import amoracle
import golden_util as gu

pytestmark = pytest.mark.skipif(not amoracle.available(), reason="oracle/libamoracle.so not built")


def test_pure_functions():
    g = gu.load("pure")
    cols = g["cols"]
    assert np.array_equal(np.array([amoracle.rgb_to_hsp(c) for c in cols], dtype=np.uint32), g["hsp"])
    assert np.array_equal(np.array([amoracle.hsp_to_rgb(c) for c in cols], dtype=np.uint32), g["back"])
    assert np.array_equal(np.array([amoracle.color_distance(a, b) for a, b in zip(cols[:-1], cols[1:])]), g["cd"])
    pts = g["pts"]
    assert np.array_equal(np.array([amoracle.point_distance(a, b) for a, b in zip(pts[:-1], pts[1:])], dtype=np.uint64), g["pd"])
    for s, row in zip(g["seeds"], g["noise"]):
        got = np.array([amoracle.octave_noise(s, x, y, 8) for x, y in g["nxy"]])
        assert np.array_equal(got, row)
    for i in range(len(g["sp_n"])):
        n = int(g["sp_n"][i])
        out = amoracle.spline_point(g["sp_ctrl"][i, :n, 0], g["sp_ctrl"][i, :n, 1], g["sp_t"][i])
        assert np.array_equal(out, g["sp_out"][i])
    w, lag, slope = g["w"], g["lag"], g["slope"]
    assert np.array_equal(np.array([amoracle.interpolate_color(a, b, x) for a, b, x in zip(cols[:2000], cols[1000:3000], w)], dtype=np.uint32), g["ic_plain"])
    assert np.array_equal(np.array([amoracle.interpolate_color(a, b, x, l, s) for a, b, x, l, s in zip(cols[:2000], cols[1000:3000], w, lag, slope)],
                                   dtype=np.uint32), g["ic_eased"])
    assert np.array_equal(np.array([amoracle.interpolate_point(a, b, x) for a, b, x in zip(pts[:1999], pts[1:2000], w)], dtype=np.uint64), g["ip"])


@pytest.mark.parametrize("case", gu.RENDER_CASES)
def test_render_bit_exact(case):
    g = gu.load(case)
    scene = gu.oracle_scene(amoracle, g)
    for t, ref in zip(g["times"], g["frames"]):
        assert np.array_equal(scene.render(t), ref), "%s t=%g" % (case, t)
    images, keys, bbox, frames, chains = gu.render_tables(g)
    assert amoracle.cost(chains) == float(g["cost"])


def test_serial_matcher_replays_the_reference():
    g = gu.load("swap")
    words, e1, gain = amoracle.morph_steps(g["before"], int(g["steps"]), int(g["cycle_length"]), int(g["e1"]))
    assert np.array_equal(words, g["after"])
    assert e1 == int(g["e1_after"])
    assert amoracle.cost([dict(words=words)]) == float(g["cost_after"])
    assert float(g["cost_before"]) - float(g["cost_after"]) == gain


@pytest.mark.parametrize("case", ["blobs_rects", "blobs_cloud"])
def test_blob_partition_and_energy(case):
    g = gu.load(case)
    n = len(g["images"])
    per_frame = []
    for i in range(n):
        labels, stats = amoracle.blobify(g["present_%d" % i], g["stored_%d" % i])
        assert np.array_equal(labels, gu.canonical(g["labels_%d" % i]))
        real = g["bsize_%d" % i] > 0
        assert len(stats) == int(real.sum())
        # match blobs through their first (smallest) position
        order = np.argsort(g["bfirst_%d" % i][real])
        ref_stats = g["bstats_%d" % i][real][order]
        assert np.array_equal(stats[:, 0], g["bsize_%d" % i][real][order].astype(float))
        assert np.allclose(stats[:, 1:], ref_stats, rtol=0, atol=1e-9)
        per_frame.append((g["bsize_%d" % i], g["bstats_%d" % i], g["bgroup_%d" % i]))
    # energy of the blob map that the blob groups describe (thread.cpp:1087-1107 over thread.cpp:1151-1174)
    w_xy, w_rgba, w_size, bbox_d = g["weights"]
    e = 0.0
    W = len(per_frame[0][0])
    by_group = [{int(gr): b for b, gr in enumerate(f[2])} for f in per_frame]
    for grp in range(W):
        for j in range(n):
            pj = (j - 1) % n
            b0, b1 = by_group[pj][grp], by_group[j][grp]
            e += amoracle.blob_distance(per_frame[pj][0][b0], per_frame[pj][1][b0], per_frame[j][0][b1], per_frame[j][1][b1], w_xy, w_rgba, w_size, int(bbox_d))
    assert abs(e - float(g["energy_best"])) <= 1e-9 * max(1.0, abs(e))


def test_fluid_step():
    g = gu.load("fluid")
    gx, gy, n = (int(v) for v in g["dims"])
    rec = g["rec"]
    for (sl, rad), ref_rec, ref_nodes in zip(g["steps"], g["after"], g["nodes"]):
        rec, nodes = amoracle.fluid_step(gx, gy, rec, int(sl), float(rad))
        act = ref_rec[:, 7] != 0
        assert np.array_equal(rec[act][:, :22], ref_rec[act][:, :22])
        assert np.array_equal(nodes, ref_nodes)


def test_reference_converged_cost_fixture_is_consistent():
    """tests/golden/c2_copt.json (BASELINE.md section 3 protocol, run on the unmodified reference by make_copt.py): the legs
    fall monotonically, and the recorded c_opt is the intercept of the least-squares line through (1/ppa, cost)."""
    import json
    import os
    import numpy as np
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_copt.json")))
    assert g["size"] == 1024 and g["atoms"] == 1024 * 1024
    ppa = np.array([leg["proposals_per_atom"] for leg in g["legs"]], dtype=np.float64)
    cost = np.array([leg["cost"] for leg in g["legs"]])
    assert list(ppa) == [1000.0, 2000.0, 4000.0]
    assert np.all(np.diff(cost) < 0) and cost[0] < g["cost_initial"] / 30
    slope, icpt = np.polyfit(1.0 / ppa, cost, 1)
    assert abs(icpt - g["c_opt"]) <= 1e-9 * g["c_opt"]
    assert 100.0 < g["k"] < 140.0                      # SURVEY.md section 8c measured k ~ 110 on this scene type
    assert g["c_opt"] < cost[-1] < 1.04 * g["c_opt"]
