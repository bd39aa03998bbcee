"""Pins the CPU oracle (oracle/rt_oracle.c) against every golden the reference's own tests hold for the
trace! -> segmentize! path (test/runtests.jl), plus a second pure-Python restatement on tiny cases."""
import math

import numpy as np
import pytest

from oracle.oracle import BC, RTOL, OracleError, OracleMesh, OracleTrackGenerator, lib
from tests.golden import runtests_goldens as G
from tests import pyref


def isapprox_pt(p, q):
    return np.linalg.norm(p - q) <= RTOL * max(np.linalg.norm(p), np.linalg.norm(q))


def bcs_tuple(**kw):
    d = dict(top="Vacuum", bottom="Vacuum", right="Vacuum", left="Vacuum")
    d.update(kw)
    return tuple(BC[d[s]] for s in ("top", "bottom", "right", "left"))


@pytest.fixture(scope="module")
def main_tg(pincell_oracle_mesh):
    tg = OracleTrackGenerator(pincell_oracle_mesh, G.MAIN["n_azim"], G.MAIN["delta"])
    tg.trace()
    tg.segmentize()
    return tg


def test_bbox(pincell_oracle_mesh):
    a, b = pincell_oracle_mesh.bbox()
    assert a.tolist() == [0.0, 0.0] and b.tolist() == [1.6, 1.6]


def test_tracing_counts(main_tg):  # test/runtests.jl:14-19
    assert main_tg.n_total_tracks == G.MAIN["n_total_tracks"]
    assert main_tg.n_tracks_x.tolist() == G.MAIN["n_tracks_x"]
    assert main_tg.n_tracks_y.tolist() == G.MAIN["n_tracks_y"]
    assert main_tg.n_tracks.tolist() == G.MAIN["n_tracks"]


def test_azimuthal_quadrature(main_tg):  # test/runtests.jl:21-28
    assert main_tg.n2 == G.MAIN["nazim2"]
    for d in main_tg.deltas:
        assert math.isclose(d, G.MAIN["delta_eff"], rel_tol=RTOL)
    a, b = main_tg.phis, np.array(G.MAIN["phis"])
    assert np.linalg.norm(a - b) <= RTOL * max(np.linalg.norm(a), np.linalg.norm(b))
    assert math.isclose(2 * main_tg.weights.sum(), 0.5, rel_tol=1e-12) or main_tg.weights.sum() > 0


def test_entry_exit_points(main_tg):  # test/runtests.jl:30-35
    t, s, off = main_tg.tracks, main_tg.seg, main_tg.seg_offsets
    for u in range(main_tg.n_total_tracks):
        a, b = off[u], off[u + 1] - 1
        assert isapprox_pt(t["p"][u], np.array([s["px"][a], s["py"][a]]))
        assert isapprox_pt(t["q"][u], np.array([s["qx"][b], s["qy"][b]]))


def test_track_length(main_tg):  # test/runtests.jl:37-43
    t, s, off = main_tg.tracks, main_tg.seg, main_tg.seg_offsets
    for u in range(main_tg.n_total_tracks):
        l2 = s["len"][off[u]:off[u + 1]].sum()
        assert math.isclose(t["len"][u], l2, rel_tol=RTOL)


@pytest.mark.parametrize("n_azim,table", [(4, G.LINKS_4), (8, G.LINKS_8)])
def test_reflection_links(pincell_oracle_mesh, n_azim, table):  # test/runtests.jl:46-335
    tg = OracleTrackGenerator(pincell_oracle_mesh, n_azim, 0.8, bcs=bcs_tuple(**G.REFLECTION_BCS))
    tg.trace()
    assert tg.n_total_tracks == len(table)
    t = tg.tracks
    for uid, (bf, bb, nf, nb, df, db) in table.items():
        u = uid - 1
        got = (t["bc_fwd"][u], t["bc_bwd"][u], t["next_fwd"][u], t["next_bwd"][u], t["dir_fwd"][u], t["dir_bwd"][u])
        assert got == (G.BC_CODE[bf], G.BC_CODE[bb], nf, nb, G.DIR_CODE[df], G.DIR_CODE[db]), uid


def test_domain_errors(pincell_oracle_mesh):  # src/azimuthal_quad.jl:22-25
    for args in [(0, 0.1), (-4, 0.1), (6, 0.1), (8, 0.0), (8, -1.0)]:
        with pytest.raises(OracleError):
            OracleTrackGenerator(pincell_oracle_mesh, *args)


def test_segmentize_before_trace(pincell_oracle_mesh):  # src/trackgenerator.jl:360-361
    tg = OracleTrackGenerator(pincell_oracle_mesh, 4, 0.8)
    with pytest.raises(OracleError):
        tg.segmentize()


def test_volumes(main_tg, pincell_model):  # src/trackgenerator.jl:371-386
    v = main_tg.volumes()
    assert math.isclose(v.sum(), 1.6 * 1.6, rel_tol=1e-9)
    xy, t = pincell_model.node_coordinates, pincell_model.triangles0()
    p0, p1, p2 = xy[t[:, 0]], xy[t[:, 1]], xy[t[:, 2]]
    area = 0.5 * np.abs((p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1]) - (p2[:, 0] - p0[:, 0]) * (p1[:, 1] - p0[:, 1]))
    assert np.abs(v / area - 1).max() < 0.6 and abs((v / area).mean() - 1) < 2e-2  # 4 angles, ~2 tracks per cell


def test_periodic_links_are_a_permutation(pincell_oracle_mesh):
    """No Periodic golden exists in the reference; check the structural property instead: with all
    sides Periodic, next_fwd and next_bwd are mutually inverse permutations keeping the angle."""
    tg = OracleTrackGenerator(pincell_oracle_mesh, 16, 0.05, bcs=(2, 2, 2, 2))
    tg.trace()
    t = tg.tracks
    n = tg.n_total_tracks
    nf, nb = t["next_fwd"] - 1, t["next_bwd"] - 1
    assert sorted(nf.tolist()) == list(range(n)) and sorted(nb.tolist()) == list(range(n))
    assert np.array_equal(nb[nf], np.arange(n))
    assert np.array_equal(t["azim_idx"][nf], t["azim_idx"])
    # a periodic forward link continues from the opposite side: q and next p differ by one period
    d = t["p"][nf] - t["q"]
    per = np.isclose(np.abs(d), 1.6, atol=1e-9) | np.isclose(d, 0.0, atol=1e-9)
    assert per.all()


def test_reflective_links_touch(pincell_oracle_mesh):
    """Reflective/Vacuum links: the next track starts (or ends, if traversed backward) where this ends."""
    tg = OracleTrackGenerator(pincell_oracle_mesh, 8, 0.05, bcs=(1, 1, 1, 1))
    tg.trace()
    t = tg.tracks
    nf = t["next_fwd"] - 1
    nxt = np.where(t["dir_fwd"][:, None] == 0, t["p"][nf], t["q"][nf])
    assert np.abs(nxt - t["q"]).max() < 1e-9
    nb = t["next_bwd"] - 1
    nxt = np.where(t["dir_bwd"][:, None] == 0, t["p"][nb], t["q"][nb])
    assert np.abs(nxt - t["p"]).max() < 1e-9


def test_kdtree_matches_bruteforce(pincell_oracle_mesh):
    rng = np.random.default_rng(7)
    om = pincell_oracle_mesh
    pts = rng.uniform(-0.1, 1.7, size=(4000, 2))
    for x, y in pts:
        assert om.nn(x, y) == om.nn_brute(x, y)
    xy = om.xy
    for x, y in pts[:500]:
        nn = om.nn(x, y)
        ids = om.knn(x, y, 5, nn)
        d2 = (xy[:, 0] - x) ** 2 + (xy[:, 1] - y) ** 2
        order = np.lexsort((np.arange(d2.size), d2)) + 1
        assert order[0] == nn and ids.tolist() == order[1:6].tolist()


def test_primitives_match_python():
    rng = np.random.default_rng(11)
    L = lib()
    for _ in range(2000):
        a = rng.uniform(-2, 2, size=6)
        abc = np.zeros(3)
        L.orc_general_form(a[0], a[1], a[2], a[3], abc)
        assert tuple(abc) == pyref.general_form((a[0], a[1]), (a[2], a[3]))
        abc2 = np.array(pyref.general_form((a[2], a[3]), (a[4], a[5])))
        xy = np.zeros(2)
        par = L.orc_intersection(abc, abc2, xy)
        par_py, x_py = pyref.intersection(tuple(abc), tuple(abc2))
        assert bool(par) == par_py and tuple(xy) == x_py
        t = rng.uniform(-0.2, 1.2)
        x = (a[0] + t * (a[2] - a[0]), a[1] + t * (a[3] - a[1]))
        assert bool(L.orc_point_in_segment(a[0], a[1], a[2], a[3], x[0], x[1])) == pyref.point_in_segment(
            (a[0], a[1]), (a[2], a[3]), x)
    # reversing an edge negates (A,B,C) exactly and leaves the intersection bit-identical (used by the kernels)
    for _ in range(500):
        a = rng.uniform(-2, 2, size=8)
        t_abc = np.array(pyref.general_form((a[4], a[5]), (a[6], a[7])))
        e1, e2 = np.zeros(3), np.zeros(3)
        L.orc_general_form(a[0], a[1], a[2], a[3], e1)
        L.orc_general_form(a[2], a[3], a[0], a[1], e2)
        assert np.array_equal(e1, -e2)
        x1, x2 = np.zeros(2), np.zeros(2)
        assert L.orc_intersection(t_abc, e1, x1) == L.orc_intersection(t_abc, e2, x2)
        assert np.array_equal(x1, x2)


def test_isapprox_semantics():
    L = lib()
    assert L.orc_isapprox_scalar(1.0, 1.0 + 1e-9, 0.0, RTOL)
    assert not L.orc_isapprox_scalar(1.0, 1.0 + 1e-7, 0.0, RTOL)
    assert L.orc_isapprox_scalar(0.0, 0.0, 0.0, RTOL) and not L.orc_isapprox_scalar(0.0, 1e-300, 0.0, RTOL)
    assert L.orc_isapprox_point(0.0, 0.0, 0.0, 0.0)
    # array isapprox scales with distance from the ORIGIN
    assert L.orc_isapprox_point(100.0, 0.0, 100.0 + 1e-7, 0.0) and not L.orc_isapprox_point(1.0, 0.0, 1.0 + 1e-7, 0.0)


@pytest.mark.parametrize("n_azim,delta", [(4, 0.8), (8, 0.4)])
def test_walk_matches_pure_python(pincell_mesh, pincell_oracle_mesh, n_azim, delta):
    tg = OracleTrackGenerator(pincell_oracle_mesh, n_azim, delta)
    tg.trace()
    tg.segmentize()
    ref = pyref.PyRef(pincell_mesh)
    t, s, off = tg.tracks, tg.seg, tg.seg_offsets
    for u in range(tg.n_total_tracks):
        segs = ref.walk(tuple(t["p"][u]), float(t["phi"][u]), tuple(t["abc"][u]))
        assert len(segs) == off[u + 1] - off[u]
        got = np.stack([s[k][off[u]:off[u + 1]] for k in ("px", "py", "qx", "qy", "len")], 1)
        assert np.array_equal(got, np.array([g[:5] for g in segs]))
        assert s["element"][off[u]:off[u + 1]].tolist() == [g[5] for g in segs]


def test_walk_matches_pure_python_on_other_meshes():
    """the second, pure-Python restatement (tests/pyref.py: brute-force nearest nodes, barycentric point test, its own walk loop)
    against the C oracle at segment level on a jittered mesh (knn branch live), the BWR lattice, an unjittered structured mesh and
    two meshes with quadrilaterals"""
    import raytracing_jl_b200 as rt

    cases = [(rt.synth.jittered_triangle_mesh(20, 20, seed=7), 8, 0.1), (rt.synth.workload("cfg2")[0], 4, 0.9),
             (rt.synth.jittered_triangle_mesh(12, 12, jitter=0.0), 8, 0.11),
             (rt.synth.mixed_quad_triangle_mesh(14, 14, seed=3), 8, 0.09),  # quadrilaterals: point_in_quadrangle, 4-edge intersections
             (rt.synth.mixed_quad_triangle_mesh(10, 12, quad_fraction=1.0, seed=5), 4, 0.3)]
    checked = 0
    for model, n_azim, delta in cases:
        mesh = rt.Mesh(model)
        tg = OracleTrackGenerator(OracleMesh.from_mesh(mesh), n_azim, delta).trace()
        tg.segmentize(check=False)
        ref = pyref.PyRef(mesh)
        t, s, off = tg.tracks, tg.seg, tg.seg_offsets
        for u in range(tg.n_total_tracks):
            if tg.seg_status[u] != 0:
                continue
            segs = ref.walk(tuple(t["p"][u]), float(t["phi"][u]), tuple(t["abc"][u]))
            assert len(segs) == off[u + 1] - off[u], (u, len(segs))
            got = np.stack([s[k][off[u]:off[u + 1]] for k in ("px", "py", "qx", "qy", "len")], 1)
            assert np.array_equal(got, np.array([g[:5] for g in segs]).reshape(-1, 5))
            assert s["element"][off[u]:off[u + 1]].tolist() == [g[5] for g in segs]
            checked += len(segs)
    assert checked > 2500


def test_oracle_on_jittered_mesh():
    """on jittered meshes the knn branch of find_element is live; no errors expected."""
    import raytracing_jl_b200 as rt

    mesh = rt.Mesh(rt.synth.jittered_triangle_mesh(60, 60, seed=1234))
    tg = OracleTrackGenerator(OracleMesh.from_mesh(mesh), 16, 0.01)
    tg.trace()
    tg.segmentize()
    assert tg.bad_status == 0
    assert math.isclose(tg.volumes().sum(), 1.0, rel_tol=1e-9)
    st = tg.stats()
    assert st["knn_fallbacks"] > 0 and st["k_retries"] == 0
    t, s, off = tg.tracks, tg.seg, tg.seg_offsets
    # consecutive segments of a track are contiguous: q_k ~ p_{k+1}
    for u in range(0, tg.n_total_tracks, 37):
        a, b = off[u], off[u + 1]
        assert np.abs(s["qx"][a:b - 1] - s["px"][a + 1:b]).max(initial=0) < 1e-7


def test_nearest_node_ties_never_decide(pincell_model):
    """The reference locates a point with NearestNeighbors.jl's exact nn / knn (src/mesh.jl:108,124), which does not document how
    it breaks distance ties; oracle and device take the lowest node id.  That freedom only exists for a query whose two nearest
    nodes are at EXACTLY the same distance -- and no find_element query of the walks has one: not on the pin cell, the BWR
    lattice, a jittered mesh, nor on the unjittered structured mesh whose tracks run through vertices (tests/nn_ties.py runs
    the same count over cfg3 at its named size and over samples of cfg4 / cfg5: 0 in 4.2e8 queries, profiles/r2_nn_ties.txt)."""
    import raytracing_jl_b200 as rt
    from oracle import oracle as O

    # the counter sees a tie when there is one: the midpoint of a mesh edge is equidistant from the edge's two nodes
    sq = rt.Mesh(rt.synth.jittered_triangle_mesh(2, 2, jitter=0.0))
    om = OracleMesh.from_mesh(sq)
    x0, y0 = om.xy[0]
    d2 = ((om.xy - om.xy[0]) ** 2).sum(1)
    j = int(np.argsort(d2)[1])  # a node nearest to node 0
    xm, ym = 0.5 * (om.xy[0] + om.xy[j])
    assert O.lib().orc_nn_is_tied(om._h, float(xm), float(ym)) == 1
    assert O.lib().orc_nn_is_tied(om._h, float(x0 + 0.25 * (om.xy[j][0] - x0) + 1e-3), float(y0 + 0.25 * (om.xy[j][1] - y0) + 2e-3)) == 0

    cases = [(pincell_model, 8, 2e-2, (1, 1, 1, 1)), (rt.synth.workload("cfg2")[0], 16, 8e-2, (1, 1, 1, 1)),
             (rt.synth.jittered_triangle_mesh(60, 60, seed=1234), 16, 0.01, (0, 1, 2, 2)),
             (rt.synth.jittered_triangle_mesh(16, 16, jitter=0.0), 8, 0.0625, (0, 0, 0, 0))]
    O.set_diag_ties(True)
    try:
        for model, n_azim, delta, bcs in cases:
            tg = OracleTrackGenerator(OracleMesh.from_mesh(rt.Mesh(model)), n_azim, delta, bcs=bcs).trace()
            tg.segmentize(check=False, fetch=False)
            st = tg.stats()
            assert st["steps"] > 2000 and st["nn_ties"] == 0, st
    finally:
        O.set_diag_ties(False)
