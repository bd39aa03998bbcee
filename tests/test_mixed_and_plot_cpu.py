"""SURVEY 8(f)-4 on the CPU: the oracle's restatement of point_in_quadrangle (src/mesh.jl:184-201) and of the 4-edge
intersections (src/intersection.jl:42-44, 81-95) on quadrilateral cells, mixed-mesh host tables, and the plot-friendly views of
src/plot_recipes.jl."""
import os

import numpy as np
import pytest

import raytracing_jl_b200 as rt
from oracle.oracle import OracleError, OracleMesh, OracleTrackGenerator


def unit_quads():
    # two unit squares side by side, nodes stored as a counter-clockwise cycle; and the same squares cut into triangles
    xy = np.array([[0, 0], [1, 0], [2, 0], [0, 1], [1, 1], [2, 1]], float)
    quads = rt.UnstructuredDiscreteModel.from_cells(xy, [(0, 1, 4, 3), (1, 2, 5, 4)])
    tris = rt.UnstructuredDiscreteModel.from_cells(xy, [(0, 1, 4), (0, 4, 3), (1, 2, 5), (1, 5, 4)])
    return quads, tris


def test_point_in_quadrangle_and_four_edge_chords():
    quads, _ = unit_quads()
    om = OracleMesh.from_mesh(rt.Mesh(quads))
    assert om.point_in_element(1, 0.9, 0.1) and om.point_in_element(1, 0.1, 0.9)  # both halves of the square, whatever triangle holds them
    assert om.point_in_element(1, 1.0, 0.5) and om.point_in_element(2, 1.0, 0.5)  # the shared edge belongs to both (tolerant test)
    assert not om.point_in_element(1, 1.5, 0.5) and om.point_in_element(2, 1.5, 0.5)
    assert om.find_element(0.25, 0.75) == 1 and om.find_element(1.75, 0.25) == 2
    # a horizontal line y = 0.5: general form of (0, .5) -> (2, .5)
    abc = np.zeros(3)
    from oracle.oracle import lib
    lib().orc_general_form(0.0, 0.5, 2.0, 0.5, abc)
    rc, pq, ed, n = om.intersections(1, abc, 0.0)
    assert rc == 0 and n == 2 and np.allclose(pq, [0, .5, 1, .5]) and sorted(ed.tolist()) == [1, 3]  # right and left edge of the cycle
    # the diagonal y = x crosses square 1 through two of its corners: four edge hits, the farthest pair wins (src/intersection.jl:81-95)
    lib().orc_general_form(0.0, 0.0, 1.0, 1.0, abc)
    rc, pq, ed, n = om.intersections(1, abc, np.pi / 4)
    assert rc == 0 and n == 4 and np.allclose(pq, [0, 0, 1, 1])


def test_mixed_mesh_walk_covers_the_area_and_matches_its_triangulation_in_length():
    model = rt.synth.mixed_quad_triangle_mesh(9, 7, 1.8, 1.4, jitter=0.2, seed=5, quad_fraction=0.5)
    assert model.has_quads and set(model.cell_sizes.tolist()) == {3, 4}
    mesh = rt.Mesh(model)
    ptrs, data = mesh.node_cells
    assert ptrs[-1] - 1 == model.cell_data.size  # every (cell, node) incidence once
    otg = OracleTrackGenerator(OracleMesh.from_mesh(mesh), 16, 0.03, bcs=(1, 1, 1, 1)).trace().segmentize()
    assert otg.bad_status == 0 and np.all(otg.seg_status == 0)
    area = rt.synth.mesh_area(model)
    assert np.isclose(area, 1.8 * 1.4, rtol=1e-12)
    assert np.isclose(otg.volumes().sum(), area, rtol=2e-3)  # the tracked volumes approximate the areas, cell by cell
    # reference invariants (test/runtests.jl:30-43): first p, last q, sum of lengths
    off, t = otg.seg_offsets, otg.tracks
    for u in range(otg.n_total_tracks):
        a, b = off[u], off[u + 1]
        assert b > a
        assert np.allclose([otg.seg["px"][a], otg.seg["py"][a]], t["p"][u], rtol=0, atol=1e-7)
        assert np.allclose([otg.seg["qx"][b - 1], otg.seg["qy"][b - 1]], t["q"][u], rtol=0, atol=1e-7)
        assert np.isclose(otg.seg["len"][a:b].sum(), t["len"][u], rtol=1.5e-8)
    quad_ids = np.nonzero(model.cell_sizes == 4)[0] + 1
    assert np.isin(otg.seg["element"], quad_ids).mean() > 0.3  # the walk does cross the quadrilaterals


def test_k_limit_is_an_error_not_a_clamp():
    d = rt.synth.jittered_triangle_mesh(6, 6, seed=2)
    otg = OracleTrackGenerator(OracleMesh.from_mesh(rt.Mesh(d)), 4, 0.2).trace()
    otg.segmentize(k=32)
    with pytest.raises(OracleError):
        otg.segmentize(k=33)


def test_plot_views_match_the_recipes():
    quads, tris = unit_quads()
    x, y = rt.plotdata.mesh_lines(rt.Mesh(quads))  # src/plot_recipes.jl:76-107: (nn + 1) x n_cells, closed cycles
    assert x.shape == (5, 2) and np.array_equal(x[:, 0], [0, 1, 1, 0, 0]) and np.array_equal(y[:, 1], [0, 0, 1, 1, 0])
    x, y = rt.plotdata.mesh_lines(rt.Mesh(tris))
    assert x.shape == (4, 4) and np.array_equal(x[0], x[-1])
    mixed = rt.synth.mixed_quad_triangle_mesh(3, 2, seed=1)
    with pytest.raises(ValueError):
        rt.plotdata.mesh_lines(rt.Mesh(mixed))  # the reference's `error("error")` for cells of different sizes
    fx, fy = rt.plotdata.flat_mesh_edges(rt.Mesh(mixed))
    assert np.isnan(fx).sum() == mixed.num_cells - 1 and fx.size == mixed.cell_data.size + 2 * mixed.num_cells - 1
    cols = dict(px=np.array([0., 1.]), py=np.array([0., 0.]), qx=np.array([1., 2.]), qy=np.array([0., 1.]), element=np.array([7, 9], np.int32))
    sx, sy, sz = rt.plotdata.segment_lines(cols)  # src/plot_recipes.jl:26-48
    assert np.array_equal(sx, [[0, 1], [1, 2]]) and np.array_equal(sz, [[7, 9], [7, 9]])
    fx, fy, fz = rt.plotdata.flat_polyline(sx, sy, sz)
    assert np.array_equal(fx[[0, 1, 3, 4]], [0, 1, 1, 2]) and np.isnan(fx[2]) and fx.size == 5 and np.array_equal(fz[[0, 1, 3, 4]], [7, 7, 9, 9])


@pytest.mark.skipif(not os.path.isdir("/root/reference/demo"), reason="the reference tree only exists in the build container")
def test_mesh_readers_on_the_reference_files():
    """Both mesh files the reference ships for its quick start (demo/pincell.json: Gridap JSON, loaded by test/runtests.jl:5-6;
    demo/pincell.msh: MSH 4.1) read into the same model, which is the committed fixture tests/golden/pincell.npz."""
    import raytracing_jl_b200 as rt

    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pincell.npz"))
    for m in (rt.DiscreteModelFromFile("/root/reference/demo/pincell.json"), rt.GmshDiscreteModel("/root/reference/demo/pincell.msh")):
        assert np.array_equal(m.node_coordinates, d["node_coordinates"])
        assert np.array_equal(m.cell_ptrs, d["cell_ptrs"]) and np.array_equal(m.cell_data, d["cell_data"])
