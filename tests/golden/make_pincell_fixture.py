"""Converts the reference's config-1 fixture (demo/pincell.json, Gridap JSON; the mesh its own tests
load at test/runtests.jl:5-6) into a compact .npz that travels to the GPU box, where /root/reference
does not exist.  Also cross-checks demo/pincell.msh against it.  Run in the build container:

    python tests/golden/make_pincell_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import raytracing_jl_b200 as rt  # noqa: E402

REF = "/root/reference/demo"
m = rt.DiscreteModelFromFile(os.path.join(REF, "pincell.json"))
g = rt.GmshDiscreteModel(os.path.join(REF, "pincell.msh"))
assert np.array_equal(m.node_coordinates, g.node_coordinates), "msh nodes differ from json nodes"
assert np.array_equal(m.cell_data, g.cell_data), "msh triangles differ from json cells"
out = os.path.join(ROOT, "tests", "golden", "pincell.npz")
np.savez_compressed(out, node_coordinates=m.node_coordinates, cell_ptrs=m.cell_ptrs, cell_data=m.cell_data)
print("wrote", out, m.num_nodes, "nodes", m.num_cells, "cells", os.path.getsize(out), "bytes")
