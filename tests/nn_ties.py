"""CPU diagnostic: how many find_element queries of a workload have two nearest nodes at EXACTLY the same distance?  Only those
depend on how an exact nearest-neighbour search breaks ties (lowest node id in the oracle and on the device; NearestNeighbors.jl,
which the reference uses at src/mesh.jl:108,124, does not document its choice).  Lives under tests/ because it runs the oracle (test infrastructure), with its
tie counter on.  usage: python tests/nn_ties.py [pincell cfg2 cfg3 cfg4:0.02 cfg5:0.002 structured]   (name[:fraction of the tracks])"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracing_jl_b200 as rt  # noqa: E402
from oracle import oracle as O  # noqa: E402

O.set_diag_ties(True)
threads = os.cpu_count() or 1
for arg in sys.argv[1:] or ["pincell", "cfg2", "cfg3"]:
    name, _, frac = arg.partition(":")
    frac = float(frac) if frac else 1.0
    if name == "pincell":
        d = np.load(os.path.join(ROOT, "tests", "golden", "pincell.npz"))
        model, n_azim, delta = rt.UnstructuredDiscreteModel(d["node_coordinates"], d["cell_ptrs"], d["cell_data"]), 8, 2e-2
    elif name == "structured":
        model, n_azim, delta = rt.synth.jittered_triangle_mesh(16, 16, jitter=0.0), 8, 0.0625  # test_structured_mesh_exact_vertex_crossings
    else:
        model, n_azim, delta = rt.synth.workload(name)
    t0 = time.time()
    otg = O.OracleTrackGenerator(O.OracleMesh.from_mesh(rt.Mesh(model)), n_azim, delta, bcs=(1, 1, 1, 1) if name != "structured" else (0, 0, 0, 0)).trace()
    n = otg.n_total_tracks
    tot = {"tracks": 0, "segments": 0, "steps": 0, "nn_ties": 0, "knn_fallbacks": 0}
    if frac >= 1.0:
        ranges = [(1, n + 1)]
    else:  # uid blocks spread over the whole track set (every angle)
        nb = 64
        w = max(1, int(n * frac / nb))
        ranges = [(int(u), min(int(u) + w, n + 1)) for u in np.linspace(1, n - w, nb)]
    for u0, u1 in ranges:
        otg.segmentize(rtol=1e-6, uid_begin=u0, uid_end=u1, nthreads=threads, fetch=False, check=False)
        st = otg.stats()
        tot["tracks"] += u1 - u0
        tot["segments"] += otg.n_segments
        for k in ("steps", "nn_ties", "knn_fallbacks"):
            tot[k] += st[k]
        otg.free_segments()
    print(f"{arg:12s} cells {model.num_cells:9d} tracks {tot['tracks']:9d} of {n:9d} segments {tot['segments']:.4e} find_element queries "
          f"{tot['steps']:9d} (knn branch {tot['knn_fallbacks']}) exact nearest-node ties {tot['nn_ties']}   [{time.time() - t0:.0f} s]", flush=True)
