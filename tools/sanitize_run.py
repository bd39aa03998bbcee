"""Workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every pipeline on the pin cell and on a jittered
mesh, tiny chunks, batched evaluation, count batches, record-pool exhaustion, compact download, sweep exports, volume correction.
usage: compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402
from raytracing_jl_b200 import _lib as L  # noqa: E402

d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pincell.npz"))
pin = rt.UnstructuredDiscreteModel(d["node_coordinates"], d["cell_ptrs"], d["cell_data"])
jit = rt.synth.jittered_triangle_mesh(48, 40, 1.5, 1.25, 0.25, 9, x0=-0.5, y0=2.0)
ref = None
runs = 0
cases = ((pin, 8, 0.02), (pin, 16, 0.08), (jit, 16, 0.01))
if os.environ.get("SAN_QUICK"):  # (initcheck keeps a shadow copy of every allocation: one mesh is enough for it)
    cases = cases[1:2]
for model, n_azim, delta in cases:
    ref = None
    for pipeline in (3, 0, 1):
        for chunk, cap, pool_slots, pool_extra in ((None, 0, 0, None), (5, 0, 0, None), (40, 3000, 0, None), (40, 0, 1024, None), (700, 0, 0, 0.0)):
            if pipeline != 3 and (pool_slots or pool_extra is not None):
                continue
            tg = rt.TrackGenerator(model, n_azim, delta, bcs=rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Vacuum, right=rt.Periodic, left=rt.Periodic),
                                   volume_correction=(chunk is None))
            rt.trace_(tg)
            tg.set_option("pipeline", pipeline)
            if os.environ.get("SAN_QUICK"):
                tg.set_option("debug_clear_pool", 1)
            if chunk:
                tg.set_option("chunk_segments", chunk)
                tg.set_option("target_walkers", 1e9)
            if cap:
                L.check(tg._ctx, L.lib().rt_set_segment_capacity(tg._ctx, cap))
            if pool_slots:
                tg.set_option("pool_slots", pool_slots)
            if pool_extra is not None:
                tg.set_option("pool_extra", pool_extra)
            for rep in range(2):  # (the second call takes the cached plan and the optimistic evaluation)
                rt.segmentize_(tg, check=False)
            off = tg.segment_offsets.copy()
            if ref is None:
                ref = off
            assert np.array_equal(off, ref), (pipeline, chunk, cap)
            full = {k: v.copy() for k, v in tg.fetch_segments().items()}
            comp = tg.fetch_segments(compact=True)
            assert all(np.array_equal(comp[k], full[k]) for k in full)
            if chunk is None:
                tg.track_view()
                tg.quadrature_device()
                tg.optical_lengths(np.ones((model.num_cells, 2)))
                tg.element_volumes()
            runs += 1
            tg.close()
tg = rt.TrackGenerator(rt.Mesh(pin, device_ingest=True), 8, 0.05)
rt.segmentize_(rt.trace_(tg), flags=rt.RT_SEG_LITERAL)
rt.segmentize_(tg, flags=rt.RT_SEG_NO_CHUNKS)
print("sanitize_run ok:", runs, "configurations", flush=True)
