"""GPU experiment: kernel-side segments/s of the small named workloads (pincell, cfg2 at both spacings, cfg3)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt

def run(name, model, n_azim, delta, reps=20):
    bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
    tg = rt.TrackGenerator(model, n_azim, delta, bcs=bcs)
    rt.trace_(tg)
    for _ in range(3):
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
    tg.timer_start()
    for _ in range(reps):
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
    ms = tg.timer_stop() / reps
    p = tg.phase_ms()
    print(f"{name}: cells {model.num_cells} tracks {tg.n_total_tracks} segments {tg.n_segments} ms/step {ms:.3f} "
          f"-> {tg.n_segments / ms * 1e3:.3e} seg/s (count {p['count']:.3f} fill {p['fill']:.3f}) bad {int((tg.segment_offsets is not None) and (tg.segment_status != 0).sum())}", flush=True)
    tg.close()

d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pincell.npz"))
run("cfg1 pincell nphi=8 delta=2e-2", rt.UnstructuredDiscreteModel(d["node_coordinates"], d["cell_ptrs"], d["cell_data"]), 8, 2e-2)
m2, na, dl = rt.synth.workload("cfg2")
run("cfg2 BWR nphi=16 delta=8e-2", m2, na, dl)
run("cfg2 BWR nphi=16 delta=2e-3", m2, na, 2e-3)
m3, na, dl = rt.synth.workload("cfg3")
run("cfg3", m3, na, dl, reps=10)
