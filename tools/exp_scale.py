"""GPU experiment: per-segment cost of the walk and the evaluation against the mesh size (are the tables L2-resident?).
usage: python tools/exp_scale.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
for scale in (0.125, 0.25, 0.5, 1.0, 2.0, 4.0):
    model, n_azim, delta = rt.synth.workload("cfg3", scale=scale)
    delta = delta * (scale ** 0.5)  # same number of segments (~5e7) on every mesh
    tg = rt.TrackGenerator(rt.Mesh(model), n_azim, delta, bcs=bcs)
    rt.trace_(tg)
    best = None
    for _ in range(4):
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
        p = tg.phase_ms()
        if best is None or p["count"] + p["fill"] < best["count"] + best["fill"]:
            best = p
    n = tg.n_segments
    print(f"scale {scale:5.3f} cells {model.num_cells:8d} tables {model.num_cells * 96 / 1e6:6.1f} MB each  segments {n:.3e} tracks {tg.n_total_tracks}"
          f"  count {best['count']:.3f} ms = {best['count'] * 1e6 / n:.2f} ps/seg... fill {best['fill']:.3f} ms = {best['fill'] * 1e6 / n:.2f} ns/1000seg", flush=True)
    tg.close()
