"""GPU experiment: phase times of rt_segmentize under different flags / chunk options (cfg3 by default)."""
import json
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
model, n_azim, delta = rt.synth.workload(name)
if len(sys.argv) > 2:
    delta = float(sys.argv[2])
mesh = rt.Mesh(model)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(mesh, n_azim, delta, bcs=bcs)
rt.trace_(tg)
print("tracks", tg.n_total_tracks, "cells", model.num_cells)


def run(label, flags=0, reps=3, **opts):
    tg.set_option("chunk_segments", opts.get("chunk_segments", 128))
    tg.set_option("target_walkers", opts.get("target_walkers", 148 * 2048 * 4))
    tg.set_option("order_grid", opts.get("order_grid", 32))
    tg.set_option("pipeline", opts.get("pipeline", 0))
    tg.set_option("eval_waves", opts.get("eval_waves", int(os.environ.get("RT_EVAL_WAVES", "1"))))
    best = None
    for _ in range(reps):
        tg.timer_start()
        rt.segmentize_(tg, flags=flags, check=False, fetch_volumes=False)
        ms = tg.timer_stop()
        p = tg.phase_ms()
        if best is None or ms < best[0]:
            best = (ms, p)
    st = tg.stats()
    print(f"{label:40s} total {best[0]:8.3f} ms  count {best[1]['count']:7.3f} fill {best[1]['fill']:7.3f} scan {best[1]['scan']:6.3f}"
          f"  nseg {tg.n_segments}  seg/s {tg.n_segments / best[0] * 1e3:.3e}  fast {st['fast_transitions']:.0f} lit {st['literal_iterations']:.0f}")


run("default (hybrid)")
print("fallbacks", tg.info("verify_fallbacks"))
if os.environ.get("RT_EXP_MIN"):
    sys.exit(0)
run("sequential", rt.RT_SEG_SEQUENTIAL)
if os.environ.get("RT_EXP_QUICK"):
    for og in (0, 4, 8, 32, 64, 128):
        run(f"order_grid={og}", order_grid=og)
    run("no volumes", rt.RT_SEG_NO_VOLUMES)
    sys.exit(0)
run("no volumes", rt.RT_SEG_NO_VOLUMES)
run("no chunks", rt.RT_SEG_NO_CHUNKS)
run("no chunks, no volumes", rt.RT_SEG_NO_CHUNKS | rt.RT_SEG_NO_VOLUMES)
run("count only", rt.RT_SEG_COUNT_ONLY | rt.RT_SEG_NO_VOLUMES)
for cs in (16, 32, 128, 256, 512):
    run(f"chunk_segments={cs}", chunk_segments=cs, target_walkers=1e9)
    run(f"chunk_segments={cs} novol", rt.RT_SEG_NO_VOLUMES, chunk_segments=cs, target_walkers=1e9)
