"""GPU experiment: steady-state phase times of segmentize! under sets of rt_set_option knobs (chunk_segments, band_min, band_div,
order_grid, order_classes, target_walkers, ...).
usage: [RT_B200_LIB=...] python tools/exp_opts.py [cfg3] [reps] [k=v,k=v ...]   (every argument is one run; options persist between runs)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
model, n_azim, delta = rt.synth.workload(name)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(rt.Mesh(model), n_azim, delta, bcs=bcs)
rt.trace_(tg)
lib = os.path.basename(os.environ.get("RT_B200_LIB", "default"))
for arg in sys.argv[3:] or [""]:
    for kv in filter(None, arg.split(",")):
        k, v = kv.split("=")
        tg.set_option(k, float(v))
    for _ in range(3):
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
    best = None
    for _ in range(3):
        tg.timer_start()
        for _ in range(reps):
            rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
        ms = tg.timer_stop() / reps
        p = tg.phase_ms()
        if best is None or ms < best[0]:
            best = (ms, p)
    ms, p = best
    print(f"{lib:16s} {arg:44s}: {ms:.4f} ms/step count {p['count']:.3f} fill {p['fill']:.3f} units {tg.info('n_units'):.0f} "
          f"nseg {tg.n_segments} fb {tg.info('verify_fallbacks'):.0f} bad {tg.bad_status}", flush=True)
