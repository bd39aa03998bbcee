"""GPU experiment: chunk length / walker count of the single-walk pipeline on a named workload.
usage: [RT_B200_LIB=build_variants/x.so] python tools/exp_chunks.py [cfg3] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
model, n_azim, delta = rt.synth.workload(name)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(rt.Mesh(model), n_azim, delta, bcs=bcs)
rt.trace_(tg)
print("lib", os.environ.get("RT_B200_LIB", "default"), flush=True)
for cs, tw in ((None, None), (64, None), (96, None), (160, None), (192, None), (256, None), (None, 2e5), (None, 8e5), (None, None)):
    if cs is not None:
        tg.set_option("chunk_segments", cs)
    if tw is not None:
        tg.set_option("target_walkers", tw)
    for _ in range(4):
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
    tg.timer_start()
    for _ in range(reps):
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
    ms = tg.timer_stop() / reps
    p = tg.phase_ms()
    print(f"chunk_segments {cs} target_walkers {tw}: {ms:.4f} ms/step count {p['count']:.3f} fill {p['fill']:.3f} units {tg.info('n_units'):.0f} fb {tg.info('verify_fallbacks'):.0f}", flush=True)
