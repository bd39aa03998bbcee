import os, sys
sys.path.insert(0, os.getcwd())
import raytracing_jl_b200 as rt
model, n_azim, delta = rt.synth.workload("cfg3")
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(rt.Mesh(model), n_azim, delta, bcs=bcs)
rt.trace_(tg)
best = None
for _ in range(6):
    tg.timer_start()
    rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
    ms = tg.timer_stop()
    p = tg.phase_ms()
    p["total"] = ms
    best = p if best is None or p["total"] < best["total"] else best
print(os.path.basename(os.environ.get("RT_B200_LIB", "default")), "total %.3f count %.3f fill %.3f" % (best["total"], best["count"], best["fill"]), "fb", tg.info("verify_fallbacks"), flush=True)
