import os, sys
sys.path.insert(0, os.getcwd())
import raytracing_jl_b200 as rt
model, n_azim, delta = rt.synth.workload("cfg3")
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(rt.Mesh(model), n_azim, delta, bcs=bcs)
rt.trace_(tg)
def run(label, **opts):
    for k, v in opts.items(): tg.set_option(k, v)
    best = None
    for _ in range(6):
        tg.timer_start()
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
        ms = tg.timer_stop()
        p = tg.phase_ms(); p["total"] = ms
        best = p if best is None or p["total"] < best["total"] else best
    print("%-40s total %.3f count %.3f fill %.3f fb %d" % (label, best["total"], best["count"], best["fill"], tg.info("verify_fallbacks")), flush=True)
for c in (16, 0, 3, 8, 32):
    run("classes=%d grid=32" % c, order_classes=c, order_grid=32)
run("classes=16 grid=8", order_classes=16, order_grid=8)
run("classes=16 grid=1", order_classes=16, order_grid=1)
