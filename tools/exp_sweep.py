"""GPU experiment: chunk-size / walker-count sweep of the default pipeline on cfg3 (or argv[1])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
model, n_azim, delta = rt.synth.workload(name)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(model, n_azim, delta, bcs=bcs)
rt.trace_(tg)
def run(label, **o):
    tg.set_option("chunk_segments", o.get("cs", 64)); tg.set_option("target_walkers", o.get("tw", 148 * 2048 * 4)); tg.set_option("order_grid", o.get("og", 16))
    best = None
    for _ in range(4):
        tg.timer_start(); rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False); ms = tg.timer_stop(); p = tg.phase_ms()
        if best is None or ms < best[0]: best = (ms, p)
    print(f"{label:34s} total {best[0]:7.3f} count {best[1]['count']:6.3f} fill {best[1]['fill']:6.3f} units {tg.info('n_units'):.0f}", flush=True)
run("default")
for cs in (96, 128, 160, 192, 256):
    run(f"cs={cs} tw=1e9", cs=cs, tw=1e9)
for tw in (148 * 2048 * 2, 148 * 2048 * 3, 148 * 2048 * 6, 148 * 2048 * 8):
    run(f"tw={tw}", tw=tw)
for og in (8, 32):
    run(f"og={og} cs=128", og=og, cs=128, tw=1e9)
