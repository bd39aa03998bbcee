"""GPU experiment: the 8 uid shards of the weak-scaling bench (delta/8) one after the other on ONE GPU: where does the time of a
shard go (phases, walk statistics)?  usage: python tools/exp_shards.py [cfg3] [world]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
div = int(sys.argv[3]) if len(sys.argv) > 3 else world  # track spacing = delta / div
model, n_azim, delta = rt.synth.workload(name)
mesh = rt.Mesh(model)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
for r in range(world):
    tg = rt.TrackGenerator(mesh, n_azim, delta / div, bcs=bcs, shard=(r, world))
    rt.trace_(tg)
    best = None
    for _ in range(4):
        tg.timer_start()
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
        ms = tg.timer_stop()
        if best is None or ms < best[0]:
            best = (ms, tg.phase_ms())
    st = tg.stats()
    az = [t.azim_idx for t in (tg.tracks_by_uid[tg.uid_begin], tg.tracks_by_uid[tg.uid_end - 1])] if hasattr(tg, "tracks_by_uid") else []
    print(f"shard {r}/{world}: uids [{tg.uid_begin}, {tg.uid_end}) segments {tg.n_segments} total {best[0]:.3f} ms count {best[1]['count']:.3f} fill {best[1]['fill']:.3f} "
          f"azim {az} fast {st['fast_transitions']:.0f} literal iters {st['literal_iterations']:.0f} nn {st['nn_queries']:.0f} fb {tg.info('verify_fallbacks'):.0f}", flush=True)
    tg.close()
