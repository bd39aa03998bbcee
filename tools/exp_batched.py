"""GPU experiment: wall (device time between two events, host gaps included) against the kernel phases of the batched path.
usage: python tools/exp_batched.py [cfg4] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model, n_azim, delta = rt.synth.workload(name)
tg = rt.TrackGenerator(rt.Mesh(model), n_azim, delta, bcs=rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective))
rt.trace_(tg)
for rep in range(reps + 1):
    t0 = time.perf_counter()
    tg.timer_start()
    rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
    ms = tg.timer_stop()
    wall = (time.perf_counter() - t0) * 1e3
    p = tg.phase_ms()
    k = p["count"] + p["scan"] + p["fill"] + p["volumes"]
    print(f"{name} rep {rep}: device {ms:8.2f} ms host wall {wall:8.2f} ms  phases count {p['count']:.2f} scan {p['scan']:.2f} fill {p['fill']:.2f} = {k:.2f} "
          f"({100 * (ms - k) / ms:.1f}% not in kernel phases)  segments {tg.n_segments:.4e} -> {tg.n_segments / ms * 1e3:.3e} seg/s  walk batches {tg.info('count_batches'):.0f} "
          f"cap {tg.info('segment_capacity'):.3e}", flush=True)
