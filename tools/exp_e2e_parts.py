"""GPU experiment: where the end-to-end step of bench.py spends its time (host wall clock per call, cfg3)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

model, n_azim, delta = rt.synth.workload("cfg3")
tg = rt.TrackGenerator(rt.Mesh(model), n_azim, delta, bcs=rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective))
tg.pin_mesh()
for rep in range(4):
    t = [time.perf_counter()]
    tg.upload_mesh(); t.append(time.perf_counter())
    rt.trace_(tg); t.append(time.perf_counter())
    rt.segmentize_(tg, rtol=1e-6, check=False); t.append(time.perf_counter())
    tg.segment_offsets; t.append(time.perf_counter())
    seg = tg.fetch_segments(pinned=True, compact=True); t.append(time.perf_counter())
    names = ["upload_mesh", "trace", "segmentize(+volumes)", "offsets", "fetch compact"]
    print("rep", rep, " ".join(f"{n} {1e3 * (b - a):.2f}" for n, a, b in zip(names, t[:-1], t[1:])), f"total {1e3 * (t[-1] - t[0]):.2f} ms",
          "GB/s of fetch %.1f" % (28 * tg.n_segments / (t[-1] - t[-2]) / 1e9), tg.phase_ms(), flush=True)
