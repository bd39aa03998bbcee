"""GPU experiment: wall-clock breakdown of the e2e step (host buffers in -> host buffers out)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
model, n_azim, delta = rt.synth.workload(name)
mesh = rt.Mesh(model)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(mesh, n_azim, delta, bcs=bcs)
for it in range(4):
    t = [time.perf_counter()]
    tg.upload_mesh(); t.append(time.perf_counter())
    rt.trace_(tg); t.append(time.perf_counter())
    rt.segmentize_(tg, check=False, rtol=1e-6); t.append(time.perf_counter())
    tg.segment_offsets; t.append(time.perf_counter())
    seg = tg.fetch_segments(pinned=True); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    gb = sum(v.nbytes for v in seg.values()) / 1e9
    print(f"it{it}: upload {d[0]:.2f} trace {d[1]:.2f} segmentize+volumes {d[2]:.2f} offsets {d[3]:.2f} download {d[4]:.2f} ms"
          f" ({gb:.2f} GB -> {gb / d[4] * 1e3:.1f} GB/s) total {sum(d):.1f} ms; phases {tg.phase_ms()}")
