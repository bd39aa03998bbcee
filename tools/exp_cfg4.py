"""GPU experiment: cfg4 at full size, pipelines 0 and 3: totals, batches, timing, per-track count differences."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402
from raytracing_jl_b200 import _lib as L  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
cap = int(float(sys.argv[2])) if len(sys.argv) > 2 else 0
model, n_azim, delta = rt.synth.workload(name)
mesh = rt.Mesh(model)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(mesh, n_azim, delta, bcs=bcs)
rt.trace_(tg)
if cap:
    L.check(tg._ctx, L.lib().rt_set_segment_capacity(tg._ctx, cap))
print("tracks", tg.n_total_tracks, "cells", model.num_cells, flush=True)
res = {}
for pipeline in (0, 3, 3, 0):
    tg.set_option("pipeline", pipeline)
    batches = []

    def on_batch(b):
        batches.append((b.uid_begin, b.uid_end, b.n_segments, b.offset_base))

    tg.timer_start()
    t0 = time.perf_counter()
    rt.segmentize_(tg, rtol=1e-6, check=False, on_batch=on_batch, fetch_volumes=False)
    ms = tg.timer_stop()
    off = tg.segment_offsets.copy()
    st = tg.segment_status.copy()
    print(f"pipeline {pipeline}: {ms:9.2f} ms  nseg {tg.n_segments} seg/s {tg.n_segments / ms * 1e3:.3e} batches {len(batches)} "
          f"count_batches {tg.info('count_batches'):.0f} fb {tg.info('verify_fallbacks'):.0f} bad {int((st != 0).sum())} phases {tg.phase_ms()}", flush=True)
    print("   ", batches[:3], "...", batches[-1], flush=True)
    if pipeline in res:
        o0, s0 = res[pipeline]
        print("    same as before:", np.array_equal(o0, off), np.array_equal(s0, st))
    res[pipeline] = (off, st)
o0, s0 = res[0]
o3, s3 = res[3]
print("offsets equal", np.array_equal(o0, o3), "status equal", np.array_equal(s0, s3))
if not np.array_equal(o0, o3):
    c0, c3 = np.diff(o0), np.diff(o3)
    d = np.nonzero(c0 != c3)[0]
    print("tracks with different counts:", d.size, d[:10], c0[d[:10]], c3[d[:10]])
if not np.array_equal(s0, s3):
    d = np.nonzero(s0 != s3)[0]
    print("tracks with different status:", d.size, d[:10], s0[d[:10]], s3[d[:10]])
