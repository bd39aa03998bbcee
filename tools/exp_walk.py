"""GPU experiment: the walk of the single-walk pipeline (k_seed + k_march) on a named workload, checksum against the hybrid pipeline.
usage: [RT_B200_LIB=build_variants/x.so] python tools/exp_walk.py [cfg3] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
model, n_azim, delta = rt.synth.workload(name)
mesh = rt.Mesh(model)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(mesh, n_azim, delta, bcs=bcs)
rt.trace_(tg)
print("lib", os.environ.get("RT_B200_LIB", "default"), "tracks", tg.n_total_tracks, "cells", model.num_cells, flush=True)


def checksum():
    s = tg.fetch_segments(pinned=True)
    w = (np.arange(s["len"].shape[0], dtype=np.int64) % 1021 + 1)
    return tuple(int((s[k].view(np.int64 if s[k].dtype.itemsize == 8 else np.int32).astype(np.int64) & 0xFFFFFFFF).dot(w)) for k in
                 ("px", "py", "qx", "qy", "len", "element")) + (int(tg.segment_offsets[-1]), int(np.count_nonzero(tg.segment_status)))


def run(label, chk=False, **opts):
    for k, v in opts.items():
        tg.set_option(k, v)
    best = None
    for _ in range(reps):
        tg.timer_start()
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
        ms = tg.timer_stop()
        p = tg.phase_ms()
        if best is None or p["count"] < best[1]["count"]:
            best = (ms, p)
    print(f"{label:40s} total {best[0]:7.3f} ms  count {best[1]['count']:6.3f} scan {best[1]['scan']:6.3f} fill {best[1]['fill']:6.3f}"
          f"  nseg {tg.n_segments} fb {tg.info('verify_fallbacks'):.0f} bad {tg.bad_status}", flush=True)
    return checksum() if chk else None


c0 = run("hybrid (pipeline 0)", chk=True, pipeline=0)
c3 = run("single walk (pipeline 3)", chk=True, pipeline=3)
print("checksums equal:", c0 == c3, flush=True)
if os.environ.get("RT_EXP_SWEEP"):
    for cs in (64, 96, 160, 192):
        run(f"single walk chunk={cs}", pipeline=3, chunk_segments=cs)
