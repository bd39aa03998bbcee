"""GPU experiment: which tracks fail the length check on a big workload, per pipeline."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt
from raytracing_jl_b200 import _lib as L
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
pipes = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,1").split(",")]
model, n_azim, delta = rt.synth.workload(name)
mesh = rt.Mesh(model)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(mesh, n_azim, delta, bcs=bcs)
rt.trace_(tg)
L.check(tg._ctx, L.lib().rt_set_segment_capacity(tg._ctx, int(float(os.environ.get("RT_CAP", "3e9")))))
base = tg._base
for pipe in pipes:
    tg.set_option("pipeline", pipe)
    rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
    off = tg.segment_offsets
    st = tg.segment_status
    bad = np.nonzero(st)[0]
    print("pipeline", pipe, "segments", tg.n_segments, "bad", bad.size, "status values", np.unique(st[bad], return_counts=True), flush=True)
    if bad.size:
        az = np.searchsorted(base, bad, side="right") - 1
        print("  bad per angle (first 40 nonzero):", [(int(a), int(c)) for a, c in zip(*np.unique(az, return_counts=True))][:40])
        print("  first bad uids:", (bad[:10] + 1).tolist(), "counts", np.diff(off)[bad[:10]].tolist(), "len", tg.track_data["len"][bad[:10]].tolist())
        np.save("gpurun_out/bad_%s_p%d.npy" % (name, pipe), bad[:200000])
