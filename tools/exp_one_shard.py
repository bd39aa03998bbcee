"""GPU experiment helper: run segmentize on ONE uid shard (for ncu).  usage: python tools/exp_one_shard.py cfg3 rank world div"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

name, r, world, div = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
model, n_azim, delta = rt.synth.workload(name)
mesh = rt.Mesh(model)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(mesh, n_azim, delta / div, bcs=bcs, shard=(r, world))
rt.trace_(tg)
for _ in range(3):
    rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
print(tg.n_segments, tg.phase_ms(), tg.stats())
