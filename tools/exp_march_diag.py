"""GPU experiment: where the lane-slots of k_march's fast loop go (needs a library built with -DRT_MARCH_DIAG, which reports them
through the literal-iteration / query counters): finished lanes that wait for the slowest walker of their warp, lanes that wait
for the slow side, lanes that step.  usage: RT_B200_LIB=build_variants/diag.so python tools/exp_march_diag.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt
model, n_azim, delta = rt.synth.workload("cfg3")
tg = rt.TrackGenerator(rt.Mesh(model), n_azim, delta, bcs=rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective))
rt.trace_(tg)
for a in sys.argv[1:]:
    k, v = a.split('=')
    tg.set_option(k, float(v))
rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
st = tg.stats()
tot = st["knn_queries"]
print("lane-slots", tot, "done-idle %.3f" % (st["literal_iterations"] / tot), "wait-idle %.3f" % (st["nn_queries"] / tot), "fast %.3f" % (st["fast_transitions"] / tot), "segments", tg.n_segments)
