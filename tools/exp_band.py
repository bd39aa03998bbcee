"""GPU experiment: the chunk plan's boundary-band rule (k_plan_chunks: heads / tails of a track that run along the bounding box get
shorter chunks) -- threshold `band_min` (in regular chunk lengths) and shortening factor `band_div` of those chunks, against the chunk length.
usage: python tools/exp_band.py [cfg3] [reps] [chunk,band_min,band_div ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
model, n_azim, delta = rt.synth.workload(name)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(rt.Mesh(model), n_azim, delta, bcs=bcs)
rt.trace_(tg)


def run(cs, bmin, bdiv):
    tg.set_option("chunk_segments", cs)
    tg.set_option("band_min", bmin)
    tg.set_option("band_div", bdiv)
    for _ in range(3):
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
    best = None
    for _ in range(3):
        tg.timer_start()
        for _ in range(reps):
            rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
        ms = tg.timer_stop() / reps
        p = tg.phase_ms()
        if best is None or ms < best[0]:
            best = (ms, p)
    ms, p = best
    print(f"chunk {cs:5.0f} band_min {bmin:7.4f} band_div {bdiv:5.1f}: {ms:.4f} ms/step count {p['count']:.3f} fill {p['fill']:.3f} "
          f"units {tg.info('n_units'):.0f} nseg {tg.n_segments} fb {tg.info('verify_fallbacks'):.0f} bad {tg.bad_status}", flush=True)


sweep = [a.split(",") for a in sys.argv[3:]] or [("192", "0.0625", "8")]
for cs, bmin, bdiv in sweep:
    run(float(cs), float(bmin), float(bdiv))
