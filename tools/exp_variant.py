"""GPU experiment: steady-state phase times of one library build on a named workload, with an order- and bit-sensitive checksum of
every Segment column (variants of a kernel must agree on it).
usage: [RT_B200_LIB=build_variants/x.so] python tools/exp_variant.py [cfg3] [reps] [chunk_segments ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
chunks = [float(a) for a in sys.argv[3:]] or [None]
model, n_azim, delta = rt.synth.workload(name)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(rt.Mesh(model), n_azim, delta, bcs=bcs)
rt.trace_(tg)
lib = os.path.basename(os.environ.get("RT_B200_LIB", "default"))


def checksum():
    tg.fetch_segments()
    s = tg.segments
    w = (np.arange(s["len"].shape[0], dtype=np.int64) % 1021 + 1)
    return hash(tuple(int((s[k].view(np.int64 if s[k].dtype.itemsize == 8 else np.int32).astype(np.int64) & 0xFFFFFFFF).dot(w)) for k in
                      ("px", "py", "qx", "qy", "len", "element"))) & 0xFFFFFFFF


for cs in chunks:
    if cs is not None:
        tg.set_option("chunk_segments", cs)
    for _ in range(4):
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
    best = None
    for _ in range(3):  # best of three blocks of `reps` steps
        tg.timer_start()
        for _ in range(reps):
            rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
        ms = tg.timer_stop() / reps
        p = tg.phase_ms()
        if best is None or ms < best[0]:
            best = (ms, p)
    ms, p = best
    ck = checksum() if cs is None or cs == chunks[0] else 0
    print(f"{lib:24s} chunk {cs}: {ms:.4f} ms/step count {p['count']:.3f} fill {p['fill']:.3f} units {tg.info('n_units'):.0f} "
          f"nseg {tg.n_segments} fb {tg.info('verify_fallbacks'):.0f} bad {tg.bad_status} cks {ck:08x}", flush=True)
