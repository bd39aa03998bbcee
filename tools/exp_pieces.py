"""GPU experiment: the steady-state segmentize! in K pieces (option "pieces"), checksum against one piece.
usage: [RT_B200_LIB=build_variants/x.so] python tools/exp_pieces.py [cfg3] [reps] [K list]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 3, 4, 6, 8]
model, n_azim, delta = rt.synth.workload(name)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(rt.Mesh(model), n_azim, delta, bcs=bcs)
rt.trace_(tg)


def checksum():
    s = tg.fetch_segments(pinned=True)
    w = (np.arange(s["len"].shape[0], dtype=np.int64) % 1021 + 1)
    return tuple(int((s[k].view(np.int64 if s[k].dtype.itemsize == 8 else np.int32).astype(np.int64) & 0xFFFFFFFF).dot(w)) for k in
                 ("px", "py", "qx", "qy", "len", "element")) + (int(tg.segment_offsets[-1]), int(np.count_nonzero(tg.segment_status)),
                                                                float(tg.volumes.sum()))


ref = None
for K in ks:
    tg.set_option("pieces", K)
    for _ in range(5):
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
    tg.timer_start()
    for _ in range(reps):
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=False)
    ms = tg.timer_stop() / reps
    p = tg.phase_ms()
    c = checksum()
    if ref is None:
        ref = c
    same = c[:8] == ref[:8] and abs(c[8] - ref[8]) < 1e-9 * abs(ref[8])
    if not same:
        print("   differs:", [i for i in range(9) if c[i] != ref[i]], c[8], ref[8])
    print(f"pieces {K}: {ms:7.4f} ms/step  count {p['count']:6.3f} scan {p['scan']:6.3f} fill {p['fill']:6.3f}  nseg {tg.n_segments} "
          f"fb {tg.info('verify_fallbacks'):.0f} same-as-first {same}", flush=True)
