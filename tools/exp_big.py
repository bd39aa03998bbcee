"""GPU experiment: one of the big named workloads (cfg4 / cfg5) through the batched fill pass, with on-device checksums of every
batch (torch over __cuda_array_interface__), for every pipeline.  usage: python tools/exp_big.py cfg4 [pipelines e.g. 0,1] [reps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
pipes = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,1").split(",")]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
t0 = time.time()
model, n_azim, delta = rt.synth.workload(name)
mesh = rt.Mesh(model)
print(f"{name}: {model.num_cells} cells, {model.num_nodes} nodes, synth+mesh {time.time() - t0:.1f} s", flush=True)
bcs = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(mesh, n_azim, delta, bcs=bcs)
rt.trace_(tg)
print("tracks", tg.n_total_tracks, "upload ms", tg.phase_ms()["upload"], flush=True)
area = rt.synth.mesh_area(model)
_lib = rt._lib if hasattr(rt, "_lib") else None
from raytracing_jl_b200 import _lib as L  # noqa: E402
L.check(tg._ctx, L.lib().rt_set_segment_capacity(tg._ctx, int(float(os.environ.get("RT_CAP", "2e9")))))

for pipe in pipes:
    tg.set_option("pipeline", pipe)
    for rep in range(reps):
        chk = {"n": 0, "len": 0.0, "elem": 0, "batches": 0, "qp": 0.0}

        def on_batch(b):
            torch.cuda.synchronize()
            cols = {k: torch.as_tensor(getattr(b, k), device="cuda") for k in ("len", "element", "qx", "px")}
            step = 1 << 27
            for lo in range(0, b.n_segments, step):
                hi = min(b.n_segments, lo + step)
                chk["len"] += float(cols["len"][lo:hi].sum(dtype=torch.float64))
                chk["elem"] += int(cols["element"][lo:hi].sum(dtype=torch.int64))
                chk["qp"] += float(cols["qx"][lo:hi].sum(dtype=torch.float64)) - float(cols["px"][lo:hi].sum(dtype=torch.float64))
            chk["n"] += b.n_segments
            chk["batches"] += 1
            torch.cuda.synchronize()

        t1 = time.time()
        tg.timer_start()
        rt.segmentize_(tg, rtol=1e-6, check=False, fetch_volumes=True, on_batch=on_batch)
        ms = tg.timer_stop()
        wall = time.time() - t1
        off = tg.segment_offsets
        st = tg.segment_status
        p = tg.phase_ms()
        print(f"pipeline {pipe} rep {rep}: segments {tg.n_segments:.4e} batches {chk['batches']} device ms {ms:.1f} (count {p['count']:.1f} fill {p['fill']:.1f}) "
              f"wall {wall:.2f} s -> {tg.n_segments / (p['count'] + p['scan'] + p['fill']) * 1e3:.3e} seg/s (kernel phases), bad tracks {(st != 0).sum()} "
              f"max seg/track {np.diff(off).max()} sum(vol)/area {tg.volumes.sum() / area:.12f} "
              f"chk n={chk['n']} len={chk['len']:.9e} elem={chk['elem']} qp={chk['qp']:.6e} fallbacks {tg.info('verify_fallbacks')}", flush=True)
