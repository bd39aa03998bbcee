"""Repeat the full-size cfg4 segmentation with every pipeline and compare per-batch-independent checksums of the Segment
stream (the check of tests/test_gpu_parity.py::test_cfg4_full_size_pipelines_agree, run R times on one context).
usage: python tools/stress_cfg4.py [rounds] [pipelines, e.g. 0,1,3] [GB of device memory to hold per round, e.g. 0,40,80]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracing_jl_b200 as rt  # noqa: E402
from raytracing_jl_b200 import _lib as L
from raytracing_jl_b200 import api

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
pipes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 3]
hogs = [float(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0.0]  # GB of device memory taken away per round
keys = ("px", "py", "qx", "qy", "len", "element")
model, n_azim, delta = rt.synth.workload("cfg4")
mesh = rt.Mesh(model)
B = rt.BoundaryConditions(top=rt.Reflective, bottom=rt.Reflective, right=rt.Reflective, left=rt.Reflective)
tg = rt.TrackGenerator(mesh, n_azim, delta, bcs=B)
rt.trace_(tg)
L.check(tg._ctx, L.lib().rt_set_segment_capacity(tg._ctx, 1_500_000_000))
n = tg.n_total_tracks
ref = None
for r in range(rounds):
    hog = None
    torch.cuda.empty_cache()
    if hogs[r % len(hogs)] > 0:
        hog = torch.empty(int(hogs[r % len(hogs)] * 2 ** 30), dtype=torch.uint8, device="cuda")
    print(f"round {r}: hog {hogs[r % len(hogs)]} GB, free {torch.cuda.mem_get_info()[0] / 2 ** 30:.1f} GB", flush=True)
    for pipeline in pipes:
        tg.set_option("pipeline", pipeline)
        sums, rows = {}, []

        def on_batch(b, sums=sums, rows=rows):
            torch.cuda.synchronize()
            cols = {k: torch.as_tensor(getattr(b, k), device="cuda") for k in keys}
            step = 1 << 27
            part = {}
            for lo in range(0, b.n_segments, step):
                hi = min(b.n_segments, lo + step)
                w = (torch.arange(lo, hi, device="cuda", dtype=torch.int64) + int(b.offset_base)) % 1021 + 1
                part["element"] = part.get("element", 0) + int((cols["element"][lo:hi].to(torch.int64) * w).sum())
                for k in keys[:5]:
                    part[k] = part.get(k, 0) + int((cols[k][lo:hi].view(torch.int64) & 0xFFFFFFFF).mul_(w).sum())
            for k, v in part.items():
                sums[k] = (sums.get(k, 0) + v) & ((1 << 64) - 1)  # the device sums wrap modulo 2^64
            rows.append((b.uid_begin, b.uid_end, b.n_segments, int(b.offset_base)))
            torch.cuda.synchronize()

        t0 = time.time()
        rt.segmentize_(tg, rtol=1e-6, check=False, on_batch=on_batch)
        res = (tg.n_segments, sums, tg.segment_offsets.copy(), tg.segment_status.copy())
        tag = f"round {r} pipeline {pipeline}: {time.time() - t0:.2f}s batches={len(rows)} nseg={res[0]} fallbacks={tg.info('verify_fallbacks')}"
        if ref is None:
            ref = res
            print(tag, "(reference)", flush=True)
            continue
        bad = []
        if res[0] != ref[0]:
            bad.append(f"n_segments {res[0]} != {ref[0]}")
        for k in keys:
            if res[1][k] != ref[1][k]:
                bad.append(f"sum[{k}] differs")
        if not np.array_equal(res[2], ref[2]):
            d = np.nonzero(res[2] != ref[2])[0]
            bad.append(f"offsets differ at {d.size} places, first {d[0]} ({res[2][d[0]]} vs {ref[2][d[0]]})")
        if not np.array_equal(res[3], ref[3]):
            d = np.nonzero(res[3] != ref[3])[0]
            bad.append(f"status differs at {d.size} tracks, first uid {d[0] + 1} ({res[3][d[0]]} vs {ref[3][d[0]]})")
        print(tag, "OK" if not bad else "MISMATCH: " + "; ".join(bad), flush=True)
        if bad:
            print("   batches:", rows[:12], flush=True)
