"""world_size-2 `gloo` test of the N>1 host logic (no GPU): shard planning by uid range, the unique-id broadcast
helper, and the one reduction of the path (sum of per-element delta*len over the shards == unsharded result).
The per-shard segmentation is done by the ORACLE here (test infrastructure standing in for the device)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    import raytracing_jl_b200 as rt
    from oracle.oracle import OracleMesh, OracleTrackGenerator
    from raytracing_jl_b200.api import _angle_tables
    from raytracing_jl_b200.distributed import broadcast_bytes, env_rank, plan_shards, shard_range

    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert env_rank() == (rank, world, rank)
    d = np.load(os.path.join(ROOT, "tests", "golden", "pincell.npz"))
    model = rt.UnstructuredDiscreteModel(d["node_coordinates"], d["cell_ptrs"], d["cell_data"])
    mesh = rt.Mesh(model)
    lay = rt.TrackLayout(mesh, 16, 0.02)
    _angle_tables(lay)
    bounds = plan_shards(lay, world)
    u0, u1 = shard_range(lay, rank, world)
    assert (u0, u1) == (int(bounds[rank]), int(bounds[rank + 1]))
    # every rank must plan the same split
    tb = torch.from_numpy(bounds.copy())
    ref = tb.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(tb, ref)
    # the 128-byte id travels intact
    payload = bytes(range(128)) if rank == 0 else None
    assert broadcast_bytes(payload, 128) == bytes(range(128))
    # shard-local segmentation + the single reduction of the path
    otg = OracleTrackGenerator(OracleMesh.from_mesh(mesh), 16, 0.02, bcs=(1, 1, 1, 1)).trace()
    otg.segmentize(uid_begin=u0, uid_end=u1)
    vol = torch.from_numpy(otg.volumes() * otg.n2)  # un-normalised sum(delta*len) of this shard
    nseg = torch.tensor([otg.n_segments], dtype=torch.int64)
    dist.all_reduce(vol)
    dist.all_reduce(nseg)
    np.save(os.path.join(out_dir, f"vol{rank}.npy"), vol.numpy() / otg.n2)
    np.save(os.path.join(out_dir, f"meta{rank}.npy"), np.array([u0, u1, int(nseg.item()), otg.n_total_tracks]))
    dist.destroy_process_group()


def test_two_rank_shards_reduce_to_the_whole(tmp_path, pincell_mesh):
    import torch.multiprocessing as mp

    from oracle.oracle import OracleMesh, OracleTrackGenerator

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    whole = OracleTrackGenerator(OracleMesh.from_mesh(pincell_mesh), 16, 0.02, bcs=(1, 1, 1, 1)).trace().segmentize()
    m0, m1 = np.load(tmp_path / "meta0.npy"), np.load(tmp_path / "meta1.npy")
    assert m0[0] == 1 and m0[1] == m1[0] and m1[1] == whole.n_total_tracks + 1  # contiguous cover of all uids
    assert m0[2] == m1[2] == whole.n_segments
    # shards are balanced by total track length
    lens = whole.tracks["len"]
    a, b = lens[: m0[1] - 1].sum(), lens[m0[1] - 1:].sum()
    assert abs(a - b) <= 2 * lens.max()
    v0, v1 = np.load(tmp_path / "vol0.npy"), np.load(tmp_path / "vol1.npy")
    assert np.array_equal(v0, v1)
    assert np.abs(v0 - whole.volumes()).max() <= 1e-12
    assert abs(v0.sum() - 2.56) < 1e-8


@pytest.mark.parametrize("parts", [1, 2, 3, 8, 64])
def test_shard_plan_properties(pincell_mesh, parts):
    import raytracing_jl_b200 as rt
    from raytracing_jl_b200.api import _angle_tables
    from raytracing_jl_b200.distributed import plan_shards, track_lengths

    lay = rt.TrackLayout(pincell_mesh, 32, 0.01)
    _angle_tables(lay)
    b = plan_shards(lay, parts)
    assert b[0] == 1 and b[-1] == lay.n_total_tracks + 1 and np.all(np.diff(b) >= 0) and b.size == parts + 1
    lens = track_lengths(lay)
    per = np.add.reduceat(lens, b[:-1] - 1)[: parts] if parts > 1 else np.array([lens.sum()])
    assert per.max() - per.min() <= 2 * lens.max() + 1e-9
