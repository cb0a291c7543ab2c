"""World-size-2 CPU test (gloo) of the multi-GPU plumbing: no device, no compute.

The path shards into independent row bands with no data-path collective; what the ranks must
agree on is the shard geometry (b200_shard_rows / the range query of b200_rms2d_tri_shard) and bench.py's
max-over-ranks timing reduction.  Each rank derives its band through the C ABI, the bands are all-gathered and must
tile the triangle exactly, in order, with balanced pair counts.
"""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nframes, out_dir):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import cpptraj_b200 as b
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        r0, r1 = b.shard_rows(nframes, rank, world)
        # the same band through the shard entry point's range query (outTri = NULL: no device needed)
        L = b.lib()
        fe, ne = C.c_size_t(0), C.c_size_t(0)
        sel = np.arange(4, dtype=np.int32)
        crd = np.zeros((nframes, 12), np.float32)
        rc = L.b200_rms2d_tri_shard(crd.ctypes.data_as(C.c_void_p), 12, nframes, None, nframes,
                                    sel.ctypes.data_as(C.c_void_p), 4, None, 1, rank, world, None, C.byref(fe), C.byref(ne))
        assert rc == 0, L.b200_last_error()
        mine = torch.tensor([r0, r1, fe.value, ne.value], dtype=torch.int64)
        allr = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, mine)
        # bench.py's timing reduction: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            np.save(os.path.join(out_dir, "bands.npy"), torch.stack(allr).numpy())
            np.save(os.path.join(out_dir, "tmax.npy"), t.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nframes", [10000, 100000, 777])
def test_two_ranks_tile_the_triangle(built, tmp_path, nframes):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), nframes, str(tmp_path)), nprocs=world, join=True)
    bands = np.load(tmp_path / "bands.npy")
    assert float(np.load(tmp_path / "tmax.npy")[0]) == float(world)
    total = nframes * (nframes - 1) // 2
    assert bands[0, 0] == 0 and bands[-1, 1] == nframes
    pos = 0
    for r in range(world):
        r0, r1, first, n = (int(x) for x in bands[r])
        assert r0 % 32 == 0 and (r1 % 32 == 0 or r1 == nframes)
        assert first == pos == nframes * r0 - r0 * (r0 + 1) // 2          # contiguous, in order, Matrix.h:110-122
        assert n == (nframes * r1 - r1 * (r1 + 1) // 2) - first
        pos += n
        if r + 1 < world:
            assert bands[r + 1, 0] == r1
    assert pos == total
    counts = bands[:, 3].astype(np.float64)
    assert counts.max() / counts.mean() < 1.0 + 64.0 * nframes / total + 1e-9   # balanced to within one 32-row group
