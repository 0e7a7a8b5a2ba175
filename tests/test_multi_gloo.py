"""World-size-2 test of the N>1 path on CPU (gloo): sample ranges tile [0,spp),
each rank's range rendered independently, one reduce(sum) to rank 0 == the
single-process image.  The renders come from the oracle (no GPU here); what is
under test is the partition + reduce logic the GPU ranks run."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, A, O

from raytracinginoneweekendincuda_b200.multigpu import reduce_accumulators, sample_range


def test_sample_ranges_tile_exactly():
    for spp in (0, 1, 7, 10, 1024, 10000):
        for world in (1, 2, 3, 4, 8):
            rs = [sample_range(r, world, spp) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == spp
            for a, b in zip(rs, rs[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sample_range(2, 2, 8)


def _worker(rank, world, port, spp, W, H, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from raytracinginoneweekendincuda_b200 import BuiltinScene
    from oracle import bindings as O
    oracle = O.load_oracle()
    sc = BuiltinScene(10)
    cam = sc.camera(W, H, spp, 50)
    s0, s1 = sample_range(rank, world, spp)
    out = np.zeros((H, W, 3))
    st = O.oracle_stats()
    oracle.oracle_render(sc.desc, C.byref(cam), s0, s1, 1984, 1, 64, 1, out.ctypes.data, C.byref(st))
    acc = torch.from_numpy(out.astype(np.float32))
    rays = torch.tensor([st.rays], dtype=torch.int64)
    reduce_accumulators(acc, dst=0)
    dist.reduce(rays, dst=0)
    if rank == 0:
        q.put((acc.numpy(), int(rays.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_split_equals_single_render(oracle):
    spp, W, H = 5, 32, 18
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, spp, W, H, q)) for r in range(2)]
    for p in procs:
        p.start()
    acc, rays = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from raytracinginoneweekendincuda_b200 import BuiltinScene
    sc = BuiltinScene(10)
    cam = sc.camera(W, H, spp, 50)
    full = np.zeros((H, W, 3))
    st = O.oracle_stats()
    oracle.oracle_render(sc.desc, C.byref(cam), 0, spp, 1984, 1, 64, 2, full.ctypes.data, C.byref(st))
    assert rays == st.rays
    assert np.allclose(acc, full.astype(np.float32), rtol=1e-6, atol=1e-7)
