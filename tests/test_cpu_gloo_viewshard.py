"""CPU, world_size 2 over gloo: the view-sharded build exchange.  Each rank
builds the partial mean volume of its own source views (oracle arithmetic stands in for the kernel),
one all_reduce(sum) gives every rank the full view-mean volume."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cer_oracle as O
    from cer_mvs_b200 import synth
    from cer_mvs_b200.dist import view_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    H, W, V = 48, 64, 3
    sc = synth.make_scene(H, W, V, seed=2)
    t = torch.from_numpy
    K = t(sc["intrinsics"]).clone()
    K[:, :, :2] /= 4
    D, incre = 64, 0.0025 / 64
    disp = torch.zeros(1, 1, H // 4, W // 4)
    vb, ve = view_range(V, rank, world)
    jj = list(range(1 + vb, 1 + ve))
    pyr, origin = O.build_volume(t(sc["fmaps"]), t(sc["poses"]), K, [0] * len(jj), jj, D, incre, disp, True)
    part = pyr[0].reshape(len(jj), -1, D).sum(0) / V          # local views, scaled by 1/V_total
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    if rank == 0:
        full, _ = O.build_volume(t(sc["fmaps"]), t(sc["poses"]), K, [0] * V, [1, 2, 3], D, incre, disp, True)
        want = full[0].reshape(V, -1, D).mean(0)
        out.put(float((part - want).abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_view_sharded_volume_equals_full_volume():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-6, err          # sum order only
