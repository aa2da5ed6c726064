"""World-size-2 gloo test (CPU) of the multi-GPU plumbing: round-robin sharding + all_gather of partial points +
order-independent summation reproduce the unsharded MSM.  The per-rank partial MSMs are computed
by the CPU oracle here; on GPUs the same plumbing carries g16_prove_partial outputs (tests/test_gpu_parity.py
::test_sharded_prove_equals_single checks the CUDA side, bench.py --gpus N the NCCL side)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orc
import pyref as P
from groth16_b200 import CurveCodec, get_curve
from groth16_b200.dist import all_gather_partials, shard_indices


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, bases, scalars, nq, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sel = shard_indices(bases.shape[0], rank, world)
    part = orc.msm_g1(1, nq, np.ascontiguousarray(bases[sel]), np.ascontiguousarray(scalars[sel]), threads=1)   # X||Y||Z
    allp = all_gather_partials(part)
    if rank == 0:
        q.put(allp)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_msm_over_gloo():
    c = P.CURVES["bn254"]
    cx = P.ctx(c)
    cd = CurveCodec(get_curve("bn254"))
    rng = P.Rng(3)
    n = 37
    pts = [cx.G1.mul(cx.g1_gen(), rng.fr(c.r)) for _ in range(n)]
    sc = [rng.fr(c.r) for _ in range(n)]
    bases, scalars = cd.enc_g1(pts), cd.fr.bigint(sc)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, bases, scalars, cd.nq, q)) for r in range(world)]
    for p in procs:
        p.start()
    allp = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert allp.shape == (world, 3 * cd.nq)
    total = None
    for r in range(world):
        total = cx.G1.add(total, cd.dec_proj_g1(allp[r]))
    assert total == cx.G1.msm_naive(pts, sc)
    # the shares tile [0, n) exactly, like Engine::shard
    assert sorted(i for r in range(3) for i in range(n)[shard_indices(n, r, 3)]) == list(range(n))
