"""CPU, world_size 2, gloo: the N > 1 host logic — contiguous edge blocks per rank, gather of the per-rank
p-value blocks, global adjustment on the gathered vector — with the oracle as each rank's compute."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    from multideggs_b200 import _abi, sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = synth.random_problem(40, 30, 157, 2, 77)          # E not divisible by world
    E = prob.dims[2]
    lo, hi, chunk = sharding.edge_block(E, rank, world)
    sub = _abi.Problem(assay=prob.assay, from_idx=prob.from_idx[lo:hi], to_idx=prob.to_idx[lo:hi], group=prob.group,
                       K=prob.K, pct=prob.pct, method=prob.method, padj=_abi.PADJ_HOST)
    mine = pyoracle.run_omic(sub, ("p_edge", "status"))
    send = torch.full((chunk,), float("nan"), dtype=torch.float64)
    send[:hi - lo] = torch.from_numpy(mine["p_edge"])
    recv = [torch.empty(chunk, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(recv, send)                               # what ncclAllGather does on the GPUs
    p_all = sharding.assemble_blocks([r.numpy() for r in recv], E)
    full = pyoracle.run_omic(prob, ("p_edge", "cat", "padj", "sig_count", "best"))
    ok = np.array_equal(p_all, full["p_edge"], equal_nan=True)
    # global BH on the gathered vector, per (percentile, category) segment, equals the single-rank result
    from multideggs_b200 import host
    padj = np.full_like(full["padj"], np.nan)
    for pi in range(len(prob.pct)):
        for k in (1, 2):
            m = full["cat"][pi] == k
            if m.any():
                padj[pi, m] = host.p_adjust(p_all[m], "BH")
    ok = ok and np.array_equal(padj, full["padj"], equal_nan=True)
    q.put((rank, bool(ok), lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_edge_blocks_cover_and_are_contiguous():
    sys.path.insert(0, ROOT)
    from multideggs_b200 import sharding
    for E in (0, 1, 7, 53035, 7998000):
        for world in (1, 2, 3, 4, 8):
            blocks = [sharding.edge_block(E, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == E
            for (lo, hi, ch), (lo2, _, _) in zip(blocks, blocks[1:]):
                assert hi == lo2 and 0 <= hi - lo <= ch
            assert all(b[2] == blocks[0][2] for b in blocks) and blocks[0][2] * world >= E


def test_two_rank_gather_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert res[0][3] == res[1][2]
