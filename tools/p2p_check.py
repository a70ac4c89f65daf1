#!/usr/bin/env python
"""torchrun worker (one process per GPU): edge-sharded jobs with the peer-memory all-gather (option p2p = 1) and with
ncclAllGather (p2p = 0) must both equal the same job on the rank's GPU alone, bit for bit - eager, captured, replayed.
python -m torch.distributed.run --nproc-per-node 2 tools/p2p_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import multideggs_b200 as md
from multideggs_b200 import _abi, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    uid = torch.frombuffer(bytearray(md.Engine.nccl_unique_id()), dtype=torch.uint8).to(dev)
dist.broadcast(uid, 0)
eng = md.Engine.for_rank(local, rank, world, bytes(uid.cpu().numpy().tobytes()))
solo = md.Engine([local])
want = ("p_edge", "status", "cutoff", "median", "tot_edges", "sig_count", "fail_count", "best", "cat_best", "padj_best")
ok = True
cases = [(300, 50, 2501, 2, _abi.METHOD_LM, _abi.PADJ_BH), (500, 64, 40000, 2, _abi.METHOD_LM, _abi.PADJ_BH),
         (200, 90, 3000, 3, _abi.METHOD_LM, _abi.PADJ_HOLM), (120, 60, 700, 2, _abi.METHOD_RLM, _abi.PADJ_BH),
         (300, 50, 2503, 2, _abi.METHOD_LM, _abi.PADJ_BONFERRONI)]
for G, n, E, K, method, padj in cases:                       # (growing and shrinking sizes: the buffers are re-mapped)
    prob = synth.random_problem(G, n, E, K, 31000 + E, method, padj)
    ref = solo.run_omic(prob, want)
    for p2p in (1, 0):
        eng.set_option("p2p", p2p)
        for rep in range(4):                                 # eager, eager, capture + replay, replay
            got = eng.run_omic(prob, want)
            if bool(int(eng.stats()["fit_path"]) & 0x4000) != bool(p2p):
                ok = False
                sys.stderr.write(f"[rank {rank}] case E={E} p2p={p2p} rep={rep}: fit_path {hex(int(eng.stats()['fit_path']))}\n")
            for k in want:
                if not np.array_equal(got[k], ref[k], equal_nan=True):
                    ok = False
                    sys.stderr.write(f"[rank {rank}] mismatch: case E={E} K={K} p2p={p2p} rep={rep} field {k}\n")
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("p2p_check ok" if int(flag.item()) == 1 else "p2p_check FAILED")
eng.close()
solo.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
