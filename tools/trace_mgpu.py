#!/usr/bin/env python
"""Per-launch timeline of one edge-sharded call on every rank (torchrun): python -m torch.distributed.run ... tools/trace_mgpu.py [cfg5]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
import multideggs_b200 as md

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    uid = torch.frombuffer(bytearray(md.Engine.nccl_unique_id()), dtype=torch.uint8).to(dev)
dist.broadcast(uid, 0)
eng = md.Engine.for_rank(local, rank, world, bytes(uid.cpu().numpy().tobytes()))
prob = bench._workload(name).problem()
cprob, cres, keep = bench._device_problem(torch, prob, dev)
eng.set_option("graph", 0)
for _ in range(3):
    eng.run_omic_raw(cprob, cres)
torch.cuda.synchronize()
dist.barrier()
eng.set_option("trace", 3)
eng.run_omic_raw(cprob, cres)
lines = eng.trace() if hasattr(eng, "trace") else []
for r in range(world):
    dist.barrier()
    if r == rank and (rank in (0, world // 2, world - 1)):
        sys.stderr.write(f"--- rank {rank} of {world} ({name}) ---\n")
        prev = {0: 0.0, 1: 0.0}
        for nm, sid, t in lines:
            sys.stderr.write(f"[r{rank}] {nm:34s} {'side' if sid else 'main'} {t:10.1f} {t - prev[sid]:10.1f}\n")
            prev[sid] = t
        sys.stderr.write(str({k: (round(v, 4) if isinstance(v, float) else v) for k, v in eng.stats().items()}) + "\n")
        sys.stderr.flush()
dist.barrier()
eng.close()
dist.destroy_process_group()
