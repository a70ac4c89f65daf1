"""Host-side timing of the e2e call (MDEGGS_HOSTTIME=1): prep + enqueue, emit, sync; python tools/hosttime_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MDEGGS_HOSTTIME"] = "1"
import torch, bench
import multideggs_b200 as md
prob = bench._workload("cfg3").problem()
eng = md.Engine([0])
hp, hr, keep, _, _ = bench._host_problem(torch, prob)
for i in range(8):
    sys.stderr.write(f"--- call {i}\n")
    eng.run_omic_raw(hp, hr)
