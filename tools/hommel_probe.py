"""Config 3 with padj_method = "hommel" on the device: per-launch timeline (python tools/hommel_probe.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import multideggs_b200 as md
from multideggs_b200 import _abi

prob = bench._workload("cfg3").problem()
prob.padj = _abi.PADJ_HOMMEL
dev = torch.device("cuda", 0)
eng = md.Engine([0])
cprob, cres, keep = bench._device_problem(torch, prob, dev)
for _ in range(2):
    eng.run_omic_raw(cprob, cres)
eng.set_option("trace", 1)
eng.run_omic_raw(cprob, cres)
eng.set_option("trace", 0)
st = eng.stats()
print({k: round(v, 4) if isinstance(v, float) else v for k, v in st.items() if k.startswith("ms_") or k == "kernel_launches"})
print("sig_count", keep[5]["sig_count"].cpu().numpy(), "best", int(keep[5]["best"].cpu().numpy()[0]))
