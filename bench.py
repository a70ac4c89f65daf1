#!/usr/bin/env python
"""bench.py — throughput of the multiDEGGs differential-correlation hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference (rank 0 only)

A "step" is one whole-job pass of the hot path over one omic layer (node filter -> gather -> gene stats ->
quantiles -> pair fits -> tails -> percolation -> BH -> best percentile) on synthetic data of BASELINE.json's
config 3 (20,000 genes x 1,000 samples, 3 categories, built-in 53,035-edge network, BH).
  value  : unique network edges tested per second, inputs already resident in HBM (device pointers).
  e2e    : the same through the C-ABI call with HOST buffers (pinned), H2D + D2H inside the timed region.
At N > 1 the main line is weak scaling over independent omic layers (one layer per GPU, the reference's
`lapply` over omics, core_functions.R:214; no data-path collective); the edge-sharded + NCCL all-gather path
of config 5 (4,000 x 2,000, all 7,998,000 pairs) is reported beside it under "also".
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if "NCCL_DEBUG" not in os.environ:  # (a pre-set level, e.g. the driver's INFO, is kept)
    os.environ["NCCL_DEBUG"] = "WARN"
# The contract is ONE JSON line on stdout.  Native libraries (NCCL's version banner) write to file descriptor 1
# behind Python's back, so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor.
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def _emit(line: dict) -> None:
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


METRIC = "gene-pair interaction tests/sec"
UNIT = "unique pair tests/s"


# ----------------------------------------------------------------------------------------------
def _peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "MEASURED_PEAKS.json (of measured)"}
    return {"hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


def _dram_traffic() -> dict:
    """Per-kernel DRAM bytes per launch from the committed ncu --set full capture (profiles/dram_traffic.json, written by
    tools/dram_traffic.py): read at run time so the number follows the profile, never a literal in this file."""
    path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh)
    except (OSError, ValueError):
        return {}


def _config(name: str, prob, st_nodes, world: int, extra: dict | None = None) -> dict:
    """The workload description both arms print (identical keys)."""
    G, n, E, K, P = prob.dims
    cfg = {"workload": name, "genes": int(G), "samples": int(n), "edges": int(E), "categories": int(K),
           "percentiles": int(P), "padj": "BH", "network_nodes": st_nodes, "parallelism": None, "l2": None,
           "timing": None, "launch_mode": None, "host_cpus": None, "note": None}
    cfg.update(extra or {})
    return cfg


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 2 ms from a thread (the timed
    region of the value leg is ~10 ms, too short for nvidia-smi's 100 ms loop, which is only the fallback)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.handle, self.samples, self.stop_flag = None, None, [], False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                self.samples.append((sm, mx, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            n = self.nvml
            names = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown,
                     "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown,
                     "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
            reasons = sorted(nm for nm, bit in names.items() if any(rs & bit for _, _, rs in self.samples))
            sm = [x[0] for x in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None,
                    "sm_max_mhz": float(max(x[1] for x in self.samples)) if sm else None, "reasons": reasons,
                    "samples": len(sm), "source": "NVML polled every 2 ms during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def _workload(name: str, seed_offset: int = 0):
    from multideggs_b200 import synth
    if name == "cfg3":
        return synth.cfg3(seed=20261019 + seed_offset)
    if name == "cfg3-small":
        return synth.cfg3(n_genes=4000, n_samples=200, max_edges=6000, seed=20261019 + seed_offset)
    if name == "cfg5":
        return synth.cfg5()
    if name == "cfg5-small":
        return synth.cfg5(n_features=600, n_samples=400)
    raise SystemExit(f"unknown workload {name}")


WANT = ("cutoff", "p_edge", "status", "tot_edges", "sig_count", "fail_count", "best", "cat_best", "padj_best")


# ----------------------------------------------------------------------------------------------
# reference arm: CPU restatement of the reference algorithm on the host cores
# ----------------------------------------------------------------------------------------------
def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1: the reference arm uses every host core this process may run on
    host_cpus = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(host_cpus)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    name = args.workload or "cfg3"
    wl = _workload(name)
    prob = wl.problem()
    E = prob.dims[2]
    want = ("tot_edges", "sig_count", "best")  # no per-edge outputs => pure reference mode (refit per percentile)
    cores = pyoracle.lib().mdeggs_oracle_threads()
    for _ in range(max(args.warmup, 0)):
        pyoracle.run_omic(prob, want, pyoracle.MODE_REFERENCE, 0)
    t0 = time.perf_counter()
    tests = 0
    n_nodes = None
    for _ in range(args.steps):
        r = pyoracle.run_omic(prob, want, pyoracle.MODE_REFERENCE, 0)
        tests += int(r["tot_edges"].sum())
    dt = time.perf_counter() - t0
    value = E * args.steps / dt
    used = min(cores, len(prob.pct))
    n_nodes = int(np.unique(np.concatenate([prob.from_idx, prob.to_idx])).size)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(name, prob, n_nodes, 1, {
            "parallelism": f"{used} host threads over the percentiles (like mclapply), {cores} available; rank 0 only",
            "l2": "n/a (host)", "timing": "wall clock around whole jobs; at most 5 steps and 1 warm-up (each step is a "
                                          "whole job of seconds on the host cores)",
            "launch_mode": "n/a (host)", "host_cpus": f"{host_cpus} CPUs in this process's affinity mask",
            "note": "R is not installed here: CPU restatement of the reference algorithm (oracle/), reference mode = "
                    "every (percentile, category, edge) refitted, parallel over percentiles only like mclapply"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": f"full {name} workload per step, {args.steps} steps, host threads available {cores}"},
        "ref_equiv_tests_per_s": tests / dt,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


# ----------------------------------------------------------------------------------------------
# CUDA arm
# ----------------------------------------------------------------------------------------------
def _device_problem(torch, prob, dev):
    """Inputs resident in HBM: torch tensors own the memory, the C struct carries their pointers."""
    from multideggs_b200 import _abi
    a = prob.assay
    assert a.flags.f_contiguous
    t_assay = torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)  # memory == column-major G x n
    t_from = torch.from_numpy(np.ascontiguousarray(prob.from_idx, dtype=np.int32)).to(dev)
    t_to = torch.from_numpy(np.ascontiguousarray(prob.to_idx, dtype=np.int32)).to(dev)
    G, n, E, K, P = prob.dims
    group = np.ascontiguousarray(prob.group, dtype=np.int32)
    pct = np.ascontiguousarray(prob.pct, dtype=np.float64)
    cprob = _abi.CProblem(t_assay.data_ptr(), G, G, n, _abi.LAYOUT_COLMAJOR, _abi.MEM_DEVICE, t_from.data_ptr(),
                          t_to.data_ptr(), E, group.ctypes.data, K, pct.ctypes.data, P, prob.method, prob.padj,
                          prob.tested_coef, prob.rlm_cov_mode)
    outs = {}
    cres = _abi.CResult()
    for nm in WANT:
        dt, shp = _abi.RESULT_FIELDS[nm]
        tdt = {np.float64: torch.float64, np.int32: torch.int32, np.int8: torch.int8, np.int64: torch.int64}[dt]
        t = torch.zeros(shp(G, n, E, K, P), dtype=tdt, device=dev)
        outs[nm] = t
        setattr(cres, nm, t.data_ptr())
    keep = (t_assay, t_from, t_to, group, pct, outs)
    return cprob, cres, keep


def _host_problem(torch, prob):
    """Pinned host buffers for the e2e leg."""
    from multideggs_b200 import _abi
    G, n, E, K, P = prob.dims
    pin = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    t_assay = pin(prob.assay.T)
    t_from = pin(np.asarray(prob.from_idx, dtype=np.int32))
    t_to = pin(np.asarray(prob.to_idx, dtype=np.int32))
    group = np.ascontiguousarray(prob.group, dtype=np.int32)
    pct = np.ascontiguousarray(prob.pct, dtype=np.float64)
    cprob = _abi.CProblem(t_assay.data_ptr(), G, G, n, _abi.LAYOUT_COLMAJOR, _abi.MEM_HOST, t_from.data_ptr(),
                          t_to.data_ptr(), E, group.ctypes.data, K, pct.ctypes.data, P, prob.method, prob.padj,
                          prob.tested_coef, prob.rlm_cov_mode)
    outs = {}
    cres = _abi.CResult()
    d2h = 0
    for nm in WANT:
        dt, shp = _abi.RESULT_FIELDS[nm]
        tdt = {np.float64: torch.float64, np.int32: torch.int32, np.int8: torch.int8, np.int64: torch.int64}[dt]
        t = torch.zeros(shp(G, n, E, K, P), dtype=tdt).pin_memory()
        outs[nm] = t
        d2h += t.numel() * t.element_size()
        setattr(cres, nm, t.data_ptr())
    h2d = t_assay.numel() * 8 + 2 * E * 4
    return cprob, cres, (t_assay, t_from, t_to, group, pct, outs), h2d, d2h


def _timed_steps(torch, eng, cprob, cres, steps, warmup, flush, dist_sync):
    """W warm-up steps, then K steps; device time per step by CUDA events on the (legacy) default stream, which
    orders against the engine's blocking streams; L2 flushed (untimed) between steps."""
    for _ in range(warmup):
        eng.run_omic_raw(cprob, cres)
    dist_sync()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    stage = {}
    t_wall0 = time.perf_counter()
    for i in range(steps):
        if flush is not None:
            flush.add_(1.0)  # touches 2 x 256 MiB > 126 MB L2
        ev[i][0].record()
        eng.run_omic_raw(cprob, cres)
        ev[i][1].record()
        st = eng.stats()
        for k, v in st.items():
            if k.startswith("ms_"):
                stage[k] = stage.get(k, 0.0) + v
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_wall0
    dist_sync()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    return ms, wall, {k: v / steps for k, v in stage.items()}, st


def _bind_near_gpu(local: int) -> str:
    """One process per GPU: run on the CPUs NVML reports as local to this rank's GPU, so that the pinned host buffers
    (allocated afterwards, first touch) live on the GPU's NUMA node and every rank's host<->device copies use its own
    memory controllers instead of all ranks sharing those of one socket. Returns a description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "unchanged (NVML reports no local CPUs inside this process's CPU set)"
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} CPUs local to the GPU (NVML cpu affinity)"
    except Exception as e:  # no NVML, no permission: keep the inherited CPU set
        return f"unchanged ({type(e).__name__})"


def _kernel_table(torch, eng, cprob, cres, flush, steps: int) -> list[dict]:
    """Per-launch device times of the eager pipeline from CUDA events on the launching streams (engine option "trace" =
    3: an event after every launch, no printing): duration = this event - the later of (previous event on the same
    stream, the main-stream event the side stream forked from).  Median over `steps` steps, L2 flushed between them."""
    eng.set_option("graph", 0)
    eng.set_option("trace", 3)
    runs = []
    for _ in range(steps + 1):
        flush.add_(1.0)
        eng.run_omic_raw(cprob, cres)
        runs.append(eng.trace())
    eng.set_option("trace", 0)
    eng.set_option("graph", 1)
    runs = runs[1:]
    table = []
    for i, (name, stream, _) in enumerate(runs[0]):
        durs = []
        for tr in runs:
            if len(tr) != len(runs[0]):
                continue
            prev_same = max([t for (_, s, t) in tr[:i] if s == stream], default=0.0)
            fork = max([t for (_, s, t) in tr[:i] if s == 0], default=0.0) if stream == 1 else 0.0
            durs.append(tr[i][2] - max(prev_same, fork))
        table.append({"kernel": name.strip("()").split("<")[0], "stream": "side" if stream else "main",
                      "us": float(np.median(durs)) if durs else None})
    return table


def _kernel_bytes(G_nodes: int, n: int, npad: int, E: int, K: int, tot_edges_sum: int) -> dict:
    """Algorithmic bytes per launch of the kernels of a config-3-shaped step."""
    vals = G_nodes * n
    return {
        "k_mark_nodes": ("hbm", 8 * E), "k_build_nodes": ("hbm", 16 * E),
        # fused K1 + K2: node rows read once (8 B / value), D (8 B) and the bin ids (2 B) written once
        "k_front_fused": ("hbm", 8 * vals + 10 * G_nodes * npad),
        "k_gather_colmajor": ("hbm", 8 * vals + 10 * G_nodes * npad), "k_gene_stats_reg": ("hbm", 10 * G_nodes * npad),
        "k_qs_compact": ("hbm", 2 * G_nodes * npad),
        # pair fits: rows A and B, two indices, the K cross moments; D is L2-resident => L2 bandwidth is the bound
        "k_fit_stream": ("l2", E * (16 * n + 16)),
        "k_finish": ("hbm", E * (8 * K + 8 + 8 + 4 + 8)),
        "k_percolate": ("hbm", E * (8 + 4 + 8)),
        "k_bsort_scatter": ("hbm", E * (8 + 4 + 12)), "k_bsort_sort": ("hbm", E * (12 + 8 + 8 + 20)),
        "k_bh_onepass": ("hbm", 24 * tot_edges_sum), "k_bh_count": ("hbm", 8 * tot_edges_sum),
        "k_bh_tilemin": ("hbm", 8 * tot_edges_sum), "k_bh_final": ("hbm", 8 * tot_edges_sum),
    }


def _bundled():
    """The reference's bundled example data (data/*.rda re-encoded by tests/golden/make_golden.py)."""
    import pandas as pd
    d = np.load(os.path.join(ROOT, "tests", "golden", "bundled.npz"), allow_pickle=False)
    out = {}
    for key, nm in (("rnaseq", "RNAseq"), ("proteomic", "Proteomics"), ("olink", "Olink")):
        out[nm] = pd.DataFrame(d[f"{key}_x"], index=d[f"{key}_genes"], columns=d[f"{key}_samples"])
    meta = pd.DataFrame({"response": pd.Categorical(d["meta_response"], categories=list(d["meta_response_levels"]))},
                        index=d["meta_samples"])
    return out, meta


def _also_small_configs(md, steps: int) -> dict:
    """BASELINE configs 1, 2 and 4 through the public API (host data frames in, deggs object out): wall clock per call,
    everything included (host prep, H2D, engine, D2H, data.frame assembly)."""
    import pandas as pd
    from multideggs_b200 import synth
    out = {}
    assays, meta = _bundled()
    for key, method, padj in (("cfg1", "lm", "bonferroni"), ("cfg2", "rlm", "BH")):
        def call():
            return md.get_diffNetworks(assayData=assays, metadata=meta, category_variable="response",
                                       regression_method=method, padj_method=padj, verbose=False, show_progressBar=False,
                                       cores=2)
        res = call()
        for _ in range(5):
            call()
        ts = []
        for _ in range(max(20, steps)):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        edges = 0
        for nm in assays:
            prep_net = res["diffNetworks"].get(nm)
            edges += 0 if prep_net is None else sum(len(v) for k, v in prep_net.items()
                                                    if isinstance(v, pd.DataFrame))
        out[key] = {"workload": f"bundled RNAseq + Proteomics + Olink, {method} + {padj} (BASELINE config {key[-1]})",
                    "api": "md.get_diffNetworks (three omic layers in flight, mdeggs_run_omics)",
                    "ms_per_call_median": 1e3 * float(np.median(ts)), "ms_per_call_min": 1e3 * float(np.min(ts)),
                    "calls": len(ts), "specific_links_reported": int(edges)}
    wl = synth.cfg4()
    x = pd.DataFrame(wl.assay.values.T, index=wl.assay.samples, columns=wl.assay.genes)
    y = wl.group.astype(float)
    rng = np.random.default_rng(20261020)
    folds = [[] for _ in range(10)]
    for c in (1, 2):  # stratified folds, as tests/test_gpu_parity.py::test_cfg4_full_size_ten_folds
        idx = rng.permutation(np.flatnonzero(wl.group == c))
        for j, smp in enumerate(idx.tolist()):
            folds[(j + (c - 1) * 5) % 10].append(smp)

    def run_all():
        npairs = []
        for hold in [None] + folds:
            keep = np.ones(len(y), dtype=bool)
            if hold is not None:
                keep[np.asarray(hold)] = False
            try:
                r = md.multiDEGGs_filter(y[keep], x.iloc[keep])
                npairs.append(int(r["pairs"].shape[0]))
            except ValueError:
                npairs.append(0)
        return npairs
    run_all()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        npairs = run_all()
        ts.append(time.perf_counter() - t0)
    E4 = int(wl.problem().dims[2])
    out["cfg4"] = {"workload": "5,000 features x 24 samples, 15 percentiles, q.value, 10 stratified folds + the full data "
                               "(BASELINE config 4)", "api": "md.multiDEGGs_filter, 11 calls",
                   "ms_per_11_calls": 1e3 * float(np.min(ts)), "edges_per_call": E4,
                   "value": 11 * E4 / float(np.min(ts)), "unit": UNIT, "pairs_kept_per_call": npairs}
    return out


def run_cuda(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_binding = _bind_near_gpu(local) if world > 1 and not os.environ.get("MDEGGS_NO_BIND") else "not bound (N = 1)"
    import torch
    import multideggs_b200 as md
    from multideggs_b200 import _abi

    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    def dist_sync():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    name = args.workload or "cfg3"
    wl = _workload(name, seed_offset=rank)  # N > 1: one independent omic layer per GPU (weak scaling)
    prob = wl.problem()
    G, n, E, K, P = prob.dims
    eng = md.Engine([local])
    flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    # ---- value: inputs resident in HBM ----
    cprob, cres, keep = _device_problem(torch, prob, dev)
    sampler = ClockSampler(local)
    sampler.start()
    ms, wall, _, st = _timed_steps(torch, eng, cprob, cres, args.steps, args.warmup, flush, dist_sync)
    clocks = sampler.stop()
    graph_replayed = bool(int(st["fit_path"]) & 0x100)
    # per-stage / per-kernel CUDA-event times need stage boundaries: a few extra steps launched eagerly
    eng.set_option("graph", 0)
    _, _, stage, _ = _timed_steps(torch, eng, cprob, cres, max(3, args.steps // 2), 1, flush, dist_sync)
    eng.set_option("graph", 1)
    ms = max_over_ranks(ms)
    value = world * E * args.steps / (ms * 1e-3)
    tot_edges = keep[5]["tot_edges"].cpu().numpy()
    ref_equiv = int(tot_edges.sum())
    best = int(keep[5]["best"].cpu().numpy()[0])
    sig = keep[5]["sig_count"].cpu().numpy()

    # ---- e2e: host buffers through the C ABI ----
    hprob, hres, hkeep, h2d, d2h = _host_problem(torch, prob)
    if os.environ.get("MDEGGS_ZERO_COPY"):
        eng.set_option("zero_copy", int(os.environ["MDEGGS_ZERO_COPY"]))
    for _ in range(args.warmup):
        eng.run_omic_raw(hprob, hres)
    dist_sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.run_omic_raw(hprob, hres)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    dist_sync()
    e2e_value = world * E * args.steps / e2e_s
    e2e_stats = eng.stats()
    e2e_stage = {k: v for k, v in e2e_stats.items() if k.startswith("ms_")}
    h2d, d2h = int(e2e_stats["h2d_bytes"]), int(e2e_stats["d2h_bytes"])  # bytes the engine actually moved per step
    assert int(hkeep[5]["best"].numpy()[0]) == best, "host and device legs disagree"
    # the same job with the gene rows shuffled: the node rows are then scattered over all G rows of the host matrix
    # instead of leading it and the band of node rows is the whole matrix; the engine's host threads gather the node rows
    # into pinned staging chunk by chunk (option host_pack) and only those 67 MB cross PCIe, behind the enqueued compute
    e2e_shuf = None
    e2e_page = None
    if rank == 0 and world == 1:
        rng_s = np.random.default_rng(99)
        perm = rng_s.permutation(G)                       # new row i holds old row perm[i]
        inv = np.empty(G, dtype=np.int64)
        inv[perm] = np.arange(G)
        from multideggs_b200 import _abi as _A
        sprob = _A.Problem(assay=np.asfortranarray(prob.assay[perm]), from_idx=(inv[prob.from_idx - 1] + 1).astype(np.int32),
                           to_idx=(inv[prob.to_idx - 1] + 1).astype(np.int32), group=prob.group, K=K, pct=prob.pct,
                           method=prob.method, padj=prob.padj, tested_coef=prob.tested_coef, rlm_cov_mode=prob.rlm_cov_mode)
        sp, sr, sk, _, _ = _host_problem(torch, sprob)
        for _ in range(args.warmup):
            eng.run_omic_raw(sp, sr)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.run_omic_raw(sp, sr)
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
        st_s = eng.stats()
        assert int(sk[5]["best"].numpy()[0]) == best, "shuffled rows change the result"
        e2e_shuf = {"value": E * args.steps / dt_s, "unit": UNIT, "ms_per_step": 1e3 * dt_s / args.steps,
                    "h2d_bytes_per_step": int(st_s["h2d_bytes"]), "d2h_bytes_per_step": int(st_s["d2h_bytes"]),
                    "host_packed": bool(int(st_s["fit_path"]) & 0x2000),
                    "note": "gene rows of the host matrix shuffled: node rows scattered over all 20,000 rows; the engine's "
                            "host threads pack the node rows into pinned staging in row chunks, each chunk is sent while "
                            "the next is packed and the front kernel consumes them as they land (option host_pack)"}
        del sp, sr, sk, sprob
        # the matrix in PAGEABLE host memory (what R hands to the .Call shim: an R matrix is never pinned)
        from multideggs_b200 import _abi as _A2
        pc = prob.to_c()
        pr, parr = _A2.alloc_result(prob.dims, ("p_edge", "stat", "status", "tot_edges", "sig_count", "best", "cat_best",
                                                "padj_best"))
        e2e_page = {"unit": "ms per step", "note": "same job, assay / edge list / results in pageable numpy memory"}
        for label, hp in (("host_pack_0_driver_staged_copy", 0), ("host_pack_1", 1)):
            eng.set_option("host_pack", hp)
            for _ in range(args.warmup):
                eng.run_omic_raw(pc, pr)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                eng.run_omic_raw(pc, pr)
            torch.cuda.synchronize()
            e2e_page[label] = 1e3 * (time.perf_counter() - t0) / args.steps
            assert int(parr["best"][0]) == best, "pageable leg disagrees"
        eng.set_option("host_pack", 1)
        del pc, pr, parr

    # ---- roofline: every kernel of the step timed by CUDA events on its own stream; the headline entry is the kernel
    # with the largest time.  Denominators: HBM = MEASURED_PEAKS.json; L2 = measured here, in this run ----
    peaks = _peaks()
    l2_peak = eng.l2_peak_gbs(64 << 20)
    g_nodes = int(st["n_nodes"])
    npad = sum((int(c) + 3) // 4 * 4 for c in np.bincount(prob.group, minlength=K + 1)[1:])
    tot_sum = int(keep[5]["tot_edges"].cpu().numpy().sum())
    ktab = _kernel_table(torch, eng, cprob, cres, flush, max(3, args.steps // 4))
    kbytes = _kernel_bytes(g_nodes, n, npad, E, K, tot_sum)
    dram = _dram_traffic()
    per_launch = dram.get("per_launch", {})
    kernels = []
    for row in ktab:
        nm, us = row["kernel"], row["us"]
        if nm not in kbytes or not us or us <= 0:
            continue
        bound, nbytes = kbytes[nm]
        pk = l2_peak if bound == "l2" else peaks["hbm_gbs"]
        ach = nbytes / (us * 1e-6) / 1e9
        tr = per_launch.get(nm)
        kernels.append({"kernel": nm, "stream": row["stream"], "us": us, "algorithmic_bytes": int(nbytes), "bound": bound,
                        "achieved": ach, "peak": pk, "frac": ach / pk,
                        "traffic": (tr["dram_read_bytes"] + tr["dram_write_bytes"]) if tr else None})
    top = max(kernels, key=lambda k: k["us"]) if kernels else None
    fit_bytes = E * (16 * n + 16)
    # whole-step algorithmic bytes (SURVEY section 8 d): gather 2*8*G_nodes*n + quantiles 4*8*G_nodes*n + pair fits
    # + adjust 3*8*sum_p tot_edges[p]
    pipe_bytes = 6 * 8 * g_nodes * n + fit_bytes + 24 * tot_sum
    step_s = ms / args.steps * 1e-3
    roofline = {
        "bound": top["bound"] if top else None, "kernel": top["kernel"] if top else None,
        "achieved": top["achieved"] if top else None, "peak": top["peak"] if top else None, "unit": "GB/s",
        "frac": top["frac"] if top else None, "traffic": top["traffic"] if top else None,
        "algorithmic_bytes_per_launch": top["algorithmic_bytes"] if top else None,
        "kernel_us": top["us"] if top else None,
        "selection": "the kernel with the largest CUDA-event time in the step (table under `kernels`)",
        "traffic_source": (dram.get("source", "") + " via profiles/dram_traffic.json (ncu --set full, dram read + write "
                           "per launch, cold cache)") if dram else None,
        "peak_source": {"hbm": peaks["source"],
                        "l2": f"own L2 read microbenchmark in this run (mdeggs_measure_l2_peak, 64 MiB buffer): {l2_peak:.0f} GB/s"},
        "kernels": kernels, "stage_ms": stage,
        "timing_note": "kernel times: CUDA events after every launch of eagerly launched steps (engine option trace), "
                       "median; the headline `value` is the graph-replayed step",
        "whole_step": {"algorithmic_bytes": pipe_bytes, "achieved": pipe_bytes / step_s / 1e9,
                       "frac": pipe_bytes / step_s / 1e9 / peaks["hbm_gbs"], "unit": "GB/s",
                       "ncu_dram_bytes_per_step": dram.get("step_total_bytes"),
                       "model": "SURVEY 8d whole-pipeline bytes / measured step time, against the HBM peak; most of "
                                "these bytes are L2 hits by construction (see ncu_dram_bytes_per_step)"}}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": _config(name, prob, int(st["n_nodes"]), world, {
            "parallelism": "1 GPU" if world == 1 else f"{world} independent omic layers, one per GPU (no collective)",
            "l2": "inputs (160 MB) exceed L2; plus a 256 MiB read-modify-write flush between steps (untimed)",
            "timing": "CUDA events per step on the default stream, max over ranks",
            "launch_mode": "compute phase replayed from a CUDA graph" if graph_replayed else "eager launches",
            "host_cpus": cpu_binding,
            "note": "front = one TMA-fed kernel (fit_path bit 10)" if int(st["fit_path"]) & 0x400 else None}),
        "ref_equiv_tests_per_s": world * ref_equiv * args.steps / (ms * 1e-3),
        "result_check": {"best_percentile_index": best, "sig_pairs": int(sig[best - 1]) if best > 0 else 0,
                         "tests_the_reference_would_run": ref_equiv},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s / args.steps, "stage_ms_last_step": e2e_stage,
                "host_input_bytes": int(prob.assay.nbytes + 8 * E),
                "api": "mdeggs_run_omic (C ABI) with pinned host buffers; only the band of network-node rows of the "
                       "assay crosses PCIe"},
        "e2e_shuffled_rows": e2e_shuf,
        "e2e_pageable": e2e_page,
        "gpu_launches": int(st["kernel_launches"]) * args.steps,
        "clocks": clocks,
        "roofline": roofline,
    }

    # ---- CPU baseline beside it (rank 0, N = 1): the oracle on the box's host cores ----
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        cores = pyoracle.lib().mdeggs_oracle_threads()
        t0 = time.perf_counter()
        pyoracle.run_omic(prob, ("tot_edges", "sig_count", "best"), pyoracle.MODE_REFERENCE, 0)
        dt_ref = time.perf_counter() - t0
        t0 = time.perf_counter()
        r_u = pyoracle.run_omic(prob, ("p_edge", "tot_edges", "sig_count", "best"), pyoracle.MODE_UNIQUE, 0)
        dt_uni = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": E / dt_ref, "unit": UNIT, "cores": min(cores, P), "kind": "port",
                                "sample": f"one full {name} job, reference mode (refit per percentile, "
                                          f"percentile-parallel like mclapply), {dt_ref:.2f} s; host threads {cores}",
                                "unique_mode": {"value": E / dt_uni, "cores": cores, "seconds": dt_uni}}
        line["result_check"]["oracle_best"] = int(r_u["best"][0])
        line["result_check"]["oracle_sig_pairs"] = int(r_u["sig_count"][int(r_u["best"][0]) - 1]) if r_u["best"][0] > 0 else 0

    # ---- also: several omic layers of one get_diffNetworks call in flight at once (mdeggs_run_omics) ----
    # Informational: the main line above runs one layer per step and waits for it (latency included); a multi-omic
    # call (core_functions.R:214-234) submits its layers on one context each before waiting, so the host latency and
    # the launch gaps of one layer hide behind the kernels of the others.
    if rank == 0 and world == 1 and not args.no_batch:
        import ctypes as C
        from multideggs_b200.engine import EnginePool
        n_layers = 3
        pool = EnginePool(local)
        engs = pool.engines(n_layers)
        probs = [_device_problem(torch, prob, dev) for _ in range(n_layers)]
        ctxs = (C.c_void_p * n_layers)(*[e._ctx for e in engs])
        cprobs = (_abi.CProblem * n_layers)(*[p[0] for p in probs])
        cress = (_abi.CResult * n_layers)(*[p[1] for p in probs])
        rcs = (C.c_int * n_layers)()
        lib = eng._lib
        for _ in range(args.warmup):
            assert lib.mdeggs_run_omics(ctxs, n_layers, cprobs, cress, rcs) == 0
        torch.cuda.synchronize()
        evb = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.add_(1.0)
            evb[i][0].record()
            lib.mdeggs_run_omics(ctxs, n_layers, cprobs, cress, rcs)
            evb[i][1].record()
        torch.cuda.synchronize()
        msb = sum(a.elapsed_time(b) for a, b in evb)
        for p_ in probs:
            assert int(p_[2][5]["best"].cpu().numpy()[0]) == best, "batched layers disagree with the single layer"
        line.setdefault("also", {})["cfg3_layers_in_flight"] = {
            "workload": name, "layers_per_step": n_layers, "api": "mdeggs_run_omics, one context per layer",
            "ms_per_step": msb / args.steps, "ms_per_layer": msb / args.steps / n_layers,
            "value": n_layers * E * args.steps / (msb * 1e-3), "unit": UNIT}
        del probs, cprobs, cress
        pool.close()
        torch.cuda.empty_cache()

    # ---- also: BASELINE configs 1, 2 and 4 through the public API ----
    if rank == 0 and world == 1 and not args.no_small:
        line.setdefault("also", {}).update(_also_small_configs(md, args.steps))

    # ---- also: config 5, edges sharded over the ranks + NCCL all-gather of p-values (strong scaling) ----
    if not args.no_cfg5:
        del cprob, cres, keep, hprob, hres, hkeep
        torch.cuda.empty_cache()
        name5 = "cfg5-small" if name.endswith("small") else "cfg5"
        wl5 = _workload(name5)
        prob5 = wl5.problem()
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                uid = torch.frombuffer(bytearray(md.Engine.nccl_unique_id()), dtype=torch.uint8).to(dev)
            dist.broadcast(uid, 0)
            eng5 = md.Engine.for_rank(local, rank, world, bytes(uid.cpu().numpy().tobytes()))
        else:
            eng5 = eng
        c5, r5, keep5 = _device_problem(torch, prob5, dev)
        steps5 = max(2, args.steps // 4)
        # timed with the compute phase (NCCL collectives included) replayed from a CUDA graph, like the main line; two
        # eager steps afterwards give the per-stage times
        eng5.set_option("graph", 1)
        ms5, _, _, _ = _timed_steps(torch, eng5, c5, r5, steps5, 3, flush, dist_sync)
        ms5 = max_over_ranks(ms5)
        eng5.set_option("graph", 0)
        _, _, stage5, st5 = _timed_steps(torch, eng5, c5, r5, 2, 1, flush, dist_sync)
        eng5.set_option("graph", 1)
        E5, n5 = prob5.dims[2], prob5.dims[1]
        line.setdefault("also", {})["cfg5"] = {
            "workload": name5, "edges": E5, "samples": n5, "features": prob5.dims[0], "scaling": "strong",
            "parallelism": f"edges sharded over {world} GPU(s)" + ("; NCCL all-gather of p-values + status" if world > 1 else ""),
            "value": E5 * steps5 / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5 / steps5, "steps": steps5,
            "stage_ms": stage5, "fit_path": int(st5["fit_path"]),
            "fit_GBps_algorithmic": (E5 / world) * (16 * n5 + 16) / (stage5["ms_fit"] * 1e-3) / 1e9 if stage5["ms_fit"] > 0 else None,
            "best": int(keep5[5]["best"].cpu().numpy()[0])}
        if world == 1:  # per-kernel times of the config-5 step and the FP64 roofline of the dense moments kernel
            kt5 = _kernel_table(torch, eng5, c5, r5, flush, 2)
            nn5 = int(st5["n_nodes"])
            dense_us = next((r["us"] for r in kt5 if r["kernel"] == "k_dense_tma"), None)
            fp64_peak = eng5.fp64_peak_tflops()
            flops5 = float(nn5) * nn5 * n5  # nn^2 / 2 pairs x n samples x 2 flops (FMA) over all categories
            line["also"]["cfg5"]["kernels_us"] = {r["kernel"]: round(r["us"], 1) for r in kt5 if r["us"]}
            if dense_us:
                line["also"]["cfg5"]["dense_roofline"] = {
                    "kernel": "k_dense_tma", "bound": "fp64", "us": dense_us, "flops": flops5,
                    "achieved": flops5 / (dense_us * 1e-6) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": flops5 / (dense_us * 1e-6) / 1e12 / fp64_peak,
                    "peak_source": "own DFMA-chain microbenchmark in this run (mdeggs_measure_fp64_peak)",
                    "flops_model": "upper-triangular 128 x 128 tile pairs: nn^2 / 2 pairs x n samples x 2 flops"}

        def bits_equal(a, b) -> bool:  # bit for bit, NaN payloads included
            if a.dtype.is_floating_point:
                a, b = a.view(torch.int64), b.view(torch.int64)
            return bool(torch.equal(a, b))

        def same_as_single(cprob_sh, keep_sh, prob_x):
            """The sharded job's outputs against the same job on this rank's GPU alone; also its single-GPU time."""
            c1, r1, keep1 = _device_problem(torch, prob_x, dev)
            eng.set_option("graph", 0)
            ms1, _, _, _ = _timed_steps(torch, eng, c1, r1, 2, 1, flush, lambda: None)
            eng.set_option("graph", 1)
            ok = all(bits_equal(keep_sh[5][k], keep1[5][k]) for k in WANT)
            t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            del c1, r1, keep1
            return bool(t.item() > 0.5), ms1 / 2

        if world > 1:
            ok5, ms5_single = same_as_single(c5, keep5, prob5)
            ms5_single = max_over_ranks(ms5_single)
            line["result_check"]["sharded_equals_single"] = {"cfg5": ok5}
            line["cfg5_strong"] = {"n_gpus": world, "ms": ms5 / steps5, "ms_1gpu_same_run": ms5_single,
                                   "speedup": ms5_single / (ms5 / steps5), "efficiency": ms5_single / (ms5 / steps5) / world}
            # the north-star sharding applied to config 3 itself: ONE job, its 53,035 edges split over the ranks, NCCL
            # all-gather of p-values + all-reduce of the counts (reported as is: the job is too small to gain from it)
            prob3 = _workload(name, seed_offset=0).problem()
            c3, r3, keep3 = _device_problem(torch, prob3, dev)
            ms3, _, stage3, _ = _timed_steps(torch, eng5, c3, r3, args.steps, args.warmup, flush, dist_sync)
            ms3 = max_over_ranks(ms3)
            line["also"]["cfg3_edge_sharded"] = {
                "workload": name, "scaling": "strong", "parallelism": f"one job, edges sharded over {world} GPUs; NCCL "
                "all-gather of p-values + status, all-reduce of the significant-pair counts",
                "value": prob3.dims[2] * args.steps / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3 / args.steps,
                "stage_ms": stage3, "best": int(keep3[5]["best"].cpu().numpy()[0])}
            ok3, _ = same_as_single(c3, keep3, prob3)
            line["result_check"]["sharded_equals_single"]["cfg3"] = ok3
            del c3, r3, keep3
            eng5.close()
        else:
            line["cfg5_strong"] = {"n_gpus": 1, "ms": ms5 / steps5, "ms_1gpu_same_run": ms5 / steps5, "speedup": 1.0,
                                   "efficiency": 1.0}

    # ---- also: rlm stress (K6 batched IRLS), FP64-pipe roofline against this GPU's own DFMA peak (rank 0) ----
    if rank == 0 and not args.no_rlm:
        from multideggs_b200 import _abi as A
        rng = np.random.default_rng(7)
        grp2 = (np.arange(n) % 2 + 1).astype(np.int32)
        rng.shuffle(grp2)
        rprob = A.Problem(assay=prob.assay, from_idx=prob.from_idx, to_idx=prob.to_idx, group=grp2, K=2, pct=prob.pct,
                          method=A.METHOD_RLM, padj=A.PADJ_BH)
        eng.set_option("graph", 0)
        want_r = ("iters", "status", "best", "tot_edges")
        res_r = eng.run_omic(rprob, want_r)                       # warm-up + iteration counts
        fit_ms_r = []
        for _ in range(3):
            eng.run_omic(rprob, want_r)
            fit_ms_r.append(eng.stats()["ms_fit"])
        eng.set_option("graph", 1)
        it = res_r["iters"][res_r["status"] != A.EDGE_SELFLOOP].astype(np.float64)
        flops_r = float(np.sum(n * (18.0 + 20.0 * it)))          # SURVEY section 8 d: n (18 + 20 iters) per pair fit
        peak64 = eng.fp64_peak_tflops()
        ms_r = float(np.median(fit_ms_r))
        line.setdefault("also", {})["rlm_stress"] = {
            "workload": f"{name} matrix and network, 2 shuffled categories, regression_method='rlm'", "edges": int(E),
            "samples": int(n), "kernel": "k_fit_rlm", "kernel_ms": ms_r, "pair_fits_per_s": E / (ms_r * 1e-3),
            "mean_irls_iterations": float(it.mean()),
            "roofline": {"bound": "fp64", "achieved": flops_r / (ms_r * 1e-3) / 1e12, "peak": peak64, "unit": "TFLOP/s",
                         "frac": flops_r / (ms_r * 1e-3) / 1e12 / peak64,
                         "peak_source": "own DFMA-chain microbenchmark in this run (mdeggs_measure_fp64_peak)",
                         "flops_model": "n (18 + 20 iters) per pair fit, FMA = 2 flops; the exact MAD selection is not credited"}}

    if rank == 0:
        _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default=None, help="cfg3 (default) | cfg3-small | cfg5 | cfg5-small")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the config-5 'also' leg")
    ap.add_argument("--no-batch", action="store_true", help="skip the layers-in-flight 'also' leg")
    ap.add_argument("--no-rlm", action="store_true", help="skip the rlm stress 'also' leg")
    ap.add_argument("--no-small", action="store_true", help="skip the config 1 / 2 / 4 'also' legs")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5  # each step is a whole job of several seconds on the host cores
        args.warmup = min(args.warmup, 1)
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
