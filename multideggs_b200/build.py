"""In-tree build of libmdeggs.so (sm_100a only). `python -m multideggs_b200.build [--verbose]`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OUT = os.path.join(CSRC, "libmdeggs.so")
SOURCES = ["mdeggs_engine.cu"]
HEADERS = ["mdeggs_kernels.cuh", "mdeggs_napair.cuh", "mdeggs_front.cuh", "mdeggs_hostpack.h", "mdeggs_math.cuh", "mdeggs_rlm.cuh", "mdeggs_dense.cuh", "mdeggs_qsel.cuh", "mdeggs_bsort.cuh", "mdeggs_qvalue_tab.h", "nccl_dl.h",
           os.path.join(ROOT, "include", "mdeggs.h")]


def nvcc_cmd(ptxas_verbose: bool = False) -> list[str]:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC,-O3", "-shared", "-ccbin", "g++", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl", "-lpthread"]
    if ptxas_verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    return cmd


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        res = subprocess.run(nvcc_cmd(verbose), capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building libmdeggs.so")
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose="--verbose" in sys.argv))
