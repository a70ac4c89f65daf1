import os
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def bundled():
    """The reference's bundled example data (data/*.rda), committed as tests/golden/bundled.npz."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "bundled.npz"), allow_pickle=False)
    out = {}
    for key, name in (("rnaseq", "RNAseq"), ("proteomic", "Proteomics"), ("olink", "Olink")):
        out[name] = pd.DataFrame(d[f"{key}_x"], index=d[f"{key}_genes"], columns=d[f"{key}_samples"])
    meta = pd.DataFrame({
        "patient_id": d["meta_patient_id"],
        "response": pd.Categorical(d["meta_response"], categories=list(d["meta_response_levels"])),
        "age": d["meta_age"], "gender": d["meta_gender"]}, index=d["meta_samples"])
    out["metadata"] = meta
    return out


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import multideggs_b200 as md
    return md.Engine([0])
