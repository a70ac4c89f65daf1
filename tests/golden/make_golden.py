"""Regenerates the committed fixtures from the reference's bundled data (run in the BUILD container only;
/root/reference does not exist on the GPU box).

  tests/golden/bundled.npz                      synthetic_{rnaseq,proteomic,Olink}Data + synthetic_metadata
                                                (reference data/*.rda; documented in R/example_data.R:1-40)
  multideggs_b200/data/omic_network.tsv.gz      the internal `omic_network` (reference R/sysdata.rda,
                                                built by data-raw/omic_network.R:1-29): 53,035 directed edges

Usage: python tests/golden/make_golden.py [/root/reference]
"""
import gzip
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from multideggs_b200 import rdata  # noqa: E402


def main(ref: str) -> None:
    out = {}
    for key, f in (("rnaseq", "synthetic_rnaseqData"), ("proteomic", "synthetic_proteomicData"),
                   ("olink", "synthetic_OlinkData")):
        m, rn, cn = rdata.as_matrix(rdata.read_rda(f"{ref}/data/{f}.rda")[f])
        out[f"{key}_x"] = m
        out[f"{key}_genes"] = np.array(rn)
        out[f"{key}_samples"] = np.array(cn)
    md = rdata.as_dataframe(rdata.read_rda(f"{ref}/data/synthetic_metadata.rda")["synthetic_metadata"])
    out["meta_samples"] = np.array(md["_row_names"])
    out["meta_response"] = np.array(md["response"])
    out["meta_response_levels"] = np.array(md["response._levels"])
    out["meta_patient_id"] = np.array(md["patient_id"])
    out["meta_age"] = np.asarray(md["age"], dtype=np.float64)
    out["meta_gender"] = np.array(md["gender"])
    np.savez_compressed(os.path.join(ROOT, "tests/golden/bundled.npz"), **out)

    net = rdata.as_dataframe(rdata.read_rda(f"{ref}/R/sysdata.rda")["omic_network"])
    buf = io.StringIO()
    buf.write("from\tto\n")
    for a, b in zip(net["from"], net["to"]):
        buf.write(f"{a}\t{b}\n")
    path = os.path.join(ROOT, "multideggs_b200/data/omic_network.tsv.gz")
    with open(path, "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as gz:
        gz.write(buf.getvalue().encode())
    print("edges", len(net["from"]), "nodes", len(set(net["from"]) | set(net["to"])))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
