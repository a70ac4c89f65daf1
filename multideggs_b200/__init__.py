"""multideggs_b200 — B200-native engine for the multiDEGGs differential-correlation hot path.

Public surface mirrors the reference R package for this path (see api.py); the compute runs in
hand-written sm_100a CUDA kernels behind the C ABI of include/mdeggs.h (csrc/).
"""
from ._abi import (METHOD_LM, METHOD_RLM, PADJ_BH, PADJ_BONFERRONI, PADJ_HOST, PADJ_NONE, RLM_COV_XTWX,
                   RLM_COV_XTX, Problem, r_seq)
from .api import (DEFAULT_PERCENTILES, FILTER_PERCENTILES, Deggs, MultiDEGGsFilter, get_diffNetworks,
                  get_diffNetworks_singleOmic, get_multiOmics_diffNetworks, get_sig_deggs, multiDEGGs_filter,
                  plot_regressions_data, predict)
from .engine import Engine, EngineError, OmicError, default_engine, load_library
from .host import NO_LINKS, Assay, Factor, IndexedNetwork, omic_network, tidy_metadata

__all__ = [n for n in dir() if not n.startswith("_")]
