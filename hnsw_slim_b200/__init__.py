"""hnsw_slim_b200 — B200-native batched query engine for the HNSW-Slim search hot path.

The product is the C-ABI shared library built from csrc/ (include/hnswslim_b200.h); this
package holds its build script, the ctypes binding used by tests and bench.py, the
synthetic-corpus generator and the multi-GPU sharding helpers.
"""
from .capi import (HS_KIND_SLIM, HS_KIND_SLIMQ, HS_METRIC_IP, HS_METRIC_L2, HsError, Index,  # noqa: F401
                   bruteforce_knn, recall)

__all__ = ["Index", "HsError", "bruteforce_knn", "recall", "HS_KIND_SLIM", "HS_KIND_SLIMQ", "HS_METRIC_L2",
           "HS_METRIC_IP"]
