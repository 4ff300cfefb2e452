"""edsparser_b200 — B200-native MSA -> EDS / l-EDS construction and l-EDS merge (hot path of draessld/EDSParser).

The product is libedsparser_b200.so (hand-written sm_100a kernels behind the C ABI in
include/edsparser_b200.h) plus the C++17 host layer in edsparser_b200/host/. This Python package is the
ctypes plumbing used by tests/, bench.py and __graft_entry__.py.
"""
from .capi import (EDS_ERR_BAD_MSA, EDS_ERR_BUDGET, EDS_ERR_CUDA, EDS_ERR_HALO, EDS_ERR_INVALID_ARGUMENT,  # noqa: F401
                   EDS_ERR_OUT_OF_RANGE, EDS_ERR_RUNTIME, EDS_OK, EXPORTS, PRODUCT_SO, Buffer, Context, EdsError,
                   Comm, Group, Library, MsaStats, MsaView, load)
