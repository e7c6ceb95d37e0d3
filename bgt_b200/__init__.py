"""bgt_b200 -- B200 (sm_100a) implementation of BGT's genotype hot path (PBWT decode + per-site AC/AN + -f filter).

The product is libbgt_b200.so (CUDA kernels behind the C ABI in include/bgt_b200.h).  This package is the thin
Python front end used by the tests and bench.py: it loads the library with ctypes and mirrors the reference's
reader flow (open .pbf -> select samples/groups/filter -> read rows).  There is no CPU implementation here:
if the library or a CUDA device is missing, calls raise.
"""
from .capi import (B200Error, Context, Pbf, Query, lib, lib_path, load_library, scan, SCAN_COUNTS, SCAN_HAP_BITS,
                   SCAN_HAP_BYTES, SCAN_DEVICE_OUT, Encoder, Sites, bgzf_inflate, synth_cohort, scan_device, collect, host_alloc, host_free, flt_eval_host, pbf_plan, load_scan, pbf_peek, scan_regions)

__all__ = ["B200Error", "Context", "Pbf", "Query", "Encoder", "Sites", "bgzf_inflate", "lib", "lib_path", "load_library", "scan", "synth_cohort", "scan_device", "collect", "host_alloc", "host_free", "flt_eval_host", "pbf_plan", "load_scan", "pbf_peek", "scan_regions",
           "SCAN_COUNTS", "SCAN_HAP_BITS", "SCAN_HAP_BYTES", "SCAN_DEVICE_OUT"]
