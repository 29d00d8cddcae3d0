#!/usr/bin/env python
"""Times one BvComp pass of a bench workload with the candidate costing on the host threads and on the GPU.
usage: python tools/time_bvcomp.py [workload]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import wga_pkg  # noqa: E402

W = wga_pkg.load()
wl = sys.argv[1] if len(sys.argv) > 1 else "eu-2015-host-shaped"
kind, n, deg, seed = bench.WORKLOADS[wl]
thr = os.cpu_count() or 1
off, succ = W.synth_graph(kind, n, deg, seed=seed, threads=thr)
print(wl, n, "nodes", succ.size, "arcs", thr, "host threads", flush=True)
tables = None
for name in ("Log2Estimator", "EntropyEstimator"):
    res = {}
    for gpu in (False, True, True):
        t0 = time.time()
        c, s = W.bvcomp_symbols(off, succ, estimator_tables=tables, chunk_nodes=bench.CHUNK_NODES, threads=thr,
                                gpu_costing=gpu, **bench.BVCOMP)
        res[gpu] = (time.time() - t0, c, s)
    assert (res[False][1] == res[True][1]).all() and (res[False][2] == res[True][2]).all()
    print("%s pass: host costing %.2f s, GPU costing %.2f s (%.1fx), %d symbols, identical" %
          (name, res[False][0], res[True][0], res[False][0] / res[True][0], res[True][2].size), flush=True)
    if tables is None:
        mb = W.ANSModel4EncoderBuilder()
        mb.push_symbols(res[True][1], res[True][2])
        tables = mb.build()[0]
