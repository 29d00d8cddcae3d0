#!/usr/bin/env python
"""End-to-end (host buffers) step time for several pipeline chunk sizes (development aid)."""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import wga_pkg  # noqa: E402

W = wga_pkg.load()
W.lib()
torch.cuda.set_device(0)
base, meta = bench.prepare_workload(W, sys.argv[1], 0, 1, None)
g = W.ANSBvGraph.load(base)
n, arcs = g.num_nodes(), g.num_arcs_hint()
h_off = torch.empty(n + 1, dtype=torch.int64).pin_memory()
h_succ = torch.empty(arcs + 1024, dtype=torch.int32).pin_memory()
got = C.c_uint64(0)
for chunk in [int(x) for x in sys.argv[2:]]:
    W.set_tuning(e2e_chunk=chunk)

    def step():
        assert W.lib().wga_upload(g._h, C.c_void_p(0)) == 0
        rc = W.lib().wga_decode_range_host(g._h, C.c_uint64(0), C.c_uint64(n), C.c_void_p(h_off.data_ptr()),
                                           C.c_void_p(h_succ.data_ptr()), C.c_uint64(h_succ.numel()), C.byref(got))
        assert rc == 0, W.lib().wga_last_error()
    step(); step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("chunk %8d: %.2f ms  %.2f Garcs/s" % (chunk, dt * 1e3, arcs / dt / 1e9), flush=True)
