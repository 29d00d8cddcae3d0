#!/usr/bin/env python
"""Runs a few decode steps of a bench workload (for ncu / quick timing): python tools/prof_decode.py [workload] [steps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import wga_pkg  # noqa: E402

W = wga_pkg.load()
wl = sys.argv[1] if len(sys.argv) > 1 else "web-1m"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
base, meta = bench.prepare_workload(W, wl, 0, 1)
g = W.ANSBvGraph.load(base)
n, arcs = g.num_nodes(), g.num_arcs_hint()
off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
succ = torch.empty(arcs + 1024, dtype=torch.int32, device="cuda")
ws = torch.empty(g.workspace_size(0, n), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
W.lib().wga_set_profiling(g._h, 1)
stages = np.zeros(8, np.float32)
for i in range(steps):
    g.decode_range_into(0, n, off, succ, ws, stream=st)
    torch.cuda.synchronize()
    W.lib().wga_last_profile(g._h, stages.ctypes.data_as(C.c_void_p))
    print(wl, "step", i, "stages ms", [round(float(x), 3) for x in stages[:4]], "sum", round(float(stages[:4].sum()), 3), flush=True)
