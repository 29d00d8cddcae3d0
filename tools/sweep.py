#!/usr/bin/env python
"""Tuning sweep of the decode kernels on one workload (development aid, not part of the bench contract).
   python tools/sweep.py WORKLOAD 'k2_tpb=128' 'k2_tpb=128,ring_log2=14' ..."""
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
W.lib()
torch.cuda.set_device(0)
wl = sys.argv[1]
base, meta = bench.prepare_workload(W, wl, 0, 1, None)
g = W.ANSBvGraph.load(base)
n, arcs = g.num_nodes(), g.num_arcs_hint()
off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
succ = torch.empty(arcs + 1024, dtype=torch.int32, device="cuda")
ws = torch.empty(g.workspace_size(0, n), dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
ref = None
for spec in ["reset=1"] + sys.argv[2:]:
    W.set_tuning(reset=1)
    kw = dict(kv.split("=") for kv in spec.split(",") if kv)
    W.set_tuning(**{k: int(v) for k, v in kw.items()})
    for _ in range(2):
        g.decode_range_into(0, n, off, succ, ws, stream=stream)
    torch.cuda.synchronize()
    W.lib().wga_set_profiling(g._h, 1)
    stages = np.zeros(8, np.float32)
    acc = np.zeros(8)
    for _ in range(3):
        g.decode_range_into(0, n, off, succ, ws, stream=stream)
        torch.cuda.synchronize()
        W.lib().wga_last_profile(g._h, stages.ctypes.data_as(C.c_void_p))
        acc += stages
    W.lib().wga_set_profiling(g._h, 0)
    acc /= 3
    chk = int(succ[:arcs].to(torch.int64).sum().item())
    if ref is None:
        ref = chk
    print("%-50s K0 %.3f K1 %.3f LV %.3f K2 %.3f total %.3f ms  %s" % (spec, acc[0], acc[1], acc[2], acc[3], acc[:4].sum(),
                                                                    "ok" if chk == ref else "CHECKSUM MISMATCH"), flush=True)
