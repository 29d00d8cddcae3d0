import torch, time
x=torch.empty(410_000_000, dtype=torch.int32, device="cuda")
h=torch.empty(410_000_000, dtype=torch.int32).pin_memory()
for n in (410_000_000, 37_000_000):
    for _ in range(2): h[:n].copy_(x[:n], non_blocking=True); torch.cuda.synchronize()
    t=time.perf_counter(); h[:n].copy_(x[:n], non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("D2H", n*4/1e9, "GB", dt*1e3, "ms", n*4/dt/1e9, "GB/s")
    t=time.perf_counter(); x[:n].copy_(h[:n], non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("H2D", n*4/1e9, "GB", dt*1e3, "ms", n*4/dt/1e9, "GB/s")
