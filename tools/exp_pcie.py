import torch, time
n = 400_000_000
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for _ in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("D2H pinned GB/s", n/dt/1e9)
for _ in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("H2D pinned GB/s", n/dt/1e9)
