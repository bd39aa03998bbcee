"""Write-only and copy bandwidth of the HBM (torch kernels): what a 2.2 GB store stream costs at best.
usage: python tools/micro/write_bw.py"""
import torch

n = 2_204_000_000 // 8
x = torch.empty(n, dtype=torch.float64, device="cuda")
y = torch.empty(n, dtype=torch.float64, device="cuda")


def timed(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


gb = n * 8 / 1e9
for name, f, traffic in (("memset (x.zero_())", lambda: x.zero_(), gb), ("fill kernel (x.fill_(1.5))", lambda: x.fill_(1.5), gb),
                         ("copy (y.copy_(x))", lambda: y.copy_(x), 2 * gb), ("read-only (x.sum())", lambda: x.sum(), gb)):
    ms = timed(f)
    print(f"{name:30s} {ms:7.3f} ms  {traffic / ms:7.1f} GB/s of DRAM traffic ({gb:.2f} GB array)")
