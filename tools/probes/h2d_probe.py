"""PCIe host->device rate for the e2e path: 2 x 1.07 GB pinned buffers, one stream vs two streams vs 8 chunks."""
import torch, time
dev = torch.device("cuda", 0)
n = 16 * 4096 * 4096
a, b = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
da, db = torch.empty(n, device=dev), torch.empty(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def one():
    da.copy_(a, non_blocking=True); db.copy_(b, non_blocking=True)
def two():
    with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
    with torch.cuda.stream(s2): db.copy_(b, non_blocking=True)
def chunks():
    c = n // 4
    for i in range(4):
        with torch.cuda.stream(s1 if i % 2 == 0 else s2):
            da[i*c:(i+1)*c].copy_(a[i*c:(i+1)*c], non_blocking=True); db[i*c:(i+1)*c].copy_(b[i*c:(i+1)*c], non_blocking=True)
for name, fn in (("one stream", one), ("two streams", two), ("4 chunks x 2 streams", chunks)):
    for _ in range(2): fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name:24s} {dt*1e3:7.2f} ms  {2*n*4/dt/1e9:6.1f} GB/s")
