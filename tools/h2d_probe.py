import torch, time
x = torch.empty(135*2**20, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for sz in (2*2**20, 16*2**20, 35*2**20, 135*2**20):
    xs, ds = x[:sz], d[:sz]
    for _ in range(3): ds.copy_(xs, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = max(1, (512*2**20)//sz)
    for _ in range(n): ds.copy_(xs, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(sz/2**20, "MB", sz*n/(e0.elapsed_time(e1)*1e-3)/1e9, "GB/s")
y = torch.empty(4608, dtype=torch.uint8).pin_memory()
