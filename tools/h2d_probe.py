import torch, time
n = 135 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def run(k, reps=20):
    streams = [torch.cuda.Stream() for _ in range(k)]
    chunk = n // k
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return n / best / 1e9
for k in (1, 2, 4, 8):
    print(k, "streams: %.1f GB/s" % run(k))
# many small sequential copies on one stream (8 frames each)
s = torch.cuda.Stream(); chunk = 8 * 1920 * 1080
torch.cuda.synchronize(); best = 1e9
for _ in range(20):
    t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for i in range(n // chunk):
            d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
    torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
print("8-frame pieces, one stream: %.1f GB/s" % ((n // chunk) * chunk / best / 1e9))
