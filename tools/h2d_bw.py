"""Pinned host -> device copy bandwidth of this box (one cudaMemcpyAsync per 36 MB image, as the extraction pipeline issues)."""
import time, torch, json
n = 32
host = [torch.empty((3000, 4000, 3), dtype=torch.uint8).pin_memory() for _ in range(n)]
dev = [torch.empty((3000, 4000, 3), dtype=torch.uint8, device="cuda") for _ in range(4)]
s = torch.cuda.Stream()
res = {}
with torch.cuda.stream(s):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            dev[i % 4].copy_(host[i], non_blocking=True)
        s.synchronize()
        dt = time.perf_counter() - t0
        res[f"h2d_GBps_rep{rep}"] = round(n * 36e6 / dt / 1e9, 2)
big = torch.empty((256 << 20,), dtype=torch.uint8).pin_memory()
dbig = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize(); res["h2d_GBps_256MB"] = round(4 * (256 << 20) / (time.perf_counter() - t0) / 1e9, 2)
t0 = time.perf_counter()
for _ in range(4): big.copy_(dbig, non_blocking=True)
torch.cuda.synchronize(); res["d2h_GBps_256MB"] = round(4 * (256 << 20) / (time.perf_counter() - t0) / 1e9, 2)
print(json.dumps(res))
