import torch, time
n = 300 * 1080 * 1920
h = torch.empty(n, dtype=torch.int32).pin_memory()
d = torch.empty(n, dtype=torch.int32, device="cuda")
for chunk in (n, n // 19, n // 300):
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        for off in range(0, n - chunk + 1, chunk):
            d[off:off + chunk].copy_(h[off:off + chunk], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"chunk {chunk*4/1e6:8.1f} MB: {n*4/dt/1e9:.1f} GB/s ({dt*1e3:.1f} ms)")
