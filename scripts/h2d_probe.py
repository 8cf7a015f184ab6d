import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.float64, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("H2D pinned 8 GiB: %.1f GB/s" % (8 * n / dt / 1e9))
    torch.cuda.synchronize(); t = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("D2H pinned 8 GiB: %.1f GB/s" % (8 * n / dt / 1e9))
