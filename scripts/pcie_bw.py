import torch, time
x = torch.empty(4 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(4 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"H2D pinned: {16 * (1<<30) / dt / 1e9:.1f} GB/s")
t0 = time.perf_counter()
for _ in range(4):
    x.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"D2H pinned: {16 * (1<<30) / dt / 1e9:.1f} GB/s")
