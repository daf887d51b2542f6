"""Pure-write HBM bandwidth on this GPU (developer tool): the roofline of a kernel that only writes (materialised Psi2)."""
import torch
x = torch.empty(2 * 1024 ** 3, dtype=torch.float64, device="cuda")   # 16 GiB
for name, fn in (("fill_ (memset-like)", lambda: x.fill_(1.5)), ("zero_", lambda: x.zero_()), ("x.mul_(2) read+write", lambda: x.mul_(2.0))):
  best = 1e9
  for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
  mult = 2 if "read" in name else 1
  print(f"{name}: {mult * x.numel() * 8 / best / 1e6:.0f} GB/s")
y = torch.empty_like(x)
best = 1e9
for _ in range(4):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
  best = min(best, e0.elapsed_time(e1))
print(f"copy (read+write): {2 * x.numel() * 8 / best / 1e6:.0f} GB/s")
