import torch, time
x = torch.empty(2 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(x, non_blocking=True)), ("d2h", lambda: x.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    print(name, round(3 * x.numel() / (time.time() - t0) / 1e9, 1), "GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
y = torch.empty(2 << 30, dtype=torch.uint8).pin_memory(); e = torch.empty_like(y, device="cuda")
torch.cuda.synchronize(); t0 = time.time()
for _ in range(3):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(e, non_blocking=True)
torch.cuda.synchronize()
print("both directions, each", round(3 * x.numel() / (time.time() - t0) / 1e9, 1), "GB/s")
