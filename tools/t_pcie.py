import torch, time
x = torch.empty(88604672, dtype=torch.uint8).pin_memory(); d = torch.empty_like(x, device="cuda")
y = torch.empty(33555456, dtype=torch.uint8).pin_memory(); dy = torch.empty_like(y, device="cuda")
for name, f in (("h2d 88.6MB", lambda: d.copy_(x, non_blocking=True)), ("d2h 33.5MB", lambda: y.copy_(dy, non_blocking=True)), ("h2d 44.3MB", lambda: d[:44302336].copy_(x[:44302336], non_blocking=True))):
    for _ in range(3): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    print(name, "%.3f ms" % (dt * 1e3))
