import torch, time
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")   # 4 GiB
y = torch.empty_like(x)
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.zero_());  print("write-only  (zero_)  GB/s", x.numel() * 4 / ms / 1e6)
ms = t(lambda: x.fill_(1.5)); print("write-only  (fill_)  GB/s", x.numel() * 4 / ms / 1e6)
ms = t(lambda: y.copy_(x));  print("copy r+w             GB/s", 2 * x.numel() * 4 / ms / 1e6)
ms = t(lambda: x.sum());     print("read-only   (sum)    GB/s", x.numel() * 4 / ms / 1e6)
ms = t(lambda: torch.add(x, 1.0, out=y)); print("add r+w              GB/s", 2 * x.numel() * 4 / ms / 1e6)
