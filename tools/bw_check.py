import torch, time
x = torch.rand(73_500_000, device="cuda")   # 294 MB
y = torch.empty_like(x)
def t(f, n=20):
    for _ in range(5): f()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
ms = t(lambda: x.sum()); print("sum  294 MB read: %.1f us -> %.2f TB/s" % (ms*1e3, 294e6/ms/1e9))
ms = t(lambda: y.copy_(x)); print("copy 294 MB r+w: %.1f us -> %.2f TB/s" % (ms*1e3, 2*294e6/ms/1e9))
big = torch.rand(500_000_000, device="cuda"); big2 = torch.empty_like(big)
ms = t(lambda: big.sum(), 10); print("sum  2 GB read: %.1f us -> %.2f TB/s" % (ms*1e3, 2e9/ms/1e9))
ms = t(lambda: big2.copy_(big), 10); print("copy 2 GB r+w: %.1f us -> %.2f TB/s" % (ms*1e3, 4e9/ms/1e9))
