"""Pure-write, pure-read and copy DRAM bandwidth on this GPU (development experiment)."""
import torch
dev = torch.device('cuda', 0)
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device=dev)       # 4 GiB
b = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, k=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
ms = t(lambda: a.fill_(1.0)); print('fill  (write only) %.1f GB/s' % (4 * n / ms / 1e6))
ms = t(lambda: a.zero_()); print('zero_ (memset)     %.1f GB/s' % (4 * n / ms / 1e6))
ms = t(lambda: a.sum()); print('sum   (read only)  %.1f GB/s' % (4 * n / ms / 1e6))
ms = t(lambda: b.copy_(a)); print('copy  (r + w)      %.1f GB/s' % (8 * n / ms / 1e6))
h = a[: n // 2]
ms = t(lambda: torch.add(h, 1.0, out=b[: n // 2])); print('add   (r + w)      %.1f GB/s' % (4 * n / ms / 1e6))
