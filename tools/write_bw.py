"""Write-only vs copy bandwidth on this GPU (context for the covariance-assembly roofline)."""
import torch
n = 16384
x = torch.empty(n, n, dtype=torch.float64, device='cuda')
y = torch.empty(n, n, dtype=torch.float64, device='cuda')
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
b = x.numel() * 8
ms = t(lambda: x.fill_(1.5)); print('fill_      %.3f ms  %.0f GB/s written' % (ms, b / ms / 1e6))
ms = t(lambda: x.zero_()); print('zero_      %.3f ms  %.0f GB/s written' % (ms, b / ms / 1e6))
ms = t(lambda: y.copy_(x)); print('copy_      %.3f ms  %.0f GB/s read+written' % (ms, 2 * b / ms / 1e6))
ms = t(lambda: x.sum()); print('sum (read) %.3f ms  %.0f GB/s read' % (ms, b / ms / 1e6))
ms = t(lambda: torch.mul(x, 2.0, out=x)); print('scale r+w  %.3f ms  %.0f GB/s read+written' % (ms, 2 * b / ms / 1e6))
