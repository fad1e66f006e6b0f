"""HBM bandwidth probes on the current GPU: copy (the MEASURED_PEAKS definition), pure write (fill), pure read (sum).
   python tools/bw_probe.py"""
import torch

def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best * 1e-3

for mb in (283, 1024, 4096):
    n = mb * (1 << 20) // 4
    x = torch.empty(n, device="cuda"); y = torch.empty(n, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    def fill(): x.fill_(1.0)
    def copy(): y.copy_(x)
    def read(): x.sum()
    def fill_cold():
        flush.zero_(); 
    print("%5d MB: fill %.0f GB/s, copy (r+w) %.0f GB/s, read(sum) %.0f GB/s" % (mb, n * 4 / t(fill) / 1e9, 2 * n * 4 / t(copy) / 1e9, n * 4 / t(read) / 1e9))
