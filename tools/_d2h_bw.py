"""host-link bandwidth of the box: copy-engine D2H into pinned memory, whole buffer and 4 MiB pieces"""
import torch, time
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for label, chunk in (("1 GiB", n), ("4 MiB pieces", 4 << 20), ("1 MiB pieces", 1 << 20)):
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for o in range(0, n, chunk):
            h[o:o + chunk].copy_(d[o:o + chunk], non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
    print("D2H", label, "%.1f GB/s" % (n / e0.elapsed_time(e1) / 1e6))
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); d.copy_(h2, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("H2D 1 GiB %.1f GB/s" % (n / e0.elapsed_time(e1) / 1e6))
