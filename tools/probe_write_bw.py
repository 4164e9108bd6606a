"""Write-only HBM bandwidth on this GPU (development probe): what a pure store stream of 4 GB achieves, to put the
materialise kernel's 4.06 TB/s beside the read+write copy figure of MEASURED_PEAKS.json."""
import json
import torch

torch.cuda.set_device(0)
n = 4_032_000_000 // 8
x = torch.empty(n, dtype=torch.float64, device="cuda")
res = {}
for name, fn in (("fill_", lambda: x.fill_(1.5)), ("zero_ (memset)", lambda: x.zero_())):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    res[name] = {"ms": ms, "gbs": n * 8 / ms / 1e6}
y = torch.empty(n // 2, dtype=torch.float64, device="cuda")
z = torch.empty(n // 2, dtype=torch.float64, device="cuda")
for _ in range(3):
    z.copy_(y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    z.copy_(y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
res["copy (read + write)"] = {"ms": ms, "gbs": 2 * (n // 2) * 8 / ms / 1e6}
print(json.dumps({"write_only_probe_4GB": res}))
