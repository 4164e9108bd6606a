"""Host-to-device copy bandwidth from pinned memory at the chunk sizes of the host-buffer path (development probe)."""
import json
import torch

torch.cuda.set_device(0)
res = {}
for mb in (1, 6, 12, 25, 50, 100, 192):
    n = mb * 1024 * 1024 // 8
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(4, 400 // mb)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    res[f"{mb}MB"] = round(n * 8 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
print(json.dumps({"h2d_pinned_gbs": res}))
