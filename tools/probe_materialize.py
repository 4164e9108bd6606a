"""Achieved HBM write bandwidth of kernel (1) alone (W materialised in the reference layout, regressor.py:46-87):
UR10, N samples, W = (6 N) x 84 doubles = 4 032 B per sample."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "figaroh-plus_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    import torch
    from conftest import load_flat, make_param
    from figaroh_b200 import engine as eng
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    fm = load_flat("ur10")
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(make_param(True, True, True))
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    q, v, a = (torch.rand((N, 6), generator=g, device="cuda", dtype=torch.float64) * 12 - 6 for _ in range(3))
    rps, cols = dm.shape(cfg)
    W = torch.empty((rps * N, cols), dtype=torch.float64, device="cuda")
    for _ in range(3):
        dm.materialize(cfg, q, v, a, out=W)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        dm.materialize(cfg, q, v, a, out=W)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = W.numel() * 8
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6557.8
    print(json.dumps({"kernel": "regressor_materialize_kernel", "N": N, "bytes_written": nbytes, "ms": ms,
                      "achieved_gbs": nbytes / ms / 1e6, "peak_gbs": peak, "frac": nbytes / ms / 1e6 / peak}))


if __name__ == "__main__":
    main()
