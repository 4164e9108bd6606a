#!/bin/bash
# Runs the FP64 probes on the GPU box; results land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/fp64_probe_smi.csv
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/fp64_probe_clocks.csv &
SMI=$!
./tools/fp64_probe > gpurun_out/fp64_probe.json
python - <<'PY' > gpurun_out/fp64_dgemm.json
import torch, json, time
torch.backends.cuda.matmul.allow_tf32 = False
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2): (a @ b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
# sustained 3 s
t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    c = a @ b; k += 1
    if k % 4 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / k
print(json.dumps({"dgemm_8192_burst_tflops": 2 * n**3 / (best * 1e-3) / 1e12,
                  "dgemm_8192_sustained_tflops": 2 * n**3 / (sus * 1e-3) / 1e12, "n_sustained": k}))
PY
kill $SMI
cat gpurun_out/fp64_probe.json gpurun_out/fp64_dgemm.json
