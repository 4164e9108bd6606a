#!/usr/bin/env python
"""BASELINE.json configs[4]: UR10 sample-count sweep 1e4 ... 1e8 beside the host path.

Run under torchrun for N GPUs (every rank calls it; it launches nothing itself - it loops bench.main in-process
is not possible because of the process group, so it shells out to bench.py with the same launcher environment).
Single-GPU use:  python tools/sweep.py --out gpurun_out/sweep_n1.jsonl
Multi-GPU use:   python tools/sweep.py --gpus 8 --out gpurun_out/sweep_n8.jsonl     (spawns torchrun per point)
Per point the TOTAL sample count is fixed (strong scaling across GPU counts, as the config says "vs host path").
The CPU path is measured where it is feasible: the reference holds W and W_mod (2 x 4 032 B per UR10 sample,
regressor.py:46-47), i.e. 81 GB at 1e7 samples - beyond ~5e6 samples it does not fit the 62 GB survey host and is
reported as "reference infeasible (memory)"; below that the port is timed on min(N, 2e5) samples (the per-sample
loop is linear in N, the two LAPACK QRs are linear in rows) and the rate is quoted with that sample stated.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--sizes", default="1e4,1e5,1e6,1e7,1e8")
    ap.add_argument("--cpu", action="store_true", help="also time the CPU port at every feasible size (1 GPU run only)")
    args = ap.parse_args()
    sizes = [int(float(s)) for s in args.sizes.split(",")]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as out:
        for n in sizes:
            steps = 20 if n <= 1e6 else (5 if n <= 1e7 else 2)
            cmd = [sys.executable]
            if args.gpus > 1:
                cmd += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr",
                        "127.0.0.1", "--master-port", str(29600 + (n % 97))]
            cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(args.gpus), "--total-samples", str(n), "--steps", str(steps),
                    "--warmup", "3", "--no-parity-check"]
            if n > 2e7:
                cmd.append("--no-e2e")  # 19 GB of pinned host memory per 1e8 samples: the device-resident number only
            if not args.cpu or n > 2e5:
                cmd.append("--no-cpu-baseline")
            else:
                cmd += ["--cpu-baseline-samples", str(n)]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=3000)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
            try:
                d = json.loads(line)
            except Exception:  # noqa: BLE001
                d = {"error": (r.stderr or "")[-500:], "total_samples": n}
            d["total_samples"] = n
            if n >= 5e6:
                d["cpu_path"] = "reference infeasible (memory): W and W_mod need %.0f GB" % (2 * 4032 * n / 1e9)
            out.write(json.dumps(d) + "\n")
            out.flush()
            print(n, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"),
                  d.get("error", "")[:200], flush=True)


if __name__ == "__main__":
    main()
