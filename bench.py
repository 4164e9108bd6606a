#!/usr/bin/env python
"""bench.py - samples/s of the identification hot path (regressor build + TSQR + OLS solve).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--samples S] [--impl b200|reference]

Workload (BASELINE.json configs[1]): UR10, 6 DoF, friction + actuator inertia + joint
offsets, S = 1,000,000 synthetic samples per GPU.  One step = one full pass
(q, dq, ddq, tau) -> eliminated columns, base-parameter index set, beta, phi_b.
With N > 1 (torchrun) every rank folds its own S samples (weak scaling) and the step
includes the all-gather + merge of the per-GPU R factors.
Prints ONE JSON line (see the task contract for the keys).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "figaroh-plus_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

MODEL_JSON = os.path.join(ROOT, "tests", "golden", "models", "ur10.json")
PARAM = dict(is_joint_torques=True, is_external_wrench=False, force_torque=None,
             has_friction=True, has_actuator_inertia=True, has_joint_offset=True,
             # UR10's yaml has None here; SURVEY 8(d): take the TX40 values
             fv=[8.05e0, 5.53e0, 1.97e0, 1.11e0, 1.86e0, 6.5e-1], fs=[7.14e0, 8.26e0, 6.34e0, 2.48e0, 3.03e0, 2.82e-1],
             Ia=[3.62e-1, 3.62e-1, 9.88e-2, 3.13e-2, 4.68e-2, 1.05e-2],
             off=[3.92e-1, 1.37, 3.26e-1, -1.02e-1, -2.88e-2, 1.27e-1])
METRIC = "samples/sec for regressor build+TSQR+OLS solve"
# algorithmic flops per sample: F_reg + 2*m_s*(p_e+1)^2, UR10 flags on: p_e = 73
F_REG = 170 * 6 + 120 * 6 + 420 * 21
F_QR = 2 * 6 * (73 + 1) ** 2
FLOPS_PER_SAMPLE = F_REG + F_QR


def fp64_peak_tflops():
    """FP64 denominator measured by tools/fp64_probe.cu on this pool (profiles/fp64_probe_r01.json)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "fp64_probe_r01.json")))
        return max(v for k, v in d.items() if k.startswith("dmma_") and k.endswith("_tflops")), "measured (DMMA m8n8k4 probe, profiles/fp64_probe_r01.json)"
    except Exception:  # noqa: BLE001
        return 37.0, "nominal"


_REAL_STDOUT = sys.stdout


def ncu_traffic(n):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed ncu
    capture of this command at the same sample count (profiles/ncu_traffic_r01.json); None if there is none."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")))
        return d["dram_bytes_per_launch"] if int(d["samples"]) == int(n) else None
    except Exception:  # noqa: BLE001
        return None


def standard_phi(fm):
    """phi_std in FIGAROH column order from the flattened model (no oracle involved)."""
    pin2fig = [9, 6, 7, 8, 0, 1, 3, 2, 4, 5]
    dyn = fm.dynamic_parameters()
    phi, names = [], []
    inertial = ("Ixx", "Ixy", "Ixz", "Iyy", "Iyz", "Izz", "mx", "my", "mz", "m")
    for i in range(1, fm.njoints):
        blk = np.empty(10)
        blk[pin2fig] = dyn[i - 1]
        phi += list(blk) + [PARAM["Ia"][i - 1], PARAM["fv"][i - 1], PARAM["fs"][i - 1], PARAM["off"][i - 1]]
        names += [n + str(i) for n in inertial] + [f"Ia{i}", f"fv{i}", f"fs{i}", f"off{i}"]
    return np.array(phi), names


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (every ~2 ms, the
    timed region of the default run lasts ~50 ms), nvidia-smi -lms as the fallback when pynvml is missing."""

    _REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.reason_bits, self.mx = [], 0, None
        self.path = f"/tmp/fg_bench_clocks_{os.getpid()}.csv"

    def _poll(self):
        import pynvml as nv
        h = self.handle
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            self.handle = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.thread = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out = {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.mx,
                       "reasons": sorted(n for n, bit in self._REASONS if self.reason_bits & bit),
                       "samples": len(self.sm), "source": "nvml"}
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                   "source": "nvidia-smi"}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def run_reference(args):
    """The reference's CPU path (oracle port: per-sample regressor loop + two tall np.linalg.qr)
    on the host cores, same workload and metric; each step is a bounded sample of it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import kernels, ref_path
    om = kernels.OracleModel(MODEL_JSON)
    std = ref_path.standard_parameters(om, PARAM)
    n = args.ref_samples
    rng = np.random.default_rng(0)
    q, v, a = (rng.uniform(-6, 6, (n, 6)) for _ in range(3))
    phi = np.array(list(std.values()), dtype=float)

    def step():
        W = ref_path.build_regressor_basic(om, q, v, a, PARAM, per_sample=True)
        tau = W @ phi
        idx_e, params_r = ref_path.get_index_eliminate(W, std, 1e-6)
        W_e = ref_path.build_regressor_reduced(W, idx_e)
        return ref_path.double_QR(tau, W_e, params_r, std)

    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "UR10 6-DoF, friction+actuator inertia+joint offsets, 1M synthetic samples per GPU",
                       "note": "CPU port of the reference path; each step processes a bounded sample"},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{n} samples per step (per-sample regressor loop through the C oracle standing in "
                                       "for pinocchio, get_index_eliminate, double_QR with two LAPACK QRs)"},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def cpu_baseline(n):
    from oracle import kernels, ref_path
    om = kernels.OracleModel(MODEL_JSON)
    std = ref_path.standard_parameters(om, PARAM)
    rng = np.random.default_rng(0)
    q, v, a = (rng.uniform(-6, 6, (n, 6)) for _ in range(3))
    phi = np.array(list(std.values()), dtype=float)
    t0 = time.perf_counter()
    W = ref_path.build_regressor_basic(om, q, v, a, PARAM, per_sample=True)
    tau = W @ phi
    idx_e, params_r = ref_path.get_index_eliminate(W, std, 1e-6)
    W_e = ref_path.build_regressor_reduced(W, idx_e)
    ref_path.double_QR(tau, W_e, params_r, std)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} of the workload's samples, once: per-sample regressor loop (C oracle standing in for "
                      f"pinocchio) + get_index_eliminate + double_QR (two LAPACK QRs), {dt:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--samples", type=int, default=1_000_000, help="samples per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-samples", type=int, default=100_000)
    ap.add_argument("--cpu-baseline-samples", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # Libraries (NCCL prints its version banner) write to fd 1: keep stdout for the ONE JSON line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from figaroh_b200 import dist as fdist
    from figaroh_b200 import engine as eng
    from figaroh_b200.model import FlatModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    fm = FlatModel.from_json(open(MODEL_JSON).read())
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(PARAM)
    rps, cols = dm.shape(cfg)
    phi_std, names = standard_phi(fm)
    N = args.samples

    # ---- synthetic inputs, generated on the device (seeded per rank); tau = W phi_std + 1% noise
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    q = (torch.rand((N, fm.nq), generator=g, device=dev, dtype=torch.float64) * 12 - 6)
    v = (torch.rand((N, fm.nv), generator=g, device=dev, dtype=torch.float64) * 12 - 6)
    a = (torch.rand((N, fm.nv), generator=g, device=dev, dtype=torch.float64) * 12 - 6)
    tau = torch.empty((rps, N), device=dev, dtype=torch.float64)
    phi_d = torch.from_numpy(phi_std).to(dev)
    chunk = 1 << 17
    for i0 in range(0, N, chunk):
        n = min(chunk, N - i0)
        Wc = dm.materialize(cfg, q[i0:i0 + n], v[i0:i0 + n], a[i0:i0 + n])
        tau[:, i0:i0 + n] = (Wc @ phi_d).view(rps, n)
        del Wc
    tau += 0.01 * tau.pow(2).mean().sqrt() * torch.randn(tau.shape, generator=g, device=dev, dtype=torch.float64)
    torch.cuda.synchronize()

    # N-independent symbolic pass: columns that are identically zero for this model are never generated
    sel = eng.structural_columns(dm, cfg)
    cfg_run = cfg.with_columns(sel)
    local = eng.TsqrState(len(sel))
    merged = eng.TsqrState(len(sel)) if world > 1 else None
    local.set_timing(True)

    def step_device():
        local.reset()
        local.accumulate_samples(dm, cfg_run, q, v, a, tau)
        st = fdist.allgather_merge(local, merged)
        return eng.solve_state(st, names, tol_e=1e-6, col_map=sel, p_full=cols)

    # pinned host copies for the end-to-end path
    qh, vh, ah, th = (x.cpu().pin_memory() for x in (q, v, a, tau))

    def step_e2e():
        local.reset()
        local.accumulate_samples_host(dm, cfg_run, qh, vh, ah, th)
        st = fdist.allgather_merge(local, merged)
        return eng.solve_state(st, names, tol_e=1e-6, col_map=sel, p_full=cols)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            sol = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, sol

    for _ in range(max(args.warmup, 3)):
        step_device()
    local.set_timing(True)  # reset the kernel timers after warm-up
    l0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, sol = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = (eng.launch_count() - l0) // args.steps
    kern_ms, kern_n = local.get_timing()
    kern_ms_avg = kern_ms / max(kern_n, 1)

    for _ in range(2):
        step_e2e()
    ms_e2e, sol_e = timed(step_e2e, max(3, args.steps // 2))

    if rank == 0:
        total = N * world
        peak, peak_src = fp64_peak_tflops()
        achieved = FLOPS_PER_SAMPLE * N / (kern_ms_avg * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": total / (ms_step * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "UR10 6-DoF, friction+actuator inertia+joint offsets, %s synthetic samples per GPU "
                                   "(BASELINE.json configs[1]%s)" % ("1M" if N == 1_000_000 else str(N),
                                                                    "" if N == 1_000_000 else "; sample count overridden: configs[4] sweep"),
                       "samples_per_gpu": N, "rows_per_sample": rps, "columns": cols, "columns_folded": int(len(sel)), "parallelism": f"dp{world}",
                       "l2": "inputs (192 B/sample x N = %.0f MB per step) exceed the 126 MB L2" % (N * 192 / 1e6),
                       "eliminated": len(sol.idx_e), "base_parameters": len(sol.idx_base)},
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(N * (fm.nq + 2 * fm.nv + rps) * 8),
                    "d2h_bytes_per_step": int(8 * (2 * (cols + 1) + len(sol.idx_base) * (len(sol.idx_regroup) + 1 + len(sol.idx_base))
                                                   + (len(sol.keep) + 1) ** 2))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "pipe": "fp64 (DFMA panels + DMMA m8n8k4 trailing updates; tcgen05 has no fp64 mode)",
                         "kernel": "gq_accumulate_kernel (one Householder TSQR stream per (warp, row group): sample-per-lane regressor rows into a 32 x w_g shared-memory tile, 8-column panels, DMMA trailing updates)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(N), "kernel_ms": kern_ms_avg, "kernel_share_of_step": kern_ms_avg / ms_step,
                         "flops_per_sample": FLOPS_PER_SAMPLE, "peak_source": peak_src},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.cpu_baseline_samples)
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
