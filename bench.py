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

MODELS = os.path.join(ROOT, "tests", "golden", "models")
METRIC = "samples/sec for regressor build+TSQR+OLS solve"
_TX40 = dict(fv=[8.05e0, 5.53e0, 1.97e0, 1.11e0, 1.86e0, 6.5e-1], fs=[7.14e0, 8.26e0, 6.34e0, 2.48e0, 3.03e0, 2.82e-1],
             Ia=[3.62e-1, 3.62e-1, 9.88e-2, 3.13e-2, 4.68e-2, 1.05e-2],
             off=[3.92e-1, 1.37, 3.26e-1, -1.02e-1, -2.88e-2, 1.27e-1])


def _cyc(vals, n):
    return [vals[i % len(vals)] for i in range(n)]


# The three robots of BASELINE.json configs[1..3].  `ranges`: (q, dq, ddq) half-widths of the uniform draws.
WORKLOADS = {
    "ur10": dict(
        model="ur10", baseline_config="configs[1]",
        title="UR10 6-DoF, friction+actuator inertia+joint offsets",
        param=dict(is_joint_torques=True, is_external_wrench=False, force_torque=None, has_friction=True,
                   has_actuator_inertia=True, has_joint_offset=True, **_TX40),  # UR10's yaml has None: SURVEY 8(d) takes TX40's
        active=None, ranges=(6.0, 6.0, 6.0), tol_e=1e-6, default_samples=1_000_000, cpu_samples=100_000),
    "tiago": dict(
        model="tiago", baseline_config="configs[2]",
        title="TIAGo torso+arm, 8 active joints of the 24-DoF tree (rows of the active joints only, "
              "examples/tiago/identification.py:157-166), friction+actuator inertia+joint offsets",
        param=dict(is_joint_torques=True, is_external_wrench=False, force_torque=None, has_friction=True,
                   has_actuator_inertia=True, has_joint_offset=True, fv=_cyc(_TX40["fv"], 24), fs=_cyc(_TX40["fs"], 24),
                   Ia=_cyc(_TX40["Ia"], 24), off=_cyc(_TX40["off"], 24)),
        active=list(range(13, 21)), ranges=(1.5, 2.0, 4.0), tol_e=1e-3, default_samples=1_000_000, cpu_samples=4_000),
    "talos": dict(
        model="talos_ff", baseline_config="configs[3]",
        title="Talos humanoid with free-flyer, external-wrench rows (force_torque=All, regressor.py:89-192)",
        param=dict(is_joint_torques=False, is_external_wrench=True, force_torque=["All"], has_friction=False,
                   has_actuator_inertia=False, has_joint_offset=False, fv=None, fs=None, Ia=None, off=None),
        active=None, ranges=(1.0, 1.0, 2.0), tol_e=1e-6, default_samples=1_000_000, cpu_samples=4_000),
}


def fp64_peak_tflops():
    """FP64 denominator measured by tools/fp64_probe.cu on this pool (profiles/fp64_probe_r01.json)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "fp64_probe_r01.json")))
        return max(v for k, v in d.items() if k.startswith("dmma_") and k.endswith("_tflops")), "measured (DMMA m8n8k4 probe, profiles/fp64_probe_r01.json)"
    except Exception:  # noqa: BLE001
        return 37.0, "nominal"


_REAL_STDOUT = sys.stdout


def ncu_traffic(workload, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed ncu
    capture of this command at the same workload and sample count (profiles/ncu_traffic_r02.json); None if none."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")))
        e = d.get(workload)
        return e["dram_bytes_per_launch"] if e and int(e["samples"]) == int(n) else None
    except Exception:  # noqa: BLE001
        return None


def standard_phi(fm, param):
    """phi_std in FIGAROH column order from the flattened model (robot.py:76-155 semantics; no oracle involved)."""
    pin2fig = [9, 6, 7, 8, 0, 1, 3, 2, 4, 5]
    dyn = fm.dynamic_parameters()
    phi, names = [], []
    inertial = ("Ixx", "Ixy", "Ixz", "Iyy", "Iyz", "Izz", "mx", "my", "mz", "m")

    def pick(flag, key, i):
        if not param[flag] or param.get(key) is None:
            return 0.0
        try:
            return float(param[key][i - 1])
        except Exception:  # noqa: BLE001
            return 0.0
    for i in range(1, fm.njoints):
        blk = np.empty(10)
        blk[pin2fig] = dyn[i - 1]
        phi += list(blk) + [pick("has_actuator_inertia", "Ia", i), pick("has_friction", "fv", i),
                            pick("has_friction", "fs", i), pick("has_joint_offset", "off", i)]
        names += [n + str(i) for n in inertial] + [f"Ia{i}", f"fv{i}", f"fs{i}", f"off{i}"]
    return np.array(phi), names


def active_rows(fm, wl):
    return None if wl["active"] is None else [int(fm.idx_v[i]) for i in wl["active"]]


def host_inputs(fm, wl, n, seed):
    """Seeded synthetic (q, dq, ddq) on the host with the workload's structure (used for the CPU arm and the
    parity subsample; the GPU arm draws on the device with the same structure)."""
    rng = np.random.default_rng(seed)
    rq, rv, ra = wl["ranges"]
    q = rng.uniform(-rq, rq, (n, fm.nq))
    v = rng.uniform(-rv, rv, (n, fm.nv))
    a = rng.uniform(-ra, ra, (n, fm.nv))
    act = None if wl["active"] is None else set(wl["active"])
    for i in range(1, fm.njoints):
        iq, iv, jt = int(fm.idx_q[i]), int(fm.idx_v[i]), int(fm.jtype[i])
        if act is not None and i not in act:  # resting joint: fixed configuration, no motion
            if jt == 2:
                q[:, iq], q[:, iq + 1] = 1.0, 0.0
            else:
                q[:, iq] = 0.1
            v[:, iv] = 0.0
            a[:, iv] = 0.0
        elif jt == 2:
            th = rng.uniform(-np.pi, np.pi, n)
            q[:, iq], q[:, iq + 1] = np.cos(th), np.sin(th)
        elif jt == 4:
            quat = rng.normal(size=(n, 4))
            q[:, iq + 3:iq + 7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    return q, v, a


def device_inputs(torch, fm, wl, n, seed, dev):
    """Same structure, drawn on the device."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    rq, rv, ra = wl["ranges"]

    def uni(cols, r):
        return torch.rand((n, cols), generator=g, device=dev, dtype=torch.float64) * (2 * r) - r
    q, v, a = uni(fm.nq, rq), uni(fm.nv, rv), uni(fm.nv, ra)
    act = None if wl["active"] is None else set(wl["active"])
    for i in range(1, fm.njoints):
        iq, iv, jt = int(fm.idx_q[i]), int(fm.idx_v[i]), int(fm.jtype[i])
        if act is not None and i not in act:
            if jt == 2:
                q[:, iq], q[:, iq + 1] = 1.0, 0.0
            else:
                q[:, iq] = 0.1
            v[:, iv] = 0.0
            a[:, iv] = 0.0
        elif jt == 2:
            th = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 2 - 1) * np.pi
            q[:, iq], q[:, iq + 1] = torch.cos(th), torch.sin(th)
        elif jt == 4:
            quat = torch.randn((n, 4), generator=g, device=dev, dtype=torch.float64)
            q[:, iq + 3:iq + 7] = quat / quat.norm(dim=1, keepdim=True)
    return q, v, a


def flops_per_sample(plan, rps, n_keep):
    """F = F_reg + 2 m_s (p_e + 1)^2 (Householder QR of an m_s-row slab against the p_e kept columns
    + the rhs), F_reg = 170 walked joints + 120 bodies + 420 (row group, body) hops."""
    nproc = len(plan["walk"]) if plan.get("ring_kernel") else plan.get("nbodies", 0)
    hops = sum(len(g["emit"]) for g in plan["groups"]) if plan.get("ring_kernel") else plan.get("hops", 0)
    f_reg = 170 * nproc + 120 * nproc + 420 * hops
    return f_reg + 2 * rps * (n_keep + 1) ** 2, f_reg


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


def oracle_path(om, wl, fm, q, v, a, per_sample):
    """The reference's script sequence on the CPU oracle (regressor.py:20-293, qrdecomposition.py:89-187)."""
    from oracle import ref_path
    param = wl["param"]
    n = len(q)
    std = ref_path.standard_parameters(om, param)
    W = ref_path.build_regressor_basic(om, q, v, a, param, per_sample=per_sample)
    rows = active_rows(fm, wl)
    if rows is not None:  # examples/tiago/identification.py:157-166
        W = np.concatenate([W[r * n:(r + 1) * n] for r in rows])
    phi = np.array(list(std.values()), dtype=float)
    rng = np.random.default_rng(1)
    tau = W @ phi
    tau = tau + 0.01 * np.sqrt(np.mean(tau ** 2)) * rng.normal(size=tau.shape)
    idx_e, params_r = ref_path.get_index_eliminate(W, std, wl["tol_e"])
    W_e = ref_path.build_regressor_reduced(W, idx_e)
    out = ref_path.double_QR(tau, W_e, params_r, std)
    return dict(std=std, tau=tau, idx_e=idx_e, params_r=params_r, W_e=W_e, out=out)


def cpu_arm(wl, n):
    """One timed pass of the CPU port on n samples of the workload; returns (seconds, description)."""
    from oracle import kernels
    mpath = os.path.join(MODELS, wl["model"] + ".json")
    om = kernels.OracleModel(mpath)
    from figaroh_b200.model import FlatModel
    fm = FlatModel.from_json(open(mpath).read())
    q, v, a = host_inputs(fm, wl, n, 0)
    t0 = time.perf_counter()
    oracle_path(om, wl, fm, q, v, a, per_sample=True)
    return time.perf_counter() - t0


def workload_string(wl, n, per_gpu=True):
    cnt = "1M" if n == 1_000_000 else ("10M" if n == 10_000_000 else str(n))
    return "%s, %s synthetic samples%s (BASELINE.json %s)" % (wl["title"], cnt, " per GPU" if per_gpu else " in total", wl["baseline_config"])


def run_reference(args):
    """The reference's CPU path (oracle port: per-sample regressor loop + get_index_eliminate + two tall LAPACK QRs)
    on the host cores, same workload and metric; each step is a bounded sample of it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:  # torchrun exports OMP_NUM_THREADS=1: give LAPACK the host cores back
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:  # noqa: BLE001
        pass
    wl = WORKLOADS[args.config]
    n = args.ref_samples or wl["cpu_samples"]
    for _ in range(min(args.warmup, 1)):
        cpu_arm(wl, n)
    dts = [cpu_arm(wl, n) for _ in range(args.steps)]
    dt = sum(dts) / len(dts)
    val = n / dt
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(wl, args.samples or wl["default_samples"]),
                       "note": "CPU port of the reference path; each step processes a bounded sample"},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{n} samples per step (per-sample regressor loop through the C oracle standing in "
                                       "for pinocchio, get_index_eliminate, double_QR with two LAPACK QRs)"},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def cpu_baseline(wl, n):
    dt = cpu_arm(wl, n)
    return {"value": n / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} of the workload's samples, once: per-sample regressor loop (C oracle standing in for "
                      f"pinocchio) + get_index_eliminate + double_QR (two LAPACK QRs), {dt:.1f} s"}


def parity_check(eng, torch, dm, fm, wl, cfg, sel, names, cols, n=4096):
    """Rank 0: the fused path and the CPU oracle on the SAME 4096-sample subsample: index sets, expression strings,
    base parameters in the norm in which they are well posed (the predicted torques), cond(R1) beside it."""
    from oracle import kernels
    om = kernels.OracleModel(os.path.join(MODELS, wl["model"] + ".json"))
    q, v, a = host_inputs(fm, wl, n, 12345)
    o = oracle_path(om, wl, fm, q, v, a, per_sample=False)
    rps = cfg_rows(dm, cfg)
    st = eng.TsqrState(len(sel))
    st.accumulate_samples(dm, cfg.with_columns(sel), q, v, a, o["tau"].reshape(rps, n))
    sol = eng.solve_state(st, names, tol_e=wl["tol_e"], col_map=sel, p_full=cols, params_std=o["std"])
    Wb = o["W_e"][:, sol.idx_base] if len(sol.keep) == o["W_e"].shape[1] else None
    out = {"subsample": n, "idx_e_equal": sol.idx_e == o["idx_e"],
           "idx_base_equal": [sol.params_r[i] for i in sol.idx_base] == [b.split(" ")[0] for b in o["out"][2]],
           "expressions_equal": sol.params_base == o["out"][2]}
    if Wb is not None and out["idx_base_equal"]:
        Q1, R1 = np.linalg.qr(Wb)
        phi_ref = np.linalg.solve(R1, Q1.T @ o["tau"])
        out["torque_norm_err"] = float(np.linalg.norm(R1 @ (sol.phi_b_raw - phi_ref)) / np.linalg.norm(o["tau"]))
        out["phi_b_max_rel_err"] = float(np.abs(sol.phi_b_raw - phi_ref).max() / np.abs(phi_ref).max())
        out["cond_R1_subsample"] = float(np.linalg.cond(R1))
    return out


def cfg_rows(dm, cfg):
    return dm.shape(cfg)[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="ur10", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="samples per GPU (weak scaling; default: the workload's)")
    ap.add_argument("--total-samples", type=int, default=0, help="strong scaling: this many samples sharded over the GPUs")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-samples", type=int, default=0)
    ap.add_argument("--cpu-baseline-samples", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--pageable", action="store_true", help="e2e from ordinary (pageable) host arrays instead of pinned ones")
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

    wl = WORKLOADS[args.config]
    fm = FlatModel.from_json(open(os.path.join(MODELS, wl["model"] + ".json")).read())
    dm = eng.DeviceModel(fm)
    rows = active_rows(fm, wl)
    cfg = eng.RegressorConfig.from_param(wl["param"], row_sel=rows)
    rps, cols = dm.shape(cfg)
    phi_std, names = standard_phi(fm, wl["param"])
    strong = args.total_samples > 0
    if strong:
        lo, hi = fdist.shard_range(args.total_samples, rank, world)
        N = hi - lo
    else:
        N = args.samples or wl["default_samples"]

    # ---- synthetic inputs, generated on the device (seeded per rank); tau = W phi_std + 1% noise
    q, v, a = device_inputs(torch, fm, wl, N, 1234 + rank, dev)
    tau = torch.empty((rps, N), device=dev, dtype=torch.float64)
    phi_d = torch.from_numpy(phi_std).to(dev)
    full_rps = dm.shape(eng.RegressorConfig(cfg.mode, cfg.flags, cfg.ft_mask))[0]
    chunk = max(1024, min(1 << 17, (1 << 31) // (8 * full_rps * cols)))
    rsel = None if rows is None else torch.as_tensor(rows, device=dev)
    for i0 in range(0, N, chunk):
        n = min(chunk, N - i0)
        Wc = dm.materialize(cfg, q[i0:i0 + n], v[i0:i0 + n], a[i0:i0 + n])
        t = (Wc @ phi_d).view(full_rps, n)
        tau[:, i0:i0 + n] = t if rsel is None else t[rsel]
        del Wc
    g = torch.Generator(device=dev)
    g.manual_seed(99 + rank)
    tau += 0.01 * tau.pow(2).mean().sqrt() * torch.randn(tau.shape, generator=g, device=dev, dtype=torch.float64)
    torch.cuda.synchronize()

    # N-independent symbolic pass: columns that are identically zero for this model are never generated
    sel = eng.structural_columns(dm, cfg)
    cfg_run = cfg.with_columns(sel)
    local = eng.TsqrState(len(sel))
    merged = eng.TsqrState(len(sel)) if world > 1 else None
    local.set_timing(True)
    tol_e = wl["tol_e"]

    def step_device():
        local.reset()
        local.accumulate_samples(dm, cfg_run, q, v, a, tau)
        st = fdist.allgather_merge(local, merged)
        return eng.solve_state(st, names, tol_e=tol_e, col_map=sel, p_full=cols)

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

    e2e = None
    if not args.no_e2e:
        # host copies for the end-to-end path: pinned (the caller's staging buffers) unless --pageable
        if args.pageable:
            qh, vh, ah, th = (x.cpu().numpy() for x in (q, v, a, tau))
        else:
            qh, vh, ah, th = (x.cpu().pin_memory() for x in (q, v, a, tau))

        def step_e2e():
            local.reset()
            local.accumulate_samples_host(dm, cfg_run, qh, vh, ah, th)
            st = fdist.allgather_merge(local, merged)
            return eng.solve_state(st, names, tol_e=tol_e, col_map=sel, p_full=cols)

        for _ in range(2):
            step_e2e()
        ms_e2e, sol_e = timed(step_e2e, max(3, args.steps // 2))
        h2d = int(N * (fm.nq + 2 * fm.nv + rps) * 8)
        e2e = {"value": (args.total_samples if strong else N * world) / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": h2d, "h2d_gbs_per_rank": h2d / (ms_e2e * 1e-3) / 1e9,
               "host_buffers": "pageable" if args.pageable else "pinned",
               "d2h_bytes_per_step": int(8 * (2 * (cols + 1) + len(sol.idx_base) * (len(sol.idx_regroup) + 1 + len(sol.idx_base))
                                              + (len(sol.keep) + 1) ** 2))}

    if rank == 0:
        total = args.total_samples if strong else N * world
        peak, peak_src = fp64_peak_tflops()
        plan = eng.describe_plan(fm, cfg_run)
        if not plan.get("ring_kernel"):  # wide kernel: count bodies / hops from the model
            depth = [0] * fm.njoints
            for i in range(1, fm.njoints):
                depth[i] = depth[int(fm.parent[i])] + 1
            plan["nbodies"] = fm.njoints - 1
            plan["hops"] = int(sum(depth))
        fps, f_reg = flops_per_sample(plan, rps, len(sol.keep))
        achieved = fps * N / (kern_ms_avg * 1e-3) / 1e12
        kernel = ("rq_accumulate_kernel (producer SMs walk the kinematic tree once per 32-sample tile and write every row group's "
                  "tile into an L2-resident ring; consumer SMs: one Householder TSQR stream per (warp, row group), cp.async.bulk "
                  "tile refill, 8-column panels, DMMA trailing updates)") if plan.get("ring_kernel") else \
                 ("matz_kernel + cq_accumulate_dense_kernel (kernel (1) materialises W a chunk of samples at a time: one sample "
                  "per lane, one walk of the tree for all row blocks, sector-aligned 14-column patches; kernel (2) is a "
                  "cooperative blocked Householder TSQR: 64-row tiles in shared memory, column blocks owned by warps, panel "
                  "look-ahead, DMMA trailing updates, factor in an L2-resident scratch); timed together")
        diag = np.asarray(sol.diag[:len(sol.keep)])
        base_diag = diag[diag > 1e-8]
        line = {
            "metric": METRIC, "value": total / (ms_step * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(wl, total if strong else N, per_gpu=not strong),
                       "samples_per_gpu": N, "rows_per_sample": rps, "columns": cols, "columns_folded": int(len(sel)),
                       "parallelism": f"dp{world}",
                       "l2": "inputs (%d B/sample x N = %.0f MB per step) exceed the 126 MB L2" % (
                           (fm.nq + 2 * fm.nv + rps) * 8, N * (fm.nq + 2 * fm.nv + rps) * 8 / 1e6),
                       "eliminated": len(sol.idx_e), "kept": len(sol.keep), "base_parameters": len(sol.idx_base),
                       "cond_R1": float(np.linalg.cond(sol.R1)),
                       "min_base_diag_over_tol_qr": float(base_diag.min() / 1e-8) if len(base_diag) else None,
                       "max_dependent_diag_over_tol_qr": float(diag[diag <= 1e-8].max() / 1e-8) if (diag <= 1e-8).any() else None},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "pipe": "fp64 (DFMA + DMMA m8n8k4; tcgen05 has no fp64 mode)",
                         "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.config, N), "kernel_ms": kern_ms_avg,
                         "kernel_share_of_step": kern_ms_avg / ms_step, "flops_per_sample": fps,
                         "flops_note": "SURVEY 8(d) dense formula F_reg + 2 m_s (p_e+1)^2 with p_e = kept columns; the row-group "
                                       "streams do fewer real flops (block sparsity), see DESIGN.md", "peak_source": peak_src},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if not args.no_parity_check:
            try:
                line["parity_check"] = parity_check(eng, torch, dm, fm, wl, cfg, sel, names, cols)
            except Exception as ex:  # noqa: BLE001
                line["parity_check"] = {"error": repr(ex)[:300]}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl, args.cpu_baseline_samples or wl["cpu_samples"])
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
