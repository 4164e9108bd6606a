"""Device-time probe of the fused accumulate path on the TIAGo (8 active joints of the 24-DoF tree) and Talos
(free-flyer, external-wrench rows) shapes of BASELINE.json configs[2] and configs[3].  Prints one JSON line each."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "figaroh-plus_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    import torch
    from conftest import load_flat, make_param, rand_inputs
    from figaroh_b200 import engine as eng
    which = sys.argv[1] if len(sys.argv) > 1 else "tiago"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    if which == "tiago":
        fm = load_flat("tiago")
        rows = [int(fm.idx_v[i]) for i in range(13, 21)]
        param = make_param(True, True, True)
        q, v, a = rand_inputs(fm, N, 9, ((-1.5, 1.5), (-2, 2), (-4, 4)))
        active = set(rows)
        for j in range(fm.nv):
            if j not in active:
                v[:, j] = 0.0
                a[:, j] = 0.0
        for i in range(1, fm.njoints):
            if int(fm.idx_v[i]) not in active:
                iq = fm.idx_q[i]
                if fm.jtype[i] == 2:
                    q[:, iq], q[:, iq + 1] = 1.0, 0.0
                else:
                    q[:, iq] = 0.1
        cfg = eng.RegressorConfig.from_param(param, row_sel=rows)
    else:
        fm = load_flat("talos_ff")
        param = dict(is_joint_torques=False, is_external_wrench=True, force_torque=["All"], has_friction=False,
                     has_actuator_inertia=False, has_joint_offset=False, fv=None, fs=None, Ia=None, off=None)
        q, v, a = rand_inputs(fm, N, 13, ((-1, 1), (-1, 1), (-2, 2)))
        cfg = eng.RegressorConfig.from_param(param)
    dm = eng.DeviceModel(fm)
    rps, cols = dm.shape(cfg)
    sel = eng.structural_columns(dm, cfg)
    cfg_run = cfg.with_columns(sel)
    qd, vd, ad = (torch.from_numpy(x).cuda() for x in (q, v, a))
    tau = torch.randn((rps, N), dtype=torch.float64, device="cuda")
    st = eng.TsqrState(len(sel))
    names = [f"c{i}" for i in range(cols)]

    def step():
        st.reset()
        st.accumulate_samples(dm, cfg_run, qd, vd, ad, tau)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sol = eng.solve_state(st, names, tol_e=1e-6, col_map=sel, p_full=cols)
    pe = len(sel)
    flops = 2 * rps * (pe + 1) ** 2
    print(json.dumps({"model": which, "N": N, "rows_per_sample": rps, "columns_folded": pe, "ms": ms,
                      "samples_per_s": N / ms * 1e3, "qr_flops_per_sample": flops,
                      "tflops_dense_formula": flops * N / ms / 1e9, "frac_of_37.19": flops * N / ms / 1e9 / 37.19,
                      "eliminated": len(sol.idx_e), "base": len(sol.idx_base),
                      "describe": {k: v for k, v in eng.describe_plan(fm, cfg_run).items() if k in ("ring_kernel", "why", "nw", "producer_ctas", "smem")}}))


if __name__ == "__main__":
    main()
