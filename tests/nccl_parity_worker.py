"""torchrun worker of tests/test_gpu_parity_oracle.py::test_nccl_ranks_reproduce_single_gpu_solution.

Every rank folds its shard of the SAME seeded UR10 problem with the fused kernel, the packed factors are
all-gathered over NCCL and merged on every rank; rank 0 also folds the whole problem alone and runs the CPU
oracle.  Index sets must be identical, phi_b within 1e-9 (well-conditioned flags) of both.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "figaroh-plus_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    from conftest import TX40_PARAM, load_flat, load_oracle_model, make_param, rand_inputs
    from figaroh_b200 import dist as fdist
    from figaroh_b200 import engine as eng

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    fm = load_flat("ur10")
    N = 50_001
    q, v, a = rand_inputs(fm, N, 77, ((-3, 3), (-3, 3), (-6, 6)))
    param = make_param(True, False, True, fv=TX40_PARAM["fv"], fs=TX40_PARAM["fs"], off=TX40_PARAM["off"])
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(param)
    _, cols = dm.shape(cfg)
    names = [f"c{i}" for i in range(cols)]
    rng = np.random.default_rng(1)
    tau = rng.normal(size=(fm.nv, N))
    lo, hi = fdist.shard_range(N, rank, world)
    local = eng.TsqrState(cols).accumulate_samples(dm, cfg, q[lo:hi], v[lo:hi], a[lo:hi], np.ascontiguousarray(tau[:, lo:hi]))
    merged = fdist.allgather_merge(local)
    sol = eng.solve_state(merged, names, tol_e=1e-6)
    # every rank must hold the same solution bit for bit (same merge order everywhere)
    mine = torch.from_numpy(np.ascontiguousarray(sol.phi_b_raw)).cuda()
    allphi = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allphi, mine)
    same = all(torch.equal(allphi[0], x) for x in allphi)
    ok = same and merged.rows == fm.nv * N
    if rank == 0:
        single = eng.TsqrState(cols).accumulate_samples(dm, cfg, q, v, a, tau)
        ref = eng.solve_state(single, names, tol_e=1e-6)
        ok &= sol.idx_e == ref.idx_e and sol.idx_base == ref.idx_base and sol.params_base == ref.params_base
        rel = np.abs(sol.phi_b_raw - ref.phi_b_raw).max() / np.abs(ref.phi_b_raw).max()
        ok &= rel < 1e-9
        # oracle on a subsample-independent check: index sets of the full problem
        from oracle import ref_path
        om = load_oracle_model("ur10")
        Ns = 4000
        W = ref_path.build_regressor_basic(om, q[:Ns], v[:Ns], a[:Ns], param, per_sample=False)
        std = ref_path.standard_parameters(om, param)
        idx_e, params_r = ref_path.get_index_eliminate(W, std, 1e-6)
        base = ref_path.get_baseIndex(ref_path.build_regressor_reduced(W, idx_e), params_r)
        ok &= sol.idx_e == idx_e and tuple(sol.idx_base) == tuple(base)
        print(f"rel={rel:.3e} same_on_all_ranks={same} idx_e={len(sol.idx_e)} base={len(sol.idx_base)}")
        print("NCCL_PARITY_OK" if ok else "NCCL_PARITY_FAILED", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
