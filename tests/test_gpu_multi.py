"""GPU: the multi-GPU reduction exercised on ONE device with simulated shards
(SURVEY.md section 4 (5)): split the samples into G shards, fold each into its own state,
export -> concatenate (what the NCCL all-gather produces) -> merge_packed, and require the
same index sets and base parameters as the single-state run."""
import numpy as np
import pytest

from conftest import TX40_PARAM, load_flat, make_param, rand_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("G", [2, 8])
def test_simulated_shards_match_single_state(G):
    import torch
    from figaroh_b200 import engine as eng
    from figaroh_b200.dist import shard_range
    fm = load_flat("ur10")
    N = 4001
    q, v, a = rand_inputs(fm, N, 21, ((-3, 3), (-3, 3), (-6, 6)))
    param = make_param(True, False, True)
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(param)
    _, cols = dm.shape(cfg)
    names = [f"c{i}" for i in range(cols)]
    rng = np.random.default_rng(2)
    tau = rng.normal(size=(6, N))
    single = eng.TsqrState(cols).accumulate_samples(dm, cfg, q, v, a, tau)
    ref = eng.solve_state(single, names, tol_e=1e-6)
    packs = []
    for g in range(G):
        lo, hi = shard_range(N, g, G)
        st = eng.TsqrState(cols).accumulate_samples(dm, cfg, q[lo:hi], v[lo:hi], a[lo:hi], np.ascontiguousarray(tau[:, lo:hi]))
        packs.append(st.export())
    merged = eng.TsqrState(cols).merge_packed(torch.cat(packs), G)
    assert merged.rows == single.rows == 6 * N
    sol = eng.solve_state(merged, names, tol_e=1e-6)
    assert sol.idx_e == ref.idx_e and sol.idx_base == ref.idx_base and sol.params_base == ref.params_base
    np.testing.assert_allclose(sol.phi_b_raw, ref.phi_b_raw, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(merged.colsq(), single.colsq(), rtol=1e-13)
    # export -> merge into an empty state is the identity on R^T R
    again = eng.TsqrState(cols).merge_packed(merged.export(), 1)
    np.testing.assert_allclose(eng.solve_state(again, names, tol_e=1e-6).phi_b_raw, sol.phi_b_raw, rtol=1e-11)
