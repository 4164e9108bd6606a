"""GPU: kernels added in round 2, each against the oracle / LAPACK.

* cooperative blocked TSQR (fg_coop.cuh: the drop-in double_QR path for wide W_e, qrdecomposition.py:89-187, and the
  Talos external-wrench fold): R^T R, squared column norms and the least-squares solution against numpy on the same
  matrix - widths that take 64-row tiles (p + 1 <= 392) and 32-row tiles (wider), ragged row counts, chunked calls;
* materialise kernel (fg_matz.cuh: build_regressor_basic, regressor.py:20-194): ragged sample counts, sector-aligned
  and unaligned leading dimensions, padded ldw, against the oracle's regressor.
Tolerances: 1e-9 relative (fp64 Householder against LAPACK Householder), entries of W 1e-9 relative to max |W|.
"""
import numpy as np
import pytest

from conftest import TX40_PARAM, load_flat, load_oracle_model, make_param, rand_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,p", [(3000, 130), (4097, 330), (1000, 330), (2500, 420), (70, 200)])
def test_cooperative_tsqr_matches_lapack(rows, p):
    import torch
    from figaroh_b200 import engine as eng
    rng = np.random.default_rng(rows + p)
    W = rng.normal(size=(rows, p)) * rng.uniform(0.1, 10.0, size=(1, p))
    W[:, 5] = 0.0                                    # a structurally zero column
    W[:, 17] = W[:, 3] - 2.0 * W[:, 11]              # an exactly dependent one
    tau = rng.normal(size=rows)
    st = eng.TsqrState(p).accumulate_dense(W, tau)
    assert st.rows == rows
    A = np.c_[W, tau]
    np.testing.assert_allclose(st.colsq(), (A * A).sum(0), rtol=1e-12, atol=1e-12)
    # the factor: R^T R = A^T A (signs of the rows of R are not unique)
    packed = st.export().cpu().numpy()
    p1 = p + 1
    R = packed[:p1 * p1].reshape(p1, p1)
    G = A.T @ A
    assert np.abs(R.T @ R - G).max() <= 1e-11 * np.abs(G).max()
    assert np.abs(np.tril(R, -1)).max() == 0.0
    # |diag R| against LAPACK's: unique only as long as the leading columns are independent (columns 0..4 here; past
    # a zero / dependent column the row of R that belongs to it is an arbitrary direction in any Householder QR), so
    # beyond them only the rank decision of qrdecomposition.py:117-121 is compared
    if rows >= p1:
        Rl = np.linalg.qr(A, mode="r")
        d, dl = np.abs(np.diag(R)), np.abs(np.diag(Rl))
        np.testing.assert_allclose(d[:5], dl[:5], rtol=1e-10)
        tol = 1e-8 * dl.max()
        assert list(d > tol) == list(dl > tol) and not (d[5] > tol) and not (d[17] > tol)
    # two calls continue the same factor
    h = rows // 2
    st2 = eng.TsqrState(p).accumulate_dense(W[:h], tau[:h]).accumulate_dense(W[h:], tau[h:])
    R2 = st2.export().cpu().numpy()[:p1 * p1].reshape(p1, p1)
    assert np.abs(R2.T @ R2 - G).max() <= 1e-11 * np.abs(G).max()


def test_wide_drop_in_double_qr_matches_oracle():
    """double_QR on a 300-column W_e (the reference: two np.linalg.qr of the tall matrix) through the cooperative kernel."""
    from figaroh.tools.qrdecomposition import double_QR
    from oracle import ref_path
    rng = np.random.default_rng(7)
    rows, p = 2400, 300
    W = rng.normal(size=(rows, p))
    for k in range(0, 40):                           # 40 dependent columns, regrouped onto earlier ones
        W[:, 7 * k + 6] = W[:, 7 * k] * 0.5 - W[:, 7 * k + 2] * 3.0
    names = [f"c{i}" for i in range(p)]
    std = {n: float(i % 7) - 3.0 for i, n in enumerate(names)}
    tau = W @ np.array(list(std.values())) + 1e-3 * rng.normal(size=rows)
    got = double_QR(tau, W, names, std)
    ref = ref_path.double_QR(tau, W, names, std)
    assert got[2] == ref[2]                          # expression strings
    np.testing.assert_allclose(got[3], ref[3], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("model,N,pad", [("ur10", 77, 0), ("ur10", 1, 0), ("ur10", 4001, 4), ("ur10", 333, 1),
                                         ("tiago", 95, 0), ("talos_ff", 50, 0)])
def test_materialise_edge_cases_match_oracle(model, N, pad):
    import torch
    from figaroh_b200 import engine as eng
    from oracle import ref_path
    fm, om = load_flat(model), load_oracle_model(model)
    if model == "talos_ff":
        param = make_param(False, False, False, is_joint_torques=False, is_external_wrench=True, force_torque=["All"])
    else:
        param = make_param(True, True, True)
    q, v, a = rand_inputs(fm, N, 5, ((-2, 2), (-2, 2), (-4, 4)))
    Wref = ref_path.build_regressor_basic(om, q, v, a, param, per_sample=False)
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(param)
    rps, cols = dm.shape(eng.RegressorConfig(cfg.mode, cfg.flags, cfg.ft_mask))
    assert Wref.shape == (rps * N, cols)
    buf = torch.full((rps * N, cols + pad), -7.0, dtype=torch.float64, device="cuda")
    out = buf[:, :cols] if pad else buf
    W = dm.materialize(cfg, q, v, a, out=out).cpu().numpy()
    assert np.abs(W - Wref).max() <= 1e-9 * np.abs(Wref).max()
    if pad:
        assert float(buf[:, cols:].min()) == -7.0 and float(buf[:, cols:].max()) == -7.0   # padding untouched
