"""CPU: the symbolic pass of the fused path (fg_plan_describe, no device needed).

The generation-5 kernel's producers execute a host-built depth-first walk program (descend / ascend steps with
saved states for side branches) and write one local column problem per row group.  Here that program is
EXECUTED IN NUMPY exactly as the CUDA producer executes it (same frames, same transports, same entry formulas)
and the resulting row-group tiles are compared with the oracle's regressor (reference semantics:
regressor.py:49-87 over pinocchio.computeJointTorqueRegressor) - so an error in the walk program of a tree
(TIAGo: head and gripper branches) is caught without a GPU.
"""
import numpy as np
import pytest

from conftest import TX40_PARAM, load_flat, load_oracle_model, make_param, rand_inputs


def _describe(fm, cfg):
    from figaroh_b200 import engine as eng
    return eng.describe_plan(fm, cfg)


def _cross(a, b):
    return np.cross(a, b)


def _axis_rot(ax, c, s):
    ax = np.asarray(ax, float)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    A = np.outer(ax, ax)
    n = len(c)
    return (1 - c)[:, None, None] * A[None] + c[:, None, None] * np.eye(3)[None] + s[:, None, None] * K[None]


def _quat_rot(x, y, z, w):
    R = np.empty((len(x), 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _actinv(R, p, lin, ang):
    t = lin - _cross(p, ang)
    return np.einsum("nji,nj->ni", R, t), np.einsum("nji,nj->ni", R, ang)


def _act(R, p, lin, ang):
    ra = np.einsum("nij,nj->ni", R, ang)
    rl = np.einsum("nij,nj->ni", R, lin)
    return rl + _cross(p, ra), ra


def _emit(sl, sa, om, al, ap):
    """s^T B for the ten pinocchio parameters (fg_ringqr.cuh::rq_emit)."""
    val = np.empty((len(sl), 10))
    val[:, 0] = np.sum(sl * ap, 1)
    ww = np.sum(om * om, 1)
    so = np.sum(sl * om, 1)
    val[:, 1:4] = _cross(sl, al) + so[:, None] * om - ww[:, None] * sl + _cross(ap, sa)
    h = _cross(sa, om)
    val[:, 4] = sa[:, 0] * al[:, 0] + h[:, 0] * om[:, 0]
    val[:, 5] = sa[:, 0] * al[:, 1] + sa[:, 1] * al[:, 0] + h[:, 0] * om[:, 1] + h[:, 1] * om[:, 0]
    val[:, 6] = sa[:, 1] * al[:, 1] + h[:, 1] * om[:, 1]
    val[:, 7] = sa[:, 0] * al[:, 2] + sa[:, 2] * al[:, 0] + h[:, 0] * om[:, 2] + h[:, 2] * om[:, 0]
    val[:, 8] = sa[:, 1] * al[:, 2] + sa[:, 2] * al[:, 1] + h[:, 1] * om[:, 2] + h[:, 2] * om[:, 1]
    val[:, 9] = sa[:, 2] * al[:, 2] + h[:, 2] * om[:, 2]
    return val


def run_walk_program(fm, d, q, v, a):
    """Execute the producers' program; returns one (N, w_g) tile per row group."""
    N = len(q)
    walk, groups = d["walk"], d["groups"]
    tiles = [np.zeros((N, g["w"])) for g in groups]
    emit = [{e[0]: (e[1], e[2]) for e in g["emit"]} for g in groups]
    AX = {}
    VS, MS = {}, {}
    vv = np.zeros((N, 6)); aa = np.zeros((N, 6))
    for op in d["ops"]:
        kind, t, vsave, msave, vload, mslot, from_u, live, own, _ = op
        if kind == 0:
            j = walk[t]
            if vsave != 255:
                VS[vsave] = (vv.copy(), aa.copy())
            if from_u:
                vv = np.zeros((N, 6)); aa = np.zeros((N, 6)); aa[:, :3] = -fm.gravity
            iq, iv, jt, ax = int(fm.idx_q[j]), int(fm.idx_v[j]), int(fm.jtype[j]), fm.axis[j]
            Rj = np.tile(np.eye(3), (N, 1, 1)); pj = np.zeros((N, 3)); vj = np.zeros((N, 6)); aj = np.zeros((N, 6))
            if jt == 1:
                Rj = _axis_rot(ax, np.cos(q[:, iq]), np.sin(q[:, iq]))
                vj[:, 3:] = ax[None] * v[:, iv:iv + 1]; aj[:, 3:] = ax[None] * a[:, iv:iv + 1]
            elif jt == 2:
                Rj = _axis_rot(ax, q[:, iq], q[:, iq + 1])
                vj[:, 3:] = ax[None] * v[:, iv:iv + 1]; aj[:, 3:] = ax[None] * a[:, iv:iv + 1]
            elif jt == 3:
                pj = ax[None] * q[:, iq:iq + 1]
                vj[:, :3] = ax[None] * v[:, iv:iv + 1]; aj[:, :3] = ax[None] * a[:, iv:iv + 1]
            else:
                Rj = _quat_rot(q[:, iq + 3], q[:, iq + 4], q[:, iq + 5], q[:, iq + 6])
                pj = q[:, iq:iq + 3].copy(); vj = v[:, iv:iv + 6].copy(); aj = a[:, iv:iv + 6].copy()
            P = fm.plc_R[j].reshape(3, 3)
            R = np.einsum("ij,njk->nik", P, Rj)
            p = pj @ P.T + fm.plc_p[j][None]
            if msave != 255:
                MS[msave] = (R.copy(), p.copy())
            tl, ta = _actinv(R, p, vv[:, :3], vv[:, 3:])
            vn = np.c_[vj[:, :3] + tl, vj[:, 3:] + ta]
            l1 = _cross(vn[:, 3:], vj[:, :3]); l2 = _cross(vn[:, :3], vj[:, 3:]); l3 = _cross(vn[:, 3:], vj[:, 3:])
            tl, ta = _actinv(R, p, aa[:, :3], aa[:, 3:])
            an = np.c_[l1 + l2 + aj[:, :3] + tl, l3 + aj[:, 3:] + ta]
            vv, aa = vn, an
            ap = _cross(vv[:, 3:], vv[:, :3]) + aa[:, :3]
            for g in range(len(groups)):
                if (live >> g) & 1:
                    AX[g] = _actinv(R, p, *AX[g])
                elif (own >> g) & 1:
                    S = np.asarray(groups[g]["S"])
                    AX[g] = (np.tile(S[:3], (N, 1)), np.tile(S[3:], (N, 1)))
                else:
                    continue
                if t in emit[g]:
                    base, mask = emit[g][t]
                    val = _emit(AX[g][0], AX[g][1], vv[:, 3:], aa[:, 3:], ap)
                    slot = base
                    for k in range(10):
                        if (mask >> k) & 1:
                            tiles[g][:, slot] = val[:, k]
                            slot += 1
        else:
            R, p = MS[mslot]
            for g in range(len(groups)):
                if (live >> g) & 1:
                    AX[g] = _act(R, p, *AX[g])
            if vload != 255:
                vv, aa = VS[vload][0].copy(), VS[vload][1].copy()
    for t, g, slot, src in d.get("stepcols", []):  # the joint's own input-fed columns, written at its walk step
        vidx = int(fm.idx_v[walk[t]])
        tiles[g][:, slot] = (a[:, vidx], v[:, vidx], np.sign(v[:, vidx]), np.ones(N))[src]
    for g, G in enumerate(groups):
        for kind, src, vidx, slot in G["simple"]:
            if kind == 2:  # COL_EXTRA
                tiles[g][:, slot] = (a[:, vidx], v[:, vidx], np.sign(v[:, vidx]), np.ones(N))[src]
            elif kind == 3:  # COL_COUPLING
                sg = np.sign(v[:, 4] + v[:, 5])
                if G["rowblock"] == 4:
                    tiles[g][:, slot] = (a[:, 5], v[:, 5], sg)[src]
                else:
                    tiles[g][:, slot] = (a[:, 4], v[:, 4], sg)[src]
    return tiles


def _check_against_oracle(model, param, rows=None, coupling=False, seed=3, lo_hi=((-2, 2), (-2, 2), (-4, 4))):
    from figaroh_b200 import engine as eng
    from oracle import ref_path
    fm, om = load_flat(model), load_oracle_model(model)
    N = 40
    q, v, a = rand_inputs(fm, N, seed, lo_hi)
    W = ref_path.build_regressor_basic(om, q, v, a, param, per_sample=False)
    if coupling:
        W = ref_path.add_coupling_TX40(W, N, v, a)
    nblk = W.shape[0] // N
    blocks = list(range(nblk)) if rows is None else list(rows)
    Wsel = np.concatenate([W[r * N:(r + 1) * N] for r in blocks])
    sel = [int(c) for c in np.flatnonzero(np.abs(Wsel).max(0) != 0)]  # what structural_columns() finds on the GPU
    cfg = eng.RegressorConfig.from_param(param, coupling=coupling, row_sel=rows, col_sel=sel)
    d = _describe(fm, cfg)
    assert d["ring_kernel"], d["why"]
    tiles = run_walk_program(fm, d, q, v, a)
    scale = np.abs(W).max()
    covered = np.zeros((len(blocks), len(sel)), bool)
    for g, G in enumerate(d["groups"]):
        rb = G["rowblock"]
        Wg = W[rb * N:(rb + 1) * N]
        for l, sc in enumerate(G["cols"][:-1] if G["cols"][-1] == len(sel) else G["cols"]):
            if sc == len(sel):
                continue
            ref = Wg[:, sel[sc]]
            assert np.abs(tiles[g][:, l] - ref).max() <= 1e-11 * scale, (g, l, sc)
            covered[blocks.index(rb) if rows is not None else rb, sc] = True
        assert G["cols"][-1] == len(sel)  # the rhs is the last local column
        assert G["nb"] == (G["w"] + 7) // 8
    # what no row group carries must be identically zero in that row block
    for bi, rb in enumerate(blocks):
        Wg = W[rb * N:(rb + 1) * N][:, sel]
        assert np.abs(Wg[:, ~covered[bi]]).max(initial=0.0) == 0.0
    return d


def test_ur10_chain_program_matches_oracle():
    d = _check_against_oracle("ur10", make_param(True, True, True))
    assert d["walk"] == [1, 2, 3, 4, 5, 6] and d["nvslots"] == 0 and d["nmslots"] == 0
    w = [g["w"] for g in d["groups"]]
    assert w[2:] == [45, 35, 25, 15] and w[0] in (54, 55) and w[1] in (53, 54)  # 1e-17 leaks of the pi/2 roundings
    assert sum(g["ninst"] for g in d["groups"]) == (148 - d["producer_ctas"]) * d["nw"]


def test_tx40_coupling_program_matches_oracle():
    _check_against_oracle("tx40", dict(TX40_PARAM), coupling=True, lo_hi=((-6, 6), (-10, 10), (-30, 30)))


def test_tiago_tree_program_matches_oracle():
    fm = load_flat("tiago")
    rows = [int(fm.idx_v[i]) for i in range(13, 21)]
    d = _check_against_oracle("tiago", make_param(True, True, True), rows=rows, lo_hi=((-1.5, 1.5), (-2, 2), (-4, 4)))
    # torso, head branch (walked first, then back up), arm chain, first finger (side trip), second finger
    assert d["walk"] == [13, 23, 24, 14, 15, 16, 17, 18, 19, 20, 21, 22]
    assert d["nvslots"] == 1 and d["nmslots"] == 2
    kinds = [op[0] for op in d["ops"]]
    assert kinds == [0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0]
    # the prismatic torso row never sees inertia-tensor columns: at most 12 bodies x (m, mx, my, mz) + Ia fv fs off + rhs
    assert d["groups"][0]["w"] <= 12 * 4 + 4 + 1 < d["groups"][1]["w"]


def test_free_flyer_row_groups_program_matches_oracle():
    """Talos in external-wrench mode restricted to the root body and one leg: six row groups on ONE joint."""
    from figaroh_b200 import engine as eng
    fm = load_flat("talos_ff")
    param = dict(is_joint_torques=False, is_external_wrench=True, force_torque=["All"], has_friction=False,
                 has_actuator_inertia=False, has_joint_offset=False, fv=None, fs=None, Ia=None, off=None)
    cols = [c for b in range(0, 7) for c in range(14 * b, 14 * b + 10)]
    cfg = eng.RegressorConfig.from_param(param, col_sel=cols)
    d = _describe(fm, cfg)
    assert d["ring_kernel"], d["why"]
    from oracle import ref_path
    om = load_oracle_model("talos_ff")
    N = 24
    q, v, a = rand_inputs(fm, N, 5, ((-1, 1), (-1, 1), (-2, 2)))
    W = ref_path.build_regressor_basic(om, q, v, a, param, per_sample=False)
    tiles = run_walk_program(fm, d, q, v, a)
    for g, G in enumerate(d["groups"]):
        Wg = W[g * N:(g + 1) * N]
        for l, sc in enumerate(G["cols"][:-1]):
            assert np.abs(tiles[g][:, l] - Wg[:, cols[sc]]).max() <= 1e-11 * np.abs(W).max()
        # force rows (no angular part) drop the six inertia-tensor columns of every body
        assert G["w"] == (7 * 4 + 1 if g < 3 else 7 * 10 + 1)


def test_unsupported_shapes_are_refused_not_mangled():
    from figaroh_b200 import engine as eng
    fm = load_flat("tiago")
    d = _describe(fm, eng.RegressorConfig.from_param(make_param(True, True, True)))  # 24 row groups
    assert not d["ring_kernel"] and d["why"]
