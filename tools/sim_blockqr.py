"""Numpy model of the block (compact-WY) TS-Householder fold used by the sm_100a kernel.

Checks the algebra the CUDA kernel relies on (test infrastructure, not shipped):
  * reflector j of a panel: u-form scalars, x_hat = x / (alpha - beta), tau = |alpha-beta| / nrm
  * V = [I_8 ; X_hat],  T^-1 = diag(1/tau) + striu(X_hat^T X_hat), column-wise back substitution
  * trailing update  Y = R_kj + X_hat^T B_j,  W = T^T Y,  R_kj -= W,  B_j -= X_hat W
"""
import numpy as np


def fold_tile(R, B, nb=8):
    """R: (P,P) upper triangular (modified in place), B: (T,P) tile (destroyed)."""
    P = R.shape[0]
    for k in range(P // nb):
        c0 = nb * k
        Xh = np.zeros((B.shape[0], nb))
        tau = np.zeros(nb)
        S = np.zeros((nb, nb))
        for j in range(nb):
            x = B[:, c0 + j].copy()
            alpha = R[c0 + j, c0 + j]
            sigma = x @ x
            if sigma == 0.0:
                continue
            mu = alpha * alpha + sigma
            nrm = np.sqrt(mu)
            beta = -nrm if alpha >= 0 else nrm
            den = alpha - beta
            inv = 1.0 / den
            tau[j] = abs(den) / nrm
            # dots with every panel column (own columns: later = current data, earlier = x_hat)
            for m in range(nb):
                if m > j:
                    d = x @ B[:, c0 + m]
                    rho = R[c0 + j, c0 + m]
                    y = rho + d * inv
                    w = tau[j] * y
                    R[c0 + j, c0 + m] = rho - w
                    B[:, c0 + m] -= (w * inv) * x
                elif m < j:
                    S[m, j] = (x @ Xh[:, m]) * inv
            Xh[:, j] = x * inv
            R[c0 + j, c0 + j] = beta
            B[:, c0 + j] = 0.0
        # T by column-wise back substitution of (diag(1/tau) + striu(S)) t = e_c
        T = np.zeros((nb, nb))
        for c in range(nb):
            t = np.zeros(nb)
            t[c] = tau[c]
            for i in range(c - 1, -1, -1):
                acc = 0.0
                for m in range(i + 1, c + 1):
                    acc += S[i, m] * t[m]
                t[i] = -tau[i] * acc
            T[:, c] = t
        for j0 in range(c0 + nb, P, nb):
            Y = R[c0:c0 + nb, j0:j0 + nb] + Xh.T @ B[:, j0:j0 + nb]
            Wm = T.T @ Y
            R[c0:c0 + nb, j0:j0 + nb] -= Wm
            B[:, j0:j0 + nb] -= Xh @ Wm
    return R


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    P, T, ntile = 80, 24, 40
    p1 = 74
    A = rng.standard_normal((T * ntile, P))
    A[:, p1:] = 0.0
    # block-triangular structure + a nearly dependent column
    A[:, 20] = A[:, 3] * 2 + 1e-9 * rng.standard_normal(T * ntile)
    A[::3, :10] = 0.0
    R = np.zeros((P, P))
    for t in range(ntile):
        fold_tile(R, A[t * T:(t + 1) * T].copy())
    Rref = np.linalg.qr(A[:, :p1], mode="r")
    d = np.abs(np.abs(np.diag(R)[:p1]) - np.abs(np.diag(Rref)))
    print("max |diag| diff", d.max(), " small diag:", abs(R[20, 20]), abs(Rref[20, 20]))
    G = R[:p1, :p1].T @ R[:p1, :p1]
    G0 = A[:, :p1].T @ A[:, :p1]
    print("rel gram err", np.abs(G - G0).max() / np.abs(G0).max())
    sgn = np.sign(np.diag(R)[:p1]) * np.sign(np.diag(Rref))
    print("max |R - Rref| (sign-fixed)", np.abs(R[:p1, :p1] * sgn[:, None] - Rref).max())
