"""CPU prototype (numpy) of the two-stage symmetric tridiagonalisation the device eigen-solver implements
(csrc/sbr_kernels.cu): dense -> band (blocked Householder panels, two-sided rank-2b updates), band ->
tridiagonal (bulge chasing on band storage, one chain of b x b tasks per sweep), and the back-transformation
of tridiagonal eigenvectors through both stages.  It mirrors the device data structures and task order one to
one, including a randomised schedule that honours only the inter-sweep dependency the persistent kernel waits
on, so the index conventions can be checked without a GPU.  Not part of the product or of the oracle."""
import sys

import numpy as np


def house(x):
    """LAPACK dlarfg: H = I - tau v v^T with v[0] = 1 maps x to (beta, 0, ...)."""
    alpha = x[0]
    xn = np.linalg.norm(x[1:]) if x.shape[0] > 1 else 0.0
    v = np.zeros_like(x)
    v[0] = 1.0
    if xn == 0.0:
        return v, 0.0, alpha
    beta = -np.copysign(np.hypot(alpha, xn), alpha)
    tau = (beta - alpha) / beta
    v[1:] = x[1:] / (alpha - beta)
    return v, tau, beta


def stage1(A, b):
    """dense -> band.  Returns (B full symmetric with bandwidth b, list of (r0, V, T))."""
    A = A.copy()
    N = A.shape[0]
    panels = []
    c0 = 0
    while True:
        r0 = c0 + b
        n = N - r0
        if n < 2:
            break
        P = A[r0:, c0:c0 + b].copy()
        nb = P.shape[1]
        V = np.zeros((n, nb))
        T = np.zeros((nb, nb))
        for j in range(min(n, nb)):
            v, tau, beta = house(P[j:, j])
            V[j:, j] = v
            P[j, j] = beta
            P[j + 1:, j] = 0.0
            if tau != 0.0 and j + 1 < nb:
                w = v @ P[j:, j + 1:]
                P[j:, j + 1:] -= tau * np.outer(v, w)
            T[j, j] = tau
            if j > 0:
                T[:j, j] = -tau * (T[:j, :j] @ (V[:, :j].T @ V[:, j]))
        A[r0:, c0:c0 + nb] = P
        A[c0:c0 + nb, r0:] = P.T
        A22 = A[r0:, r0:]
        Y = A22 @ V @ T
        S = T.T @ (V.T @ Y)
        W = Y - 0.5 * V @ S
        A22 -= V @ W.T + W @ V.T
        panels.append((r0, V, T))
        c0 += b
    return A, panels


def to_band(B, b):
    """lower band storage with room for the bulge: Bb[c, d] = B[c + d, c], d = 0 .. 2b."""
    N = B.shape[0]
    Bb = np.zeros((N, 2 * b + 1))
    for c in range(N):
        for d in range(0, min(b, N - 1 - c) + 1):
            Bb[c, d] = B[c + d, c]
    return Bb


class Band(object):
    """symmetric access to the lower band storage"""

    def __init__(self, Bb):
        self.Bb = Bb

    def get(self, rows, cols):
        out = np.zeros((len(rows), len(cols)))
        for a, r in enumerate(rows):
            for c_, c in enumerate(cols):
                lo, hi = (c, r) if r >= c else (r, c)
                out[a, c_] = self.Bb[lo, hi - lo] if hi - lo < self.Bb.shape[1] else 0.0
        return out

    def put_lower(self, rows, cols, M):
        for a, r in enumerate(rows):
            for c_, c in enumerate(cols):
                if r >= c:
                    assert r - c < self.Bb.shape[1], (r, c)
                    self.Bb[c, r - c] = M[a, c_]


def jrange(i, k, b, N):
    lo = i + 1 + k * b
    hi = min(i + (k + 1) * b, N - 1)
    return list(range(lo, hi + 1))


def ntasks(i, b, N):
    """number of tasks of sweep i: blocks J_k that are non-empty"""
    return max(0, -(-(N - 1 - i) // b))


def task(band, i, k, b, N, V2, TAU2, state):
    """task k of sweep i.  state[i] carries the reflector handed from task k-1 to task k."""
    J = jrange(i, k, b, N)
    if k == 0:
        x = np.array([band.Bb[i, r - i] for r in J])
        v, tau, beta = house(x)
        band.Bb[i, 1] = beta
        for r in J[1:]:
            band.Bb[i, r - i] = 0.0
    else:
        v, tau = state[i]
    V2[i, J[0]:J[0] + len(J)] = v
    TAU2[i, k] = tau
    # two-sided on the diagonal block
    D = band.get(J, J)
    H = np.eye(len(J)) - tau * np.outer(v, v)
    band.put_lower(J, J, H @ D @ H)
    Jn = jrange(i, k + 1, b, N)
    if Jn:
        E = band.get(Jn, J) @ H
        v2, tau2, beta2 = house(E[:, 0].copy())
        H2 = np.eye(len(Jn)) - tau2 * np.outer(v2, v2)
        E = H2 @ E
        E[0, 0] = beta2
        E[1:, 0] = 0.0
        band.put_lower(Jn, J, E)
        state[i] = (v2, tau2)


def stage2(Bb, b, schedule="sequential", seed=0):
    """band -> tridiagonal by bulge chasing.  Returns d, e, V2 (N x N: row = sweep), TAU2 (N x ntask)."""
    Bb = Bb.copy()
    N = Bb.shape[0]
    band = Band(Bb)
    V2 = np.zeros((N, N))
    TAU2 = np.zeros((N, ntasks(0, b, N) + 1))
    state = {}
    nsweep = max(0, N - 2)
    nt = [ntasks(i, b, N) for i in range(nsweep)]
    if schedule == "sequential":
        for i in range(nsweep):
            for k in range(nt[i]):
                task(band, i, k, b, N, V2, TAU2, state)
    else:
        # any order that keeps a sweep's tasks in order and starts (i, k) only after sweep i-1 has finished
        # task k+1 (or all of its tasks): what the persistent kernel's progress counters enforce
        rs = np.random.RandomState(seed)
        prog = [0] * nsweep
        while any(prog[i] < nt[i] for i in range(nsweep)):
            ready = [i for i in range(nsweep) if prog[i] < nt[i] and
                     (i == 0 or prog[i - 1] >= min(prog[i] + 2, nt[i - 1]))]
            i = ready[rs.randint(len(ready))]
            task(band, i, prog[i], b, N, V2, TAU2, state)
            prog[i] += 1
    d = Bb[:, 0].copy()
    e = Bb[:-1, 1].copy()
    return d, e, V2, TAU2, Bb


def apply_q2(z, V2, TAU2, b):
    """z <- Q2 z: reflectors in reverse order of generation (sweeps descending; tasks of one sweep commute)."""
    N = z.shape[0]
    for i in range(N - 3, -1, -1):
        for k in range(ntasks(i, b, N)):
            J = jrange(i, k, b, N)
            v = V2[i, J[0]:J[0] + len(J)]
            tau = TAU2[i, k]
            z[J] -= tau * np.outer(v, v @ z[J]) if z.ndim == 2 else tau * v * (v @ z[J])
    return z


def apply_q1(z, panels):
    for r0, V, T in reversed(panels):
        z[r0:] -= V @ (T @ (V.T @ z[r0:]))
    return z


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 61
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rs = np.random.RandomState(1)
    X = rs.standard_normal((N + 7, N)) * np.linspace(1.0, 0.01, N)[None, :]
    A = X.T @ X
    B, panels = stage1(A, b)
    assert np.abs(np.tril(B, -b - 1)).max() < 1e-13 * np.abs(A).max()
    lamA = np.linalg.eigvalsh(A)
    print("stage 1: band eigenvalues", np.abs(np.linalg.eigvalsh(B) - lamA).max() / lamA.max())
    Bb = to_band(B, b)
    d, e, V2, TAU2, Bfin = stage2(Bb, b)
    assert np.abs(Bfin[:, 2:]).max() < 1e-13 * np.abs(A).max(), np.abs(Bfin[:, 2:]).max()
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    lamT, Z = np.linalg.eigh(Tm)
    print("stage 2: tridiagonal eigenvalues", np.abs(lamT - lamA).max() / lamA.max())
    d2, e2, V2b, TAU2b, _ = stage2(Bb, b, schedule="random", seed=3)
    print("random schedule == sequential:", np.array_equal(d, d2) and np.array_equal(e, e2) and np.array_equal(V2, V2b))
    Zb = apply_q2(Z.copy(), V2, TAU2, b)
    print("Q2: band residual", np.abs(B @ Zb - Zb * lamT[None, :]).max() / lamA.max())
    Za = apply_q1(Zb.copy(), panels)
    print("Q1: dense residual", np.abs(A @ Za - Za * lamT[None, :]).max() / lamA.max(),
          "orthogonality", np.abs(Za.T @ Za - np.eye(N)).max())


if __name__ == "__main__":
    main()
