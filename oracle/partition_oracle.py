"""TEST INFRASTRUCTURE -- NumPy restatement of the partitioned block-tridiagonal solve of vinsat_b200/csrc/kernels_chain.cu
(k_chain_forward3<SPIKE>, k_seg_backrec, k_reduced_build, k_chain_backward, k_seg_backsub and the second partition level of
`Level2`, csrc/batch.h).  Not a restatement of the reference: the reference solves the damped normal equations with a dense
`torch.linalg.solve` (BA_filtering.py:60-64); this file pins the ALGEBRA the kernels use to reach the same solution, so that a
failing GPU test can be split into "the algorithm" (checked here on the CPU against a dense solve) and "its implementation".

System: block rows i = 0..n-1 of 9x9 blocks,  L[i-1] x_{i-1} + D[i] x_i + U[i] x_{i+1} = b[i]   (L[i] = A(i+1, i), U[i] = A(i, i+1);
the BA system has L[i] = U[i]^T, the reduced systems do not).

A segment k owns the interior elements [a_k, s_k) and the separator s_k; its left separator is s_{k-1} (none for k = 0).
  forward  (interior, left to right)   S_i = D_i - L_{i-1} W_{i-1},  W_i = S_i^-1 U_i,  y_i = S_i^-1 (b_i - L_{i-1} y_{i-1}),
                                       Z_i = S_i^-1 Zt_i  with  Zt_a = L_left (the coupling of a to the left separator),
                                       Zt_i = -L_{i-1} Z_{i-1}                  => x_i = y_i - W_i x_{i+1} - Z_i x_left
           left part of the separator's row:   Dl = -L_{s-1} W_{s-1},  bl = -L_{s-1} y_{s-1},  Ll = -L_{s-1} Z_{s-1}
  backward recurrence (right to left)  Wh_i = -W_i Wh_{i+1},  Zh_i = Z_i - W_i Zh_{i+1},  yh_i = y_i - W_i yh_{i+1}
                                       (started with W, Z, y of s-1)             => x_a = yh_a - Wh_a x_s - Zh_a x_left
           right part of the LEFT separator's row:  Dr = -U_left Zh_a,  Ur = -U_left Wh_a,  br = -U_left yh_a
  reduced system over the separators   D~_k = D_s + Dl_k + Dr_{k+1},  U~_k = Ur_{k+1},  b~_k = b_s + bl_k + br_{k+1},  L~_{k-1} = Ll_k
  it is block tridiagonal again (explicit lower blocks): solved by a plain sweep, or by THIS algorithm one level up
  back-substitution of the interiors   x_i = y_i - W_i x_{i+1} - Z_i x_left,  i = s-1 .. a
Segments without interior (a == s) couple their separator directly: Ll = L_left, Ur = U_left, everything else zero.
"""
import numpy as np


def dense(D, U, L, b):
    n = len(D)
    A = np.zeros((9 * n, 9 * n))
    for i in range(n):
        A[9 * i:9 * i + 9, 9 * i:9 * i + 9] = D[i]
        if i + 1 < n:
            A[9 * i:9 * i + 9, 9 * i + 9:9 * i + 18] = U[i]
            A[9 * i + 9:9 * i + 18, 9 * i:9 * i + 9] = L[i]
    return A


def plain_sweep(D, U, L, b):
    """One chain: forward elimination + back-substitution (k_chain_forward3<false> + k_chain_backward, one-sided)."""
    n = len(D)
    W = np.zeros((n, 9, 9)); y = np.zeros((n, 9))
    for i in range(n):
        S = D[i] - (L[i - 1] @ W[i - 1] if i > 0 else 0.0)
        r = b[i] - (L[i - 1] @ y[i - 1] if i > 0 else 0.0)
        W[i] = np.linalg.solve(S, U[i]) if i + 1 < n else 0.0
        y[i] = np.linalg.solve(S, r)
    x = np.zeros((n, 9))
    for i in range(n - 1, -1, -1):
        x[i] = y[i] - (W[i] @ x[i + 1] if i + 1 < n else 0.0)
    return x


def even_segments(n, S):
    """[lo, hi) per segment as the library cuts them (csrc/batch.cu make_segments / batch.h plan_level2)."""
    return [((n * k) // S, (n * (k + 1)) // S) for k in range(S)]


def partitioned_solve(D, U, L, b, levels):
    """levels = [S1, S2, ...]: number of segments at each partition level; [] = plain sweep."""
    n = len(D)
    if not levels or n == 0:
        return plain_sweep(D, U, L, b)
    S = max(1, min(levels[0], n))
    segs = even_segments(n, S)
    W = np.zeros((n, 9, 9)); Z = np.zeros((n, 9, 9)); y = np.zeros((n, 9))
    Dl = np.zeros((S, 9, 9)); Ll = np.zeros((S, 9, 9)); bl = np.zeros((S, 9))
    Dr = np.zeros((S, 9, 9)); Ur = np.zeros((S, 9, 9)); br = np.zeros((S, 9))
    for k, (lo, hi) in enumerate(segs):
        a, s = lo, hi - 1
        left = lo - 1 if k > 0 else -1
        if s == a:                                     # no interior
            if left >= 0:
                Ll[k] = L[left]; Ur[k] = U[left]
            continue
        for i in range(a, s):                          # forward, with the spike block
            if i == a:
                Sm, r = D[i], b[i]
                Zt = L[left] if left >= 0 else np.zeros((9, 9))
            else:
                Sm, r, Zt = D[i] - L[i - 1] @ W[i - 1], b[i] - L[i - 1] @ y[i - 1], -L[i - 1] @ Z[i - 1]
            W[i], y[i], Z[i] = np.linalg.solve(Sm, U[i]), np.linalg.solve(Sm, r), np.linalg.solve(Sm, Zt)
        Dl[k], bl[k], Ll[k] = -L[s - 1] @ W[s - 1], -L[s - 1] @ y[s - 1], -L[s - 1] @ Z[s - 1]
        if left >= 0:                                  # backward recurrence
            Wh, Zh, yh = W[s - 1], Z[s - 1], y[s - 1]
            for i in range(s - 2, a - 1, -1):
                Wh, Zh, yh = -W[i] @ Wh, Z[i] - W[i] @ Zh, y[i] - W[i] @ yh
            Dr[k], Ur[k], br[k] = -U[left] @ Zh, -U[left] @ Wh, -U[left] @ yh
    seps = [hi - 1 for _, hi in segs]
    Dt = np.stack([D[s] + Dl[k] + (Dr[k + 1] if k + 1 < S else 0.0) for k, s in enumerate(seps)])
    Ut = np.stack([Ur[k + 1] if k + 1 < S else np.zeros((9, 9)) for k in range(S)])
    bt = np.stack([b[s] + bl[k] + (br[k + 1] if k + 1 < S else 0.0) for k, s in enumerate(seps)])
    Lt = np.stack([Ll[k + 1] if k + 1 < S else np.zeros((9, 9)) for k in range(S)])       # Lt[k] = A~(k+1, k)
    xs = partitioned_solve(Dt, Ut, Lt, bt, levels[1:])                                     # the next level (or a plain sweep)
    x = np.zeros((n, 9))
    for k, (lo, hi) in enumerate(segs):
        a, s = lo, hi - 1
        x[s] = xs[k]
        xl = xs[k - 1] if k > 0 else np.zeros(9)
        for i in range(s - 1, a - 1, -1):
            x[i] = y[i] - W[i] @ x[i + 1] - Z[i] @ xl
    return x
