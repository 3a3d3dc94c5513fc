"""The algebra of the partitioned (one- and two-level) block-tridiagonal solve the kernels implement
(oracle/partition_oracle.py restates csrc/kernels_chain.cu) against a dense LAPACK solve, on the CPU: symmetric-coupled
systems like the BA normal equations and non-symmetric ones like the reduced systems, every segmentation shape the library can
produce (segments without interior, one segment, as many segments as elements)."""
import numpy as np
import pytest

import partition_oracle as po


def _system(n, seed, symmetric_coupling):
    rng = np.random.default_rng(seed)
    U = rng.normal(size=(n, 9, 9)) * 0.3
    L = np.stack([u.T for u in U]) if symmetric_coupling else rng.normal(size=(n, 9, 9)) * 0.3
    D = rng.normal(size=(n, 9, 9)) * 0.2 + 3.0 * np.eye(9)          # block diagonally dominant, non-symmetric like the BA blocks
    b = rng.normal(size=(n, 9))
    U[-1] = 0.0; L[-1] = 0.0
    return D, U, L, b


@pytest.mark.parametrize("symmetric", [True, False])
@pytest.mark.parametrize("n,levels", [(1, [1]), (2, [2]), (7, [3]), (40, [6]), (40, [40]), (41, [13, 3]), (200, [50, 7]),
                                       (200, [67, 8, 2]), (64, [64, 8]), (30, [30, 30]), (97, [1]), (97, [])])
def test_partitioned_solve_equals_dense_solve(n, levels, symmetric):
    D, U, L, b = _system(n, 1000 + n + len(levels), symmetric)
    x = po.partitioned_solve(D, U, L, b, levels)
    ref = np.linalg.solve(po.dense(D, U, L, b), b.reshape(-1)).reshape(n, 9)
    assert np.abs(x - ref).max() < 1e-11 * max(1.0, np.abs(ref).max()), (n, levels)


def test_segments_tile_the_chain_like_the_library():
    for n, S in [(10, 3), (2_400_000, 3552), (3552, 60), (5, 5), (9, 1)]:
        segs = po.even_segments(n, S)
        assert segs[0][0] == 0 and segs[-1][1] == n and all(a[1] == b[0] for a, b in zip(segs, segs[1:]))
        assert all(hi > lo for lo, hi in segs)
    # sequential depth of a 2.4 M-frame arc: one level sqrt(T)+sqrt(T) eliminations, two levels len1 + len2 + S2
    T = 2_400_000
    one = 2 * int(T ** 0.5)
    two = -(-T // 3552) + -(-3552 // 60) + 60
    assert one == 3098 and two == 796
