"""GPU parity: plf_match_grid_points / plf_match_grid_lines vs oracle/matchgrid.py (bit-exact match vectors)."""
import numpy as np
import pytest

import plslam_b200 as plf
from oracle import matchgrid as mg

pytestmark = pytest.mark.gpu
COLS, ROWS = 64, 48   # stvo-pl GRID_COLS x GRID_ROWS


def _descs(rng, n, protos=None, flip=0.1):
    if protos is None:
        return rng.integers(0, 256, (n, 32), dtype=np.uint8)
    flips = rng.random((n, 256)) < flip
    return (protos[rng.integers(0, len(protos), n)] ^ np.packbits(flips, axis=1)).astype(np.uint8)


@pytest.mark.parametrize("best_lr", [True, False])
@pytest.mark.parametrize("seed,n1,n2,ws", [(1, 300, 320, 1), (2, 700, 650, 2), (3, 40, 900, 0), (4, 500, 30, 3)])
def test_points_vs_oracle(fe, seed, n1, n2, ws, best_lr):
    rng = np.random.default_rng(seed)
    protos = rng.integers(0, 256, (60, 32), dtype=np.uint8)       # shared prototypes -> realistic near-duplicates and ties
    d1, d2 = _descs(rng, n1, protos), _descs(rng, n2, protos)
    q = np.stack([rng.integers(-2, COLS + 2, n1), rng.integers(-2, ROWS + 2, n1)], 1)
    t = np.stack([rng.integers(-1, COLS + 1, n2), rng.integers(-1, ROWS + 1, n2)], 1)
    w = (ws, ws, ws, ws)
    grid = mg.grid_from_points(t, ROWS, COLS)
    grid.push(3, 3, n2 + 5); grid.push(4, 4, -1)                  # stray indices are ignored (i2 < 0 || i2 >= desc2.rows)
    ref, nref = mg.match_grid_points(q, d1, grid, d2, w, 0.9, best_lr)
    got, ngot = fe.match_grid_points(q, d1, *grid.csr(), d2, COLS, ROWS, w, 0.9, best_lr)
    assert ngot == nref and np.array_equal(got, ref)
    assert nref > 0


@pytest.mark.parametrize("best_lr", [True, False])
def test_lines_vs_oracle(fe, best_lr):
    rng = np.random.default_rng(11)
    n1, n2 = 220, 260
    protos = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    d1, d2 = _descs(rng, n1, protos), _descs(rng, n2, protos)

    def lines(n):
        a = np.stack([rng.integers(-2, COLS + 2, n), rng.integers(-2, ROWS + 2, n)], 1)
        b = a + rng.integers(-9, 10, (n, 2))
        return np.concatenate([a, b], 1)

    ql, tl = lines(n1), lines(n2)
    ql[:10, 2:] = ql[:10, :2]                                      # zero-length queries (NaN direction)
    td = (tl[:, 2:] - tl[:, :2]).astype(np.float64)
    nrm = np.linalg.norm(td, axis=1, keepdims=True)
    td = np.where(nrm > 0, td / np.maximum(nrm, 1e-300), np.array([[1.0, 0.0]]))
    w = (1, 1, 1, 1)
    grid = mg.grid_from_lines(tl, ROWS, COLS)
    ref, nref = mg.match_grid_lines(ql, d1, grid, td, d2, w, 0.9, 0.75, best_lr)
    got, ngot = fe.match_grid_lines(ql, d1, *grid.csr(), td, d2, COLS, ROWS, w, 0.9, 0.75, best_lr)
    assert ngot == nref and np.array_equal(got, ref)
    assert nref > 0


def test_edge_cases(fe):
    rng = np.random.default_rng(0)
    d = rng.integers(0, 256, (5, 32), dtype=np.uint8)
    g5 = mg.grid_from_points(np.zeros((5, 2)), ROWS, COLS).csr()
    g0 = mg.Grid(ROWS, COLS).csr()
    m, n = fe.match_grid_points(np.zeros((0, 2)), d[:0], *g5, d, COLS, ROWS, (1, 1, 1, 1), 0.9)
    assert len(m) == 0 and n == 0
    m, n = fe.match_grid_points(np.zeros((5, 2)), d, *g0, d[:0], COLS, ROWS, (1, 1, 1, 1), 0.9)
    assert list(m) == [-1] * 5 and n == 0
    m, n = fe.match_grid_points(np.zeros((5, 2)), d, *g0, d, COLS, ROWS, (1, 1, 1, 1), 0.9)     # empty grid: no candidates
    assert list(m) == [-1] * 5 and n == 0
    with pytest.raises(plf.PlfError, match="bad arguments"):
        fe.match_grid_points(np.zeros((5, 2)), d, *g5, d, 0, ROWS, (1, 1, 1, 1), 0.9)
