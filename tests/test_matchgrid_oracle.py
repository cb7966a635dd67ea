"""CPU checks of the matchGrid oracle (oracle/matchgrid.py: stvo-pl matchGrid + GridStructure, [UPSTREAM-RECALL])."""
import numpy as np

from oracle import matchgrid as mg


def _d(*rows):
    out = np.zeros((len(rows), 32), np.uint8)
    for r, nbits in enumerate(rows):
        for b in range(nbits):
            out[r, b // 8] |= 1 << (b % 8)
    return out


W1 = (1, 1, 1, 1)


def test_bresenham_is_8_connected_and_includes_end_points():
    for (a, b, c, d) in [(0, 0, 5, 2), (3, 7, 1, 1), (4, 4, 4, 4), (0, 5, 9, 5), (2, 0, 2, 6)]:
        cells = mg.bresenham(a, b, c, d)
        assert (a, b) in cells and (c, d) in cells and len(cells) == max(abs(c - a), abs(d - b)) + 1
        for (x0, y0), (x1, y1) in zip(cells, cells[1:]):
            assert max(abs(x1 - x0), abs(y1 - y0)) == 1


def test_window_restricts_candidates_and_single_candidate_is_accepted():
    # query in cell (5,5); train 0 three cells away (outside the +-1 window), train 1 adjacent
    m, n = mg.match_grid_points([[5, 5]], _d(10), mg.grid_from_points([[8, 5], [6, 5]], 48, 64), _d(10, 60), W1, 0.9, True)
    assert list(m) == [1] and n == 1          # best_d2 stays INT_MAX -> ratio test passes with one candidate


def test_out_of_grid_train_cell_is_never_a_candidate():
    m, n = mg.match_grid_points([[0, 0]], _d(10), mg.grid_from_points([[-1, 0], [0, 0]], 48, 64), _d(10, 200), W1, 0.9, True)
    assert list(m) == [1]


def test_order_dependence_with_best_lr():
    # both queries see both train rows; query 0 takes train 0 at distance 4 (and records 30 on train 1);
    # for query 1, train 0 (d = 6) does not beat query 0's 4 -> skipped; train 1 (d = 20 < 30) is its only candidate
    d1 = _d(4, 6); d2 = _d(0, 26)
    d2[1] = _d(26)[0]
    m, n = mg.match_grid_points([[5, 5], [5, 5]], d1, mg.grid_from_points([[5, 5], [5, 6]], 48, 64), d2, W1, 0.9, True)
    assert list(m) == [0, 1] and n == 2
    # without the mutual bookkeeping both queries pick train 0 (4 < 0.9*22 and 6 < 0.9*20)
    m, n = mg.match_grid_points([[5, 5], [5, 5]], d1, mg.grid_from_points([[5, 5], [5, 6]], 48, 64), d2, W1, 0.9, False)
    assert list(m) == [0, 0] and n == 2


def test_lines_direction_gate_and_zero_length_query():
    d1 = _d(5, 5); d2 = _d(5, 90)
    t_line = [[2, 2, 8, 2], [2, 3, 2, 9]]                 # horizontal, vertical
    t_dir = [[1.0, 0.0], [0.0, 1.0]]
    q_line = [[2, 2, 9, 2], [2, 2, 2, 2]]                 # horizontal query; zero-length query (NaN direction)
    m, n = mg.match_grid_lines(q_line, d1, mg.grid_from_lines(t_line, 48, 64), t_dir, d2, W1, 0.9, 0.75, False)
    # query 0: the vertical train line is gated out (|cos| = 0 < 0.75) -> single candidate 0
    # query 1: NaN fails the '<' -> both stay candidates: best 0 (d 0), second 85 -> accepted
    assert list(m) == [0, 0] and n == 2


def test_grid_csr_layout():
    g = mg.grid_from_points([[0, 0], [0, 1], [63, 47], [0, 1]], 48, 64)
    start, items = g.csr()
    assert len(start) == 64 * 48 + 1 and list(items) == [0, 1, 3, 2]
    assert start[0] == 0 and start[1] == 1 and start[2] == 3 and start[-1] == 4 and start[35 * 48 + 47] == 3
