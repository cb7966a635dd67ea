"""TEST INFRASTRUCTURE (oracle) - CPU restatement of stvo-pl's windowed greedy matcher `matchGrid` (points and lines
overloads) and of the `GridStructure` it queries, as pl-slam calls them at src/mapHandler.cpp:251-271 (points),
:382-418 (lines), :580-591, :686-706.

stvo-pl is NOT on disk: this restates SURVEY Appendix A.3 plus the published stvo-pl sources
from memory ([UPSTREAM-RECALL]) - "parity unpinned".  One deliberate definition: stvo-pl iterates the candidate set
through a std::unordered_set<int> (implementation-defined order, which decides ties between equally distant
candidates); here candidates are visited in ASCENDING index order.

Semantics restated:
  GridStructure(rows, cols): cells addressed (x, y), 0 <= x < cols, 0 <= y < rows; at(x, y) outside the grid refers to
    a bin that is never returned; get(x, y, w) = union of the cells [max(0, x - w.width.first), min(cols, x +
    w.width.second + 1)) x [max(0, y - w.height.first), min(rows, y + w.height.second + 1)).
  matchGrid: queries in index order; per query best / second-best over its candidates with strict '<' updates; with
    best_lr_matches a candidate i2 is considered for query i1 only if d(i1, i2) < the smallest distance any EARLIER
    query achieved on i2 (then that record and matches_21[i2] are updated); accept when
    best_d < best_d2 * nnr (f32 product, best_d2 = INT_MAX when there was a single candidate); finally, with
    best_lr_matches, drop i1 unless matches_21[matches_12[i1]] == i1.
  Lines: the query's cells are the Bresenham cells of its (integer) end points, each train line is registered in the
    Bresenham cells of its own end points, and a candidate is skipped (before its distance is computed) when
    |v . directions2[i2]| < line_sim_th, v = normalised direction of the query's end points (unguarded division: a
    zero-length query yields NaN, which fails the '<' and therefore passes the gate)."""
import numpy as np

from oracle.mapfeatures import hamming

INT_MAX = 2147483647


def bresenham(x1, y1, x2, y2):
    """Cells of the 8-connected line (x1,y1)-(x2,y2), end points included (stvo-pl lineIterator, used as a SET)."""
    x1, y1, x2, y2 = int(x1), int(y1), int(x2), int(y2)
    steep = abs(y2 - y1) > abs(x2 - x1)
    if steep:
        x1, y1, x2, y2 = y1, x1, y2, x2
    if x1 > x2:
        x1, x2, y1, y2 = x2, x1, y2, y1
    dx, dy = x2 - x1, abs(y2 - y1)
    err, ystep, y = dx // 2, (1 if y1 < y2 else -1), y1
    out = []
    for x in range(x1, x2 + 1):
        out.append((y, x) if steep else (x, y))
        err -= dy
        if err < 0:
            y += ystep
            err += dx
    return out


def _window(cx, cy, w, cols, rows):
    return max(0, cx - w[0]), min(cols, cx + w[1] + 1), max(0, cy - w[2]), min(rows, cy + w[3] + 1)


def _greedy(n1, n2, cand_fn, d1, d2, nnr, best_lr):
    m12 = np.full(n1, -1, np.int32)
    m21 = np.full(n2, -1, np.int64)
    dist = np.full(n2, INT_MAX, np.int64)
    matches = 0
    nnr = np.float32(nnr)
    for i1 in range(n1):
        best_d, best_d2, best_idx = INT_MAX, INT_MAX, -1
        for i2 in cand_fn(i1):                      # ascending index order (see header)
            d = hamming(d1[i1], d2[i2])
            if best_lr:
                if d < dist[i2]:
                    dist[i2] = d
                    m21[i2] = i1
                else:
                    continue
            if d < best_d:
                best_d2, best_d, best_idx = best_d, d, i2
            elif d < best_d2:
                best_d2 = d
        if np.float32(best_d) < np.float32(best_d2) * nnr:
            m12[i1] = best_idx
            matches += 1
    if best_lr:
        for i1 in range(n1):
            i2 = m12[i1]
            if i2 >= 0 and m21[i2] != i1:
                m12[i1] = -1
                matches -= 1
    return m12, matches


class Grid:
    """GridStructure: rows x cols lists of train indices; at(x, y).push_back(idx) outside the grid goes to a bin that is
    never returned."""

    def __init__(self, rows, cols):
        self.rows, self.cols = int(rows), int(cols)
        self.cells = [[[] for _ in range(self.rows)] for _ in range(self.cols)]

    def push(self, x, y, idx):
        x, y = int(x), int(y)
        if 0 <= x < self.cols and 0 <= y < self.rows:
            self.cells[x][y].append(int(idx))

    def get(self, x, y, w):
        x0, x1, y0, y1 = _window(int(x), int(y), w, self.cols, self.rows)
        out = set()
        for xx in range(x0, x1):
            for yy in range(y0, y1):
                out.update(self.cells[xx][yy])
        return out

    def csr(self):
        """(cell_start int32[cols*rows+1], cell_items int32[]) with cell (x, y) at index x*rows + y - the C ABI's form."""
        start, items = [0], []
        for x in range(self.cols):
            for y in range(self.rows):
                items += self.cells[x][y]
                start.append(len(items))
        return np.asarray(start, np.int32), np.asarray(items, np.int32)


def grid_from_points(t_cell, rows, cols):
    """src/mapHandler.cpp:260-264: every train point is pushed into the cell of its (scaled, truncated) position."""
    g = Grid(rows, cols)
    for idx, (x, y) in enumerate(np.asarray(t_cell, np.int64).reshape(-1, 2)):
        g.push(x, y, idx)
    return g


def grid_from_lines(t_line, rows, cols):
    """src/mapHandler.cpp:398-411: every train line is pushed into each cell of its getLineCoords() walk."""
    g = Grid(rows, cols)
    for idx, ln in enumerate(np.asarray(t_line, np.int64).reshape(-1, 4)):
        for (x, y) in bresenham(*ln):
            g.push(x, y, idx)
    return g


def match_grid_points(q_cell, d1, grid, d2, w, nnr, best_lr=True):
    """q_cell [n1,2] integer grid coordinates (x, y); grid: Grid over the train features; w = (width.first,
    width.second, height.first, height.second)."""
    q_cell = np.asarray(q_cell, np.int64).reshape(-1, 2)
    n1, n2 = len(q_cell), len(d2)

    def cand(i1):
        return sorted(i2 for i2 in grid.get(q_cell[i1, 0], q_cell[i1, 1], w) if 0 <= i2 < n2)

    return _greedy(n1, n2, cand, d1, d2, nnr, best_lr)


def match_grid_lines(q_line, d1, grid, t_dir, d2, w, nnr, line_sim_th, best_lr=True):
    """q_line [n1,4]: integer grid coordinates (x1, y1, x2, y2) of the projected query lines; grid: Grid over the train
    lines; t_dir [n2,2]: the train lines' unit directions (directions2)."""
    q_line = np.asarray(q_line, np.int64).reshape(-1, 4)
    t_dir = np.asarray(t_dir, np.float64).reshape(-1, 2)
    n1, n2 = len(q_line), len(d2)

    def cand(i1):
        c = set()
        for (x, y) in bresenham(*q_line[i1]):
            c |= grid.get(x, y, w)
        vx, vy = float(q_line[i1, 2] - q_line[i1, 0]), float(q_line[i1, 3] - q_line[i1, 1])
        with np.errstate(invalid="ignore", divide="ignore"):
            nrm = np.sqrt(np.float64(vx * vx + vy * vy))     # normalize(v) is unguarded: a query whose end points share
            vx, vy = np.float64(vx) / nrm, np.float64(vy) / nrm  # a cell gives 0/0 = NaN, and NaN < th is false -> kept
        out = []
        for i2 in sorted(c):
            if not (0 <= i2 < n2):
                continue
            with np.errstate(invalid="ignore"):
                if abs(vx * t_dir[i2, 0] + vy * t_dir[i2, 1]) < line_sim_th:
                    continue
            out.append(i2)
        return out

    return _greedy(n1, n2, cand, d1, d2, nnr, best_lr)
