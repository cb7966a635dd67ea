"""Round-2 groundwork (test infrastructure, CPU only): functional simulation of SPECULATIVE PARALLEL region growing with
in-order commit, checked for exactness against the sequential LSD region growing on the same data.

Sequential LSD (what k_lsd_grow reproduces): seeds in pseudo-order; an unused seed grows a region over unused, aligned
neighbours (the region angle is updated by every acceptance); the region's pixels become used.

Speculative scheme (a reorder buffer over regions): M workers each take the next seed (in order) that is unused in the
COMMITTED map and grow it against that live map, keeping their acceptances private.  Regions commit strictly in seed
order.  At commit time region k is
  * dropped        if its seed has meanwhile been committed by an earlier region (the sequential loop would skip it);
  * re-executed    if it was aborted or any pixel it accepted has meanwhile been committed by an earlier region (then
                   every earlier region is committed, so the re-execution against the live map is the sequential one);
  * committed      otherwise - its trace is the sequential trace: pixels it examined and rejected as unaligned are
                   rejected whatever their used state, pixels it saw as used were committed by earlier regions only.
A worker aborts early when it accepts the SEED of an earlier, still unresolved region (that pixel will belong to an
earlier region whatever happens), which kills the duplicates started on the same edge after one step.

Usage: python oracle/tools/spec_grow_sim.py [--workers 8,16,32,64]
Prints, per window size, the makespan in region-point steps against the sequential chain and the waste / retry counts.
"""
import argparse
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "pl-slam_b200"))

NOTDEF = -1024.0
PREC = math.pi * 22.5 / 180.0
NB = [(-1, -1), (0, -1), (1, -1), (-1, 0), (0, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]   # (dx, dy), OpenCV's loop order


def level_lines(img, scale=1.2, quant=2.0, n_bins=1024):
    import cv2
    sigma = 0.6 / scale if scale < 1 else 0.6
    k = int(math.ceil(sigma * math.sqrt(2 * 3 * math.log(10.0)))) * 2 + 1
    g = cv2.GaussianBlur(img, (k, k), sigma)
    g = cv2.resize(g, (int(round(img.shape[1] * scale)), int(round(img.shape[0] * scale))), interpolation=cv2.INTER_LINEAR_EXACT)
    a = g.astype(np.float64)
    H, W = a.shape
    DA = a[1:, 1:] - a[:-1, :-1]; BC = a[:-1, 1:] - a[1:, :-1]
    gx, gy = DA + BC, DA - BC
    norm = np.sqrt((gx * gx + gy * gy) / 4.0)
    rho = quant / math.sin(PREC)
    ang = np.full((H, W), NOTDEF)
    mag = np.zeros((H, W))
    inner = np.arctan2(gx, -gy)
    ang[:-1, :-1] = np.where(norm <= rho, NOTDEF, inner)
    mag[:-1, :-1] = norm
    mx = mag.max()
    ys, xs = np.nonzero(ang != NOTDEF)
    bins = (mag[ys, xs] * ((n_bins - 1) / mx)).astype(np.int64)
    order = np.lexsort((xs, ys, -bins))                # bins descending, raster order inside a bin
    return ang, list(zip(xs[order].tolist(), ys[order].tolist()))


def aligned(a, theta):
    if a == NOTDEF:
        return False
    d = abs(theta - a)
    if d > 1.5 * math.pi:
        d = abs(d - 2 * math.pi)
    return d <= PREC


class Region:
    __slots__ = ("k", "seed", "pts", "mine", "r", "theta", "sx", "sy", "aborted", "steps")

    def __init__(self, k, seed, ang):
        self.k, self.seed = k, seed
        self.pts = [seed]; self.mine = {seed}; self.r = 0
        self.theta = ang[seed[1]][seed[0]]
        self.sx, self.sy = math.cos(self.theta), math.sin(self.theta)
        self.aborted = False; self.steps = 0

    def done(self):
        return self.aborted or self.r >= len(self.pts)

    def step(self, ang, used, W, H, seed_owner=None):
        """Processes one region point (its 8 neighbours in order)."""
        x, y = self.pts[self.r]; self.r += 1; self.steps += 1
        for dx, dy in NB:
            xx, yy = x + dx, y + dy
            if xx < 0 or yy < 0 or xx >= W or yy >= H:
                continue
            q = (xx, yy)
            if used[yy][xx] or q in self.mine:
                continue
            a = ang[yy][xx]
            if aligned(a, self.theta):
                if seed_owner is not None:
                    j = seed_owner.get(q)
                    if j is not None and j < self.k:       # the seed of an earlier unresolved region: abort
                        self.aborted = True
                        return
                self.mine.add(q); self.pts.append(q)
                self.sx += math.cos(a); self.sy += math.sin(a)
                self.theta = math.atan2(self.sy, self.sx)


def sequential(ang, order):
    H, W = len(ang), len(ang[0])
    used = [[False] * W for _ in range(H)]
    out, steps = [], 0
    for k, s in enumerate(order):
        if used[s[1]][s[0]]:
            continue
        R = Region(k, s, ang)
        while not R.done():
            R.step(ang, used, W, H)
        for (x, y) in R.pts:
            used[y][x] = True
        out.append((k, R.pts)); steps += R.steps
    return out, steps


def speculative(ang, order, M, R=None):
    """M workers, a reorder buffer of R >= M regions (a worker that finishes a region leaves it in the buffer until it
    can commit and starts the next one)."""
    R = R or M
    H, W = len(ang), len(ang[0])
    used = [[False] * W for _ in range(H)]
    rob, out = [], []                      # rob entries: [region, non_speculative]
    nxt, ticks, work, retries, dropped = 0, 0, 0, 0, 0
    seed_owner = {}
    while True:
        # dispatch: fill the buffer with the next seeds that are unused in the committed map
        while len(rob) < R and nxt < len(order):
            s = order[nxt]
            if not used[s[1]][s[0]]:
                rob.append([Region(nxt, s, ang), False]); seed_owner[s] = nxt
            nxt += 1
        if not rob:
            break
        ticks += 1
        busy = 0
        for ent in rob:                    # the M oldest unfinished regions run this tick
            Rg = ent[0]
            if not Rg.done():
                Rg.step(ang, used, W, H, None if ent[1] else seed_owner); work += 1
                busy += 1
                if busy == M:
                    break
        # in-order commit
        while rob and rob[0][0].done():
            Rg, exact = rob[0]
            if used[Rg.seed[1]][Rg.seed[0]]:
                rob.pop(0); seed_owner.pop(Rg.seed, None); dropped += 1
                continue
            if not exact and (Rg.aborted or any(used[y][x] for (x, y) in Rg.pts)):
                retries += 1
                rob[0] = [Region(Rg.k, Rg.seed, ang), True]     # every earlier region is committed: this run is exact
                break
            for (x, y) in Rg.pts:
                used[y][x] = True
            out.append((Rg.k, Rg.pts)); rob.pop(0); seed_owner.pop(Rg.seed, None)
    return out, dict(ticks=ticks, work=work, retries=retries, dropped=dropped)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", default="4,8,16,32,64")
    ap.add_argument("--rob", type=int, default=1, help="reorder-buffer entries per worker")
    ap.add_argument("--scale", type=float, default=1.2)
    ap.add_argument("--crop", default="", help="w,h crop of the bench frame (the pure-Python simulation is slow)")
    args = ap.parse_args()
    from plslam_b200 import synth
    import plslam_b200 as plf
    world = synth.World(seed=7)
    L = next(iter(synth.stream(plf.KITTI_CAMERA, 1, world=world, seed=42)))[0]
    if args.crop:
        w, h = (int(v) for v in args.crop.split(","))
        L = np.ascontiguousarray(L[:h, :w])
    ang, order = level_lines(L, args.scale)
    ang = ang.tolist()
    ref, seq_steps = sequential(ang, order)
    sizes = sorted((len(p) for _, p in ref), reverse=True)
    print(f"image {L.shape[1]}x{L.shape[0]}: {len(order)} seeds, {len(ref)} regions, sequential chain {seq_steps} region points; largest regions {sizes[:5]}")
    for M in (int(v) for v in args.workers.split(",")):
        out, st = speculative(ang, order, M, M * args.rob)
        ok = out == ref
        print(f"M={M:3d} (buffer {M * args.rob}): identical={ok}  makespan {st['ticks']} steps (x{seq_steps / st['ticks']:.2f} vs sequential), work {st['work']} (x{st['work'] / seq_steps:.2f}), "
              f"re-executions {st['retries']}, dropped duplicates {st['dropped']}")
        assert ok, "speculative result differs from the sequential one"


if __name__ == "__main__":
    main()
