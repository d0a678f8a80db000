"""Lattice geometry of the ProbabilityField (host side, integers and a handful of doubles).

The reference grid is node-centred, anchored on the target well and grows one whole cell at
a time by REPEATED addition/subtraction of the spacing (oneka/probabilityfield.py:140-146,
229-245).  LatticeGeom reproduces exactly that arithmetic, so that a lattice grown in one go
to a bounding box has the same xmin/xmax/ymin/ymax doubles the reference reaches after its
sequence of per-trace expansions.
"""
import math
from dataclasses import dataclass

from ._cabi import Lattice


@dataclass(frozen=True)
class LatticeGeom:
    deltax: float
    deltay: float
    xmin: float
    xmax: float
    ymin: float
    ymax: float
    nrows: int
    ncols: int

    # -- constructors ---------------------------------------------------------------------------
    @staticmethod
    def anchored(deltax, deltay, xo, yo):
        """3 x 3 nodes centred on (xo, yo): probabilityfield.py:140-146."""
        if deltax <= 0 or deltay <= 0:
            raise ValueError("<deltax>, <deltay> must be > 0.")
        return LatticeGeom(deltax, deltay, xo - deltax, xo + deltax, yo - deltay, yo + deltay, 3, 3)

    @staticmethod
    def of_field(pf):
        return LatticeGeom(pf.deltax, pf.deltay, pf.xmin, pf.xmax, pf.ymin, pf.ymax, int(pf.nrows), int(pf.ncols))

    # -- probabilityfield.py:229-245 ------------------------------------------------------------------
    def expanded(self, xmin, xmax, ymin, ymax):
        """Grow by whole cells until the box is STRICTLY inside; returns the new geometry."""
        return self.expanded_with_shift(xmin, xmax, ymin, ymax)[0]

    def expanded_with_shift(self, xmin, xmax, ymin, ymax):
        """-> (geometry, rshift, cshift): rows/columns added below/left (probabilityfield.py:225-245)."""
        if xmin > xmax or ymin > ymax:
            raise ValueError("min must be <= max")
        if not all(math.isfinite(v) for v in (xmin, xmax, ymin, ymax)):
            raise ValueError("non-finite bounding box")
        x0, x1, y0, y1 = self.xmin, self.xmax, self.ymin, self.ymax
        nrows, ncols = self.nrows, self.ncols
        rshift = cshift = 0
        while xmin <= x0:
            x0 -= self.deltax
            ncols += 1
            cshift += 1
        while xmax >= x1:
            x1 += self.deltax
            ncols += 1
        while ymin <= y0:
            y0 -= self.deltay
            nrows += 1
            rshift += 1
        while ymax >= y1:
            y1 += self.deltay
            nrows += 1
        return LatticeGeom(self.deltax, self.deltay, x0, x1, y0, y1, nrows, ncols), rshift, cshift

    # -- queries ------------------------------------------------------------------------------------
    def strictly_contains(self, bbox):
        """True when expand(bbox) would be a no-op (every vertex strictly inside the outer nodes)."""
        return (bbox[0] > self.xmin) and (bbox[1] < self.xmax) and (bbox[2] > self.ymin) and (bbox[3] < self.ymax)

    def offset_of(self, inner):
        """(row, col) index of `inner`'s node (0, 0) in this lattice (same anchor, same spacing)."""
        j0 = int(round((inner.xmin - self.xmin) / self.deltax))
        i0 = int(round((inner.ymin - self.ymin) / self.deltay))
        if i0 < 0 or j0 < 0 or i0 + inner.nrows > self.nrows or j0 + inner.ncols > self.ncols:
            raise ValueError("inner lattice is not contained in this one")
        return i0, j0

    def as_lattice(self, umbra) -> Lattice:
        return Lattice(xmin=float(self.xmin), ymin=float(self.ymin), deltax=float(self.deltax),
                       deltay=float(self.deltay), nrows=int(self.nrows), ncols=int(self.ncols), umbra=float(umbra))


def final_geometry(deltax, deltay, xo, yo, bbox):
    """The extents the reference's auto-expanding field ends with when the union of all trace
    bounding boxes is `bbox` (successive expansions commute with one expansion to the union)."""
    return LatticeGeom.anchored(deltax, deltay, xo, yo).expanded(*bbox)


def clip_windows(torch, base, final, bb, prior=None):
    """Per-path raster windows of the reference's auto-expanding grid, as lattice index ranges of `final`.

    `base` is the grid before the first path (3 x 3 on the target for a fresh field); path n is inserted
    after the grid has been expanded to the union of the bounding boxes of paths 0..n (rasterize(),
    probabilityfield.py:335, in (realization, path) order) -- a running min/max (torch.cummin/cummax, on
    whatever device `bb` lives).  expand() moves xmin down by whole cells until it is strictly below the
    box (:229-245): k = floor((xmin0 - c)/delta) + 1 cells when c <= xmin0, else 0; likewise upwards.
    `prior` = box of everything inserted earlier (other ranks' shards).
    bb: float64 tensor [R, P, 4] = min x, max x, min y, max y.  Returns int32 [R, P, 4] = left, right,
    bottom, top (half-open)."""
    import math
    R, P = int(bb.shape[0]), int(bb.shape[1])
    if R * P == 0:
        return torch.zeros((R, P, 4), dtype=torch.int32, device=bb.device)
    f = bb.reshape(-1, 4)
    lo_x = torch.cummin(f[:, 0], 0).values
    hi_x = torch.cummax(f[:, 1], 0).values
    lo_y = torch.cummin(f[:, 2], 0).values
    hi_y = torch.cummax(f[:, 3], 0).values
    if prior is not None and all(math.isfinite(v) for v in prior):
        lo_x = torch.clamp(lo_x, max=float(prior[0]))
        hi_x = torch.clamp(hi_x, min=float(prior[1]))
        lo_y = torch.clamp(lo_y, max=float(prior[2]))
        hi_y = torch.clamp(hi_y, min=float(prior[3]))
    i0, j0 = final.offset_of(base)                       # where the base grid sits inside the final lattice

    def cells_below(g0, c, d):                            # expansions so that g0 - k d < c
        return torch.where(c <= g0, torch.floor((g0 - c) / d) + 1.0, torch.zeros_like(c))

    def cells_above(g1, c, d):                            # expansions so that g1 + k d > c
        return torch.where(c >= g1, torch.floor((c - g1) / d) + 1.0, torch.zeros_like(c))

    left = j0 - cells_below(base.xmin, lo_x, base.deltax)
    right = j0 + base.ncols + cells_above(base.xmax, hi_x, base.deltax)
    bottom = i0 - cells_below(base.ymin, lo_y, base.deltay)
    top = i0 + base.nrows + cells_above(base.ymax, hi_y, base.deltay)
    return torch.stack([left, right, bottom, top], dim=1).to(torch.int32).reshape(R, P, 4).contiguous()


def affected_paths(torch, final, bb, clip, umbra):
    """bool [R, P]: could the reference's grid-at-that-moment have clipped ANY segment window of the path?

    insert() needs, for a segment, the nodes floor((min - umbra - xmin)/delta) .. floor((max + umbra - xmin)/delta)
    (probabilityfield.py:298-301); every segment of a path lies inside the path's bounding box, so its windows lie
    inside the box's window.  A path whose box window, widened by ONE MORE cell on every side (the floors here are not
    the reference's operation-for-operation floors), sits inside its clip window was rasterised exactly as the
    reference did even without the clip.  Conservative by construction: a path wrongly called affected only costs time.
    bb: float64 [R, P, 4] (min x, max x, min y, max y); clip: int32 [R, P, 4] from clip_windows (indices of `final`)."""
    u = float(umbra)
    c = clip.to(torch.float64)
    need_l = torch.floor((bb[..., 0] - u - final.xmin) / final.deltax) - 1.0
    need_r = torch.floor((bb[..., 1] + u - final.xmin) / final.deltax) + 2.0        # half-open upper end + one cell
    need_b = torch.floor((bb[..., 2] - u - final.ymin) / final.deltay) - 1.0
    need_t = torch.floor((bb[..., 3] + u - final.ymin) / final.deltay) + 2.0
    inside = (need_l >= c[..., 0]) & (need_r <= c[..., 1]) & (need_b >= c[..., 2]) & (need_t <= c[..., 3])
    return ~inside                                               # (a nan box compares false everywhere -> affected)
