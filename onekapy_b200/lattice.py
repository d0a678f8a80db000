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
    left, right, bottom, top = _grid_windows(torch, base, final, lo_x, hi_x, lo_y, hi_y)
    return torch.stack([left, right, bottom, top], dim=1).to(torch.int32).reshape(R, P, 4).contiguous()


def _grid_windows(torch, base, final, lo_x, hi_x, lo_y, hi_y):
    """Index ranges (in `final`) of the reference's grid after expand() to the box (lo_x, hi_x, lo_y, hi_y) [tensors of one
    shape; +-inf = nothing inserted yet = the base grid]: left, right, bottom, top (half-open) as float64 tensors."""
    i0, j0 = final.offset_of(base)                       # where the base grid sits inside the final lattice

    def cells_below(g0, c, d):                            # expansions so that g0 - k d < c   (probabilityfield.py:229-245)
        return torch.where(c <= g0, torch.floor((g0 - c) / d) + 1.0, torch.zeros_like(c))

    def cells_above(g1, c, d):                            # expansions so that g1 + k d > c
        return torch.where(c >= g1, torch.floor((c - g1) / d) + 1.0, torch.zeros_like(c))

    left = j0 - cells_below(base.xmin, lo_x, base.deltax)
    right = j0 + base.ncols + cells_above(base.xmax, hi_x, base.deltax)
    bottom = i0 - cells_below(base.ymin, lo_y, base.deltay)
    top = i0 + base.nrows + cells_above(base.ymax, hi_y, base.deltay)
    return left, right, bottom, top


def realization_boxes(torch, bb):
    """[R, P, 4] per-path boxes -> [R, 4] per-realization boxes (min x, max x, min y, max y)."""
    return torch.stack([bb[..., 0].amin(dim=1), bb[..., 1].amax(dim=1), bb[..., 2].amin(dim=1), bb[..., 3].amax(dim=1)], dim=1)


def union_before(torch, rb, prior=None):
    """[R, 4]: the union of the boxes of realizations 0..r-1 (and of `prior`, the box of everything other ranks inserted
    earlier) -- what the reference's grid had been expanded to when realization r began; +-inf where nothing precedes."""
    import math
    R = int(rb.shape[0])
    inf = float("inf")
    first = [inf, -inf, inf, -inf]
    if prior is not None and all(math.isfinite(v) for v in prior):
        first = [float(v) for v in prior]
    out = torch.empty_like(rb)
    out[0] = torch.tensor(first, dtype=rb.dtype, device=rb.device)
    if R > 1:
        out[1:, 0] = torch.clamp(torch.cummin(rb[:-1, 0], 0).values, max=first[0])
        out[1:, 1] = torch.clamp(torch.cummax(rb[:-1, 1], 0).values, min=first[1])
        out[1:, 2] = torch.clamp(torch.cummin(rb[:-1, 2], 0).values, max=first[2])
        out[1:, 3] = torch.clamp(torch.cummax(rb[:-1, 3], 0).values, min=first[3])
    return out


def affected_realizations(torch, base, final, rb, before, umbra):
    """bool [R]: could the reference's grid-at-that-moment have clipped ANY segment window of realization r?

    insert() needs, for a segment, the nodes floor((min - umbra - xmin)/delta) .. floor((max + umbra - xmin)/delta)
    (probabilityfield.py:298-301).  Every segment of realization r lies inside its box rb[r], and every grid the
    realization's paths met contains the grid as expanded to `before[r]` (the grid only grows).  So a realization whose box
    window, widened by ONE MORE cell on every side (the floors here are not the reference's operation-for-operation
    floors), sits inside that grid was rasterised exactly as the reference did even without any clip.  Conservative by
    construction: a realization wrongly called affected only costs time."""
    u = float(umbra)
    left, right, bottom, top = _grid_windows(torch, base, final, before[:, 0], before[:, 1], before[:, 2], before[:, 3])
    need_l = torch.floor((rb[:, 0] - u - final.xmin) / final.deltax) - 1.0
    need_r = torch.floor((rb[:, 1] + u - final.xmin) / final.deltax) + 2.0        # half-open upper end + one cell
    need_b = torch.floor((rb[:, 2] - u - final.ymin) / final.deltay) - 1.0
    need_t = torch.floor((rb[:, 3] + u - final.ymin) / final.deltay) + 2.0
    inside = (need_l >= left) & (need_r <= right) & (need_b >= bottom) & (need_t <= top)
    return ~inside                                               # (a nan box compares false everywhere -> affected)


def clip_windows_rows(torch, base, final, bb, before):
    """Per-path raster windows for SELECTED realizations: bb [n, P, 4] their per-path boxes, before [n, 4] the union of
    everything inserted before each of them (union_before).  Path p of such a realization met the grid as expanded to
    before U boxes of its paths 0..p.  Returns int32 [n, P, 4] = left, right, bottom, top (half-open, indices of `final`)."""
    n, P = int(bb.shape[0]), int(bb.shape[1])
    if n * P == 0:
        return torch.zeros((n, P, 4), dtype=torch.int32, device=bb.device)
    lo_x = torch.minimum(torch.cummin(bb[..., 0], 1).values, before[:, None, 0])
    hi_x = torch.maximum(torch.cummax(bb[..., 1], 1).values, before[:, None, 1])
    lo_y = torch.minimum(torch.cummin(bb[..., 2], 1).values, before[:, None, 2])
    hi_y = torch.maximum(torch.cummax(bb[..., 3], 1).values, before[:, None, 3])
    w = _grid_windows(torch, base, final, lo_x, hi_x, lo_y, hi_y)
    return torch.stack(w, dim=2).to(torch.int32).contiguous()
