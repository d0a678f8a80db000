"""ProbabilityField: drop-in for oneka/probabilityfield.py.

Same attributes (deltax, deltay, nrows, ncols, xmin, xmax, ymin, ymax, total_weight,
pgrid float64[nrows, ncols] with row = y / col = x, rgrid bool) and methods.  Geometry
bookkeeping (expand) is host arithmetic identical to the reference's; the vector-to-raster
work of insert()/rasterize() -- the reference's hot loops 4 and 5 -- runs on the GPU through
oneka_raster_traces (the same device rasteriser the fused capture kernel uses).
"""
import math

import numpy as np

from .model import RangeError
from ..lattice import LatticeGeom


class ProbabilityField:
    # Pickles written by the reference's archive (oneka/archive.py:46-89: a bz2-compressed pickle of a dict whose 'pfield'
    # entry is this object) name the class `oneka.probabilityfield.ProbabilityField`; the drop-in module `oneka.probabilityfield`
    # re-exports this class, so archives written here load in the reference and vice versa (tests/test_archive_roundtrip.py).
    __module__ = "oneka.probabilityfield"

    def __init__(self, deltax, deltay, xo=np.nan, yo=np.nan):
        """oneka/probabilityfield.py:125-151."""
        if deltax <= 0:
            raise RangeError("<deltax> must be > 0.")
        if deltay <= 0:
            raise RangeError("<deltay> must be > 0.")
        self.deltax = deltax
        self.deltay = deltay
        if np.isnan(xo) or np.isnan(yo):
            self.nrows = 0
            self.ncols = 0
        else:
            self.xmin = xo - deltax
            self.xmax = xo + deltax
            self.ymin = yo - deltay
            self.ymax = yo + deltay
            self.nrows = 3
            self.ncols = 3
            self.pgrid = np.zeros((self.nrows, self.ncols), dtype=float)
            self.rgrid = np.zeros((self.nrows, self.ncols), dtype=bool)
            self.total_weight = 0.0

    def __repr__(self):
        if self.nrows == 0:
            return "ProbabilityField({0.deltax}, {0.deltay})".format(self)
        return ("ProbabilityField({0.deltax}, {0.deltay}, {0.nrows}, {0.ncols}, "
                "{0.xmin}, {0.xmax}, {0.ymin}, {0.ymax}, {0.total_weight})".format(self))

    def __str__(self):
        if self.nrows == 0:
            return "ProbabilityField({0.deltax}, {0.deltay}, {0.nrows}, {0.ncols})".format(self)
        return ("ProbabilityField({0.deltax}, {0.deltay}, {0.nrows}, {0.ncols}, "
                "{0.xmin}, {0.xmax}, {0.ymin}, {0.ymax}, {0.total_weight}, "
                "\n{0.pgrid!r}, \n{0.rgrid!r})".format(self))

    # -------------------------------------------------------------------------------------------
    @classmethod
    def from_counts(cls, geom, counts, total_weight, weight=1.0):
        """Materialise the field the reference would hold after `total_weight/weight` register()
        calls: pgrid = weight * (number of realizations that marked the node)."""
        pf = cls(geom.deltax, geom.deltay)
        pf.xmin, pf.xmax, pf.ymin, pf.ymax = geom.xmin, geom.xmax, geom.ymin, geom.ymax
        pf.nrows, pf.ncols = int(geom.nrows), int(geom.ncols)
        pf.pgrid = np.asarray(counts).astype(float)
        if weight != 1.0:
            pf.pgrid *= weight
        pf.rgrid = np.zeros((pf.nrows, pf.ncols), dtype=bool)
        pf.total_weight = float(total_weight)
        return pf

    def geometry(self):
        return LatticeGeom.of_field(self)

    # -------------------------------------------------------------------------------------------
    def expand(self, xmin, xmax, ymin, ymax):
        """Grow the grids so that the box is strictly inside (oneka/probabilityfield.py:175-261)."""
        if xmin > xmax:
            raise RangeError("<xmin> must be <= <xmax>.")
        if ymin > ymax:
            raise RangeError("<ymin> must be <= <ymax>.")
        if (self.ncols == 0) or (self.nrows == 0):
            # empty field: (xmin, ymin) lands in index [1, 1]   (:205-220)
            self.xmin = xmin - self.deltax
            self.ymin = ymin - self.deltay
            self.ncols = max(3, math.ceil((xmax - self.xmin) / self.deltax) + 2)
            self.nrows = max(3, math.ceil((ymax - self.ymin) / self.deltay) + 2)
            self.xmax = self.xmin + (self.ncols - 1) * self.deltax
            self.ymax = self.ymin + (self.nrows - 1) * self.deltay
            self.pgrid = np.zeros((self.nrows, self.ncols), dtype=float)
            self.rgrid = np.zeros((self.nrows, self.ncols), dtype=bool)
            self.total_weight = 0.0
            return
        g, rshift, cshift = LatticeGeom.of_field(self).expanded_with_shift(xmin, xmax, ymin, ymax)
        self.xmin, self.xmax, self.ymin, self.ymax = g.xmin, g.xmax, g.ymin, g.ymax
        if (g.nrows != self.nrows) or (g.ncols != self.ncols):
            pgrid = np.zeros((g.nrows, g.ncols), dtype=float)
            rgrid = np.zeros((g.nrows, g.ncols), dtype=bool)
            rgrid[rshift:rshift + self.nrows, cshift:cshift + self.ncols] = self.rgrid
            pgrid[rshift:rshift + self.nrows, cshift:cshift + self.ncols] = self.pgrid
            self.rgrid = rgrid
            self.pgrid = pgrid
            self.nrows = g.nrows
            self.ncols = g.ncols

    # -------------------------------------------------------------------------------------------
    def _raster(self, tracks, umbra):
        from ..engine import default_engine
        if self.nrows == 0 or self.ncols == 0:
            return
        c = default_engine().raster_traces(self.geometry(), umbra, tracks)
        self.rgrid |= (c != 0)

    def insert(self, ax, ay, bx, by, umbra):
        """Mark the nodes within umbra of segment [(ax, ay), (bx, by)], clipped to the current grid
        (oneka/probabilityfield.py:264-310)."""
        self._raster([np.array([[ax, ay], [bx, by]], dtype=float)], umbra)

    def rasterize(self, x, y, umbra):
        """expand() to the track's bounding box, then insert() every segment
        (oneka/probabilityfield.py:313-339)."""
        self.expand(min(x), max(x), min(y), max(y))
        self._raster([np.stack([np.asarray(x, dtype=float), np.asarray(y, dtype=float)], axis=1)], umbra)

    def insert_tracks(self, tracks, umbra):
        """Extension: insert() every segment of many tracks in one launch (no expand)."""
        self._raster(tracks, umbra)

    def register(self, weight):
        """pgrid += weight where rgrid; clear rgrid (oneka/probabilityfield.py:342-359)."""
        self.total_weight += weight
        self.pgrid[self.rgrid] += weight
        self.rgrid[:] = False

    def reset(self):
        """Discard the current realization's registration (oneka/probabilityfield.py:362-376)."""
        self.rgrid[:] = False

    @staticmethod
    def distancesquared(ax, ay, bx, by, cx, cy):
        """Distance squared from c to the SEGMENT [a, b] (oneka/probabilityfield.py:379-427).

        Scalar utility kept for API compatibility: the device rasteriser carries its own copy of
        this formula (exact_distancesquared in csrc/oneka_device.cuh); this one is plain Python."""
        bax = bx - ax
        bay = by - ay
        cax = cx - ax
        cay = cy - ay
        perpdot = bax * cay - bay * cax
        dot = bax * cax + bay * cay
        length2 = bax * bax + bay * bay
        with np.errstate(all="ignore"):
            alpha2 = np.float64(perpdot * perpdot) / np.float64(length2)
            beta2 = np.float64(dot * dot) / np.float64(length2)
        if dot < 0:
            d2 = alpha2 + beta2
        elif beta2 > length2:
            d2 = alpha2 + beta2 - 2 * dot + length2
        else:
            d2 = alpha2
        return d2
