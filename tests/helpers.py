"""Shared helpers for the parity tests (fixture decoding; no compute)."""
import numpy as np


def scal(g):
    """Decode the `scal` vector written by tests/golden/make_golden.py:gen_capture."""
    xt, yt, rt, P, dur, spacing, umbra, tol, maxstep, base, conf = g["scal"]
    return dict(xt=xt, yt=yt, rt=rt, P=int(P), duration=dur, spacing=spacing, umbra=umbra, tol=tol,
                maxstep=maxstep, base=base, confined=bool(conf > 0))


def geom(g, prefix):
    xmin, xmax, ymin, ymax, dx, dy, nrows, ncols, tw = g[prefix + "geom"]
    return dict(xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax, deltax=dx, deltay=dy, nrows=int(nrows),
                ncols=int(ncols), total_weight=tw)


def traces_of(g):
    off = g["offsets"]
    v = g["verts"]
    return [v[off[i]:off[i + 1]] for i in range(len(off) - 1)]
