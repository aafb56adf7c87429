"""Piecewise-uniform abscissae (host side, numpy only).

A time series of constant cadence with gaps is not a uniform grid, but it is one
piece by piece.  `tile_layout` finds the runs of points spaced `dx` apart, cuts them
into whole tiles of 128 points and returns the reordering that puts those tiles
first and the points that fill no tile last: the layout `mc3b_chisq_opts_t.tile_x`
describes (include/mc3b200.h).  The reference has no counterpart: it evaluates the
user's model point by point (mc3/chain.py:316-319)."""
import numpy as np

TILE = 128


def tile_layout(x, max_leftover=0.02):
    """x: 1-D float64 abscissa.  Returns None when x is not (mostly) piecewise uniform,
    else a dict with
      dx       the common step
      starts   index into x of the first point of every whole tile
      perm     permutation of range(n): the tiles' points in order, then the leftover points
      nleft    number of leftover points (at most max(64, max_leftover * n))."""
    x = np.asarray(x, dtype=np.double)
    n = x.size
    if x.ndim != 1 or n < 2*TILE:
        return None
    d = np.diff(x)
    if not np.all(np.isfinite(x)) or not np.all(d > 0.0):
        return None
    tol = 8*np.finfo(float).eps*float(np.max(np.abs(x)))
    dx = float(np.median(d))
    on = np.abs(d - dx) <= 4*tol                    # steps that continue a run
    # refine dx on the longest run (a median of rounded differences is only good to ~eps |x|)
    edges = np.flatnonzero(~on)
    b = np.concatenate(([0], edges + 1))            # first point of each run
    e = np.concatenate((edges, [n - 1]))            # last point of each run
    k = int(np.argmax(e - b))
    if e[k] - b[k] < TILE:
        return None
    dx = float((x[e[k]] - x[b[k]])/(e[k] - b[k]))
    on = np.abs(d - dx) <= 4*tol
    edges = np.flatnonzero(~on)
    b = np.concatenate(([0], edges + 1))
    e = np.concatenate((edges, [n - 1]))
    nt = (e - b + 1)//TILE
    if nt.sum() == 0:
        return None
    starts = np.concatenate([b[i] + TILE*np.arange(nt[i]) for i in range(b.size) if nt[i] > 0])
    # every tile must follow its own origin to within the tolerance (drift inside long runs)
    idx = starts[:, None] + np.arange(TILE)[None, :]
    dev = np.abs(x[idx] - (x[starts][:, None] + np.arange(TILE)[None, :]*dx)).max(axis=1)
    starts = starts[dev <= tol]
    if starts.size == 0:
        return None
    intile = np.zeros(n, dtype=bool)
    idx = starts[:, None] + np.arange(TILE)[None, :]
    intile[idx.ravel()] = True
    left = np.flatnonzero(~intile)
    if left.size > max(64, int(max_leftover*n)):
        return None
    perm = np.concatenate((idx.ravel(), left))
    return dict(dx=dx, starts=starts, perm=perm, nleft=int(left.size))
