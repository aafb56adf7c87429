"""ORACLE (test infrastructure only): numpy forms of the built-in models.

The reference has no built-in models (its `func` is user Python); these are the
numpy callables the reference is run with when it is compared with, or timed
beside, the CUDA built-ins of mc3_b200/csrc/models.cuh (same formulas, same
parameter order).  Signature is the reference's: func(params, x).
"""
import numpy as np


def polynomial(p, x):
    """y = sum_k p[k] x**k   (get_started's quad() is the 3-parameter case)."""
    y = np.zeros_like(x, dtype=float)
    for c in p[::-1]:
        y = y*x + c
    return y


def quad(p, x):
    """examples/get_started.py:5-16, verbatim formula."""
    return p[0] + p[1]*x + p[2]*x**2.0


def sinusoid(p, x):
    """y = p0 sin(2 pi x/p1 + p2) + p3 + p4 x   (BASELINE config 2)."""
    return p[0]*np.sin(2.0*np.pi*x/p[1] + p[2]) + p[3] + p[4]*x


def gaussian(p, x):
    """y = p0 exp(-0.5 ((x-p1)/p2)^2) + p3."""
    return p[0]*np.exp(-0.5*((x - p[1])/p[2])**2) + p[3]


def box(p, x):
    """y = p3 - p0 [ |x-p1| < p2/2 ]   (transit-like box, BASELINE config 3)."""
    return p[3] - p[0]*(np.abs(x - p[1]) < 0.5*p[2])


MODELS = {'polynomial': polynomial, 'sinusoid': sinusoid,
          'gaussian': gaussian, 'box': box}
