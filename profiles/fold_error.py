"""numpy emulation of the folded uniform-grid sinusoid kernel (csrc/chisq_grid.cu, k_sinefold): chi-squared of
A sin(k x + ph) + c0 + sl x on x_i = x0 + i dx from point PAIRS mirrored about the centre of each 16-point block,
    r+ + r- = 2 [A sin(th_c) cos(dl h) + L_c - (d+ + d-)/2],   r+ - r- = 2 [A cos(th_c) sin(dl h) + sl dx dl - (d+ - d-)/2]
    r+^2 + r-^2 = 2 (u^2 + v^2)
against the direct evaluation in double and long double.  Runs on the CPU."""
import numpy as np
LD = np.longdouble
TILE, BLK, RESTART, REANCHOR = 128, 16, 4, 4


def fold(d):
    """The chain-independent pre-pass (mc3b_fold_data): per 16-point block, pair p = 0..7 joins points 7-p and 8+p;
    out[2p] = -(d+ + d-)/2, out[2p+1] = -(d+ - d-)/2."""
    nb = d.size//BLK
    b = d[:nb*BLK].reshape(nb, BLK)
    lo, hi = b[:, 7::-1], b[:, 8:]
    out = np.empty((nb, BLK))
    out[:, 0::2] = -0.5*(hi + lo)
    out[:, 1::2] = -0.5*(hi - lo)
    return out.ravel()


def folded_chisq(P, x0, dx, f, n):
    """P [nc, 5]; returns sum of squared residuals over the full tiles (nc,)"""
    amp, k, ph, c0, sl = P[:, 0], 2*np.pi/P[:, 1], P[:, 2], P[:, 3], P[:, 4]
    h = k*dx
    # tables: cos/sin((p + 1/2) h) by rotation from h/2
    c1, s1 = np.cos(h), np.sin(h)
    cp, sp = [np.cos(0.5*h)], [np.sin(0.5*h)]
    for p in range(1, 8):
        cp.append(cp[-1]*c1 - sp[-1]*s1)
        sp.append(sp[-1]*c1 + cp[-2]*s1)
    c16, s16 = np.cos(BLK*h), np.sin(BLK*h)
    cT, sT = np.cos(RESTART*TILE*h), np.sin(RESTART*TILE*h)
    gs = sl*dx
    dL16 = 16.0*gs
    acc = np.zeros(P.shape[0])
    rcount = 0
    ntile = n//TILE
    for t in range(ntile):
        xc = x0 + (t*TILE)*dx + 7.5*dx                # centre of the tile's first block
        if t % RESTART == 0:
            if rcount == 0:
                th = xc*k + ph
                S0, C0 = amp*np.sin(th), amp*np.cos(th)
            else:
                S0, C0 = C0*sT + S0*cT, -S0*sT + C0*cT
            rcount = (rcount + 1) % REANCHOR
            Sc, Cc = S0.copy(), C0.copy()
        Lt = sl*xc + c0
        q = np.zeros_like(acc)
        for b in range(TILE//BLK):
            Lc = dL16*b + Lt
            fb = f[t*TILE + b*BLK: t*TILE + (b + 1)*BLK]
            for p in range(8):
                u = Sc*cp[p] + (Lc + fb[2*p])
                v = Cc*sp[p] + (gs*(p + 0.5) + fb[2*p + 1])
                q = q + u*u
                q = q + v*v
            Sc, Cc = Cc*s16 + Sc*c16, -Sc*s16 + Cc*c16
        acc += q
    return 2.0*acc


def direct(P, x, d, dtype=float):
    P = P.astype(dtype)
    x, d = x.astype(dtype), d.astype(dtype)
    m = P[:, 0:1]*np.sin(2*dtype(np.pi)*x[None, :]/P[:, 1:2] + P[:, 2:3]) + P[:, 3:4] + P[:, 4:5]*x[None, :]
    return ((m - d[None, :])**2).sum(1)


if __name__ == '__main__':
    rs = np.random.RandomState(3)
    for n, off in ((100_000, 5.0), (20_480, 5e4)):
        x = np.linspace(0, 10, n)
        pt = np.array([1.0, 2.5, 0.3, off, -0.2])
        d = pt[0]*np.sin(2*np.pi*x/pt[1] + pt[2]) + pt[3] + pt[4]*x + rs.normal(0, 0.5, n)
        nc = 64
        P = pt*(1 + 0.02*rs.standard_normal((nc, 5)))
        P[:8, 1] = 10**rs.uniform(-3.5, -1, 8)         # short periods (up to ~pi rad per sample)
        x0, dx = x[0], (x[-1] - x[0])/(n - 1)
        nt = n//TILE*TILE
        got = folded_chisq(P, x0, dx, fold(d), n)
        ref = direct(P, x[:nt], d[:nt], LD).astype(float)
        dbl = direct(P, x[:nt], d[:nt])
        print(f'n={n} offset={off}: folded vs long double {np.max(np.abs(got/ref - 1)):.2e}, '
              f'direct double vs long double {np.max(np.abs(dbl/ref - 1)):.2e}')
