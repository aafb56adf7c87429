"""numpy emulation of the sufficient-statistics form of the uniform-grid sinusoid kernel (csrc/chisq_grid.cu,
k_sinemom): with the data centred on a reference line and folded into pair half-sums e and half-differences o,
    sum_16 r^2 = 2 [ Sc (Sc Kcc + 2 Lc Kc - 2 Pe) + Cc (Cc Kss + 2 g Ksd - 2 Po) ] + 2 [line/data terms from tile moments]
    Pe = sum_p cp_p e_p,  Po = sum_p sp_p o_p        (the only per-point work: one FMA per point)
The expansion subtracts sums of size ~ (|s| + |L'| + |d'|)^2 to leave chi-squared, so its relative error is
eps_eff * amp with amp = (|s| + |L'| + |d'|)^2 / (sigma^2 chi^2); the kernel's guard sends chains with
amp > AMP_MAX to the exact evaluation.  This script measures eps_eff against long double.  Runs on the CPU."""
import numpy as np
LD = np.longdouble
TILE, BLK, RESTART, REANCHOR = 128, 16, 4, 4
DEL = np.arange(8) + 0.5


def prepare(d, x0, dx, c0r, slr):
    """chain-independent pass: centred, folded, pre-scaled data and the per-tile moments"""
    n = d.size//TILE*TILE
    i = np.arange(n)
    dc = d[:n] - (c0r + slr*(x0 + i*dx))
    b = dc.reshape(-1, BLK)
    lo, hi = b[:, 7::-1], b[:, 8:]
    e, o = 0.5*(hi + lo), 0.5*(hi - lo)
    f = np.empty_like(b)
    f[:, 0::2], f[:, 1::2] = -2*e, -2*o
    nt = n//TILE
    E1 = e.sum(1).reshape(nt, 8)
    O1 = (o*DEL).sum(1).reshape(nt, 8)
    D2 = (e*e + o*o).sum(1).reshape(nt, 8)
    bp = np.arange(8) - 3.5
    mom = np.stack([E1.sum(1), 16*(E1*bp).sum(1) + O1.sum(1), D2.sum(1)], 1)
    return f.ravel(), mom, float((dc*dc).sum())


def moment_chisq(P, x0, dx, f, mom, n, c0r, slr):
    amp, k, ph = P[:, 0], 2*np.pi/P[:, 1], P[:, 2]
    c0, sl = P[:, 3] - c0r, P[:, 4] - slr
    h = k*dx
    c1, s1 = np.cos(h), np.sin(h)
    cp, sp = [np.cos(0.5*h)], [np.sin(0.5*h)]
    for p in range(1, 8):
        cp.append(cp[-1]*c1 - sp[-1]*s1)
        sp.append(sp[-1]*c1 + cp[-2]*s1)
    Kcc, Kc2 = sum(c*c for c in cp), 2*sum(cp)
    Kss, Ksd = sum(s*s for s in sp), sum(s*d for s, d in zip(sp, DEL))
    c16, s16 = np.cos(BLK*h), np.sin(BLK*h)
    cT, sT = np.cos(RESTART*TILE*h), np.sin(RESTART*TILE*h)
    g = sl*dx
    gK = 2*g*Ksd
    dL16 = 16.0*g
    acc = np.zeros(P.shape[0])
    rcount = 0
    for t in range(n//TILE):
        xc = x0 + (t*TILE + 7.5)*dx
        if t % RESTART == 0:
            if rcount == 0:
                th = xc*k + ph
                S0, C0 = amp*np.sin(th), amp*np.cos(th)
            else:
                S0, C0 = C0*sT + S0*cT, -S0*sT + C0*cT
            rcount = (rcount + 1) % REANCHOR
            Sc, Cc = S0.copy(), C0.copy()
        L0 = sl*xc + c0
        q = np.zeros_like(acc)
        for b in range(8):
            Lc = dL16*b + L0
            fb = f[t*TILE + b*BLK: t*TILE + (b + 1)*BLK]
            Pe = sum(cp[p]*fb[2*p] for p in range(8))
            Po = sum(sp[p]*fb[2*p + 1] for p in range(8))
            q = q + Sc*(Sc*Kcc + (Lc*Kc2 + Pe))
            q = q + Cc*(Cc*Kss + (gK + Po))
            Sc, Cc = Cc*s16 + Sc*c16, -Sc*s16 + Cc*c16
        Lt = L0 + 56.0*g                       # line at the tile centre (63.5 - 7.5 points on)
        q = q + (Lt*(64.0*Lt - 2*mom[t, 0]) + (g*(87376.0*g - 2*mom[t, 1]) + mom[t, 2]))
        acc += q
    return 2.0*acc


def amp_bound(P, n, x0, dx, D2, c0r, slr, chi):
    """(|s|max + |L'| + |d'|)^2 / chi"""
    c0, sl = P[:, 3] - c0r, P[:, 4] - slr
    xm = x0 + 0.5*(n - 1)*dx
    L2 = n*(c0 + sl*xm)**2 + (sl*dx)**2*n*(n*n - 1.0)/12.0
    return (np.abs(P[:, 0])*np.sqrt(n) + np.sqrt(L2) + np.sqrt(D2))**2/chi


def direct(P, x, d, dtype=float):
    P, x, d = P.astype(dtype), x.astype(dtype), d.astype(dtype)
    m = P[:, 0:1]*np.sin(2*dtype(np.pi)*x[None, :]/P[:, 1:2] + P[:, 2:3]) + P[:, 3:4] + P[:, 4:5]*x[None, :]
    return ((m - d[None, :])**2).sum(1)


if __name__ == '__main__':
    rs = np.random.RandomState(3)
    print('   n   S/N  offset | max amp | max rel err | err/amp (eps_eff)')
    for n, snr, off in ((100_000, 2.0, 5.0), (100_000, 30.0, 5.0), (20_480, 300.0, 5e4), (20_480, 1.0, -3.0),
                        (8192, 3000.0, 1.0)):
        x = np.linspace(0, 10, n)
        pt = np.array([1.0, 2.5, 0.3, off, -0.2])
        sig = pt[0]/snr
        d = pt[0]*np.sin(2*np.pi*x/pt[1] + pt[2]) + pt[3] + pt[4]*x + rs.normal(0, sig, n)
        nc = 96
        P = pt*(1 + 0.02/snr*rs.standard_normal((nc, 5)))
        P[:8, 1] = 10**rs.uniform(-3.2, -1, 8)
        x0, dx = x[0], (x[-1] - x[0])/(n - 1)
        slr, c0r = np.polyfit(x, d, 1)
        f, mom, D2 = prepare(d, x0, dx, c0r, slr)
        nt = n//TILE*TILE
        got = moment_chisq(P, x0, dx, f, mom, n, c0r, slr)
        ref = direct(P, x[:nt], d[:nt], LD).astype(float)
        amp = amp_bound(P, nt, x0, dx, D2, c0r, slr, ref)
        err = np.abs(got/ref - 1)
        print(f'{n:7d} {snr:6.0f} {off:7.0f} | {amp.max():9.2e} | {err.max():.2e} | {np.max(err/amp):.2e}')
