"""numpy emulation of SineGridModel's recurrences (models.cuh): error of A sin against long double,
measured from the double-rounded phase of each anchor tile (so that the rounding of the argument
itself, which any evaluation order shares, is not counted).  Runs on the CPU."""
import numpy as np
LD = np.longdouble


def run(dth, th0, A, ntile=16, TILE=128, reanchor=8):
    sd1, cd1 = np.sin(dth), np.cos(dth)
    sh, ch = np.sin(2*dth), np.cos(2*dth)
    sD, hk = 2*sh*ch, 2*sh*sh
    nkap = -2*hk
    sdT, cdT = np.sin(TILE*dth), np.cos(TILE*dth)
    out = np.zeros((ntile*TILE,) + np.shape(th0))
    ref = np.zeros_like(out)
    for t in range(ntile):
        th = th0 + t*TILE*dth
        if t % reanchor == 0:
            S0, C0, anchor, ta = A*np.sin(th), A*np.cos(th), th.copy(), t
        else:
            S0, C0 = C0*sdT + S0*cdT, -S0*sdT + C0*cdT
        s, c = [S0], [C0]
        for u in range(1, 4):
            s.append(c[u-1]*sd1 + s[u-1]*cd1)
            c.append(-s[u-1]*sd1 + c[u-1]*cd1)
        du = [c[u]*sD + s[u]*hk for u in range(4)]
        for i in range(0, TILE, 4):
            for u in range(4):
                j = (t - ta)*TILE + i + u
                out[t*TILE + i + u] = s[u]
                ref[t*TILE + i + u] = (A*np.sin(anchor.astype(LD) + j*dth.astype(LD))).astype(float)
                du[u] = nkap*s[u] + du[u]
                s[u] = s[u] + du[u]
    return out, ref


if __name__ == '__main__':
    rng = np.random.default_rng(5)
    for lo, hi in ((-5, -2), (-2, -0.5), (-0.5, 0.4)):
        worst = 0.0
        for trial in range(10):
            n = 4000
            dth = 10**rng.uniform(lo, hi, n)*rng.choice([-1, 1], n)
            ok = np.cos(2*dth)**2 >= 0.005          # the kernel sends the others to the direct path
            th0, A = rng.uniform(-50, 50, n), rng.uniform(0.5, 2, n)
            out, ref = run(dth, th0, A)
            worst = max(worst, (np.abs(out - ref)/A)[:, ok].max())
        print(f'k dx in 10^[{lo}, {hi}] rad per sample: worst |error|/A = {worst:.2e}')
