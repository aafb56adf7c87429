// mc3_b200 -- the headline kernel: sinusoid + line on a uniform abscissa grid with
// one uncertainty for all points (BASELINE config 2), fused with chi-squared and,
// optionally, the Metropolis step.  Bound: FP64 pipe.
#include "chisq_args.cuh"

namespace {

// ---- sinusoid + line on a uniform abscissa grid ------------------------------
// With x_i = x_0 + i dx the abscissa is never streamed, and with one uncertainty
// for all points (USIG, BASELINE config 2: sigma = 0.5) neither are the weights:
//   chisq = sigma^-2 sum_i (A sin(k x_i + ph) + c0 + sl x_i - d_i)^2.
// Only the data tile is staged (TMA bulk copy, NSTAGE deep, full/empty mbarriers:
// no CTA-wide barrier in the tile loop).  Six FP64 instructions per chain-point,
//   t  = Lb + s                      block base of the line + A sin
//   y  = fma(dL4, j, t)              exact line of step j (j an immediate): the line
//                                    is never a running sum, so offsets of 1e5 sigma
//                                    keep 1e-10 on chi-squared
//   r  = y - d ;  q = fma(r, r, q)
//   du = fma(nkap, s, du) ; s = s + du      Reinsch recurrence (models.cuh SineGridModel)
// each reading at most two fresh vector registers.  What bounds the loop is the
// register file, not the FP64 pipe's issue rate: one 64-bit operand per cycle and
// scheduler (profiles/r2_fp64_probe.md: DADD/DFMA with two fresh registers 2.04
// cycles, three fresh 3.03, a shared multiplier 2.2, + 1.2 per LDS.128; DMMA.8x8x4
// runs on the same pipe at 16 cycles, so the tensor form of the line gains nothing).
// Four interleaved sequences (points 0..3 mod 4) per chain, one chain per lane.
// Sequences restart every RESTART tiles from the first point of the interval, which
// comes from fast_sincos_core every REANCHOR-th restart and one rotation in
// between.  Accuracy as SineGridModel; the same guard sends stiff or huge-argument
// chains to the direct evaluation.
// Per-point uncertainties (!USIG): the stage also holds 1/sigma; every warp scales
// the data tile into a private buffer once (d/sigma, warp-synchronous) and the
// residual is fma(y, 1/sigma, -d/sigma).
namespace grid {
constexpr int TILE = 128;           // points per TMA stage and per schedule unit
#ifndef MC3B_GRID_NSTAGE
#define MC3B_GRID_NSTAGE 4
#endif
constexpr int NSTAGE = MC3B_GRID_NSTAGE;
#ifndef MC3B_GRID_RESTART
#define MC3B_GRID_RESTART 4
#endif
constexpr int RESTART = MC3B_GRID_RESTART;   // tiles between restarts of the sequences
constexpr int REANCHOR = 4;         // restarts between direct sincos evaluations of the anchor
}

// 4 CTAs per SM: 120 registers hold the sequences and the per-chain constants without
// spilling (6 CTAs at 80 registers measured 0.1829 ms, 4 at 120: 0.1778 ms at config 2)
#ifndef MC3B_GRID_MINB
#define MC3B_GRID_MINB 4
#endif
template <bool USIG>
__global__ void __launch_bounds__(WARPS * 32, MC3B_GRID_MINB) k_sinegrid(ChisqArgs<double> a) {
    using namespace grid;
    asm volatile("griddepcontrol.launch_dependents;");   // the next proposal kernel may start its draws
    __shared__ __align__(128) double sd[NSTAGE][TILE];
    __shared__ __align__(128) double sw[USIG ? 1 : NSTAGE][USIG ? 2 : TILE];
    __shared__ __align__(128) double pw[USIG ? 1 : WARPS][USIG ? 2 : TILE];      // per-warp d/sigma
    __shared__ __align__(8) uint64_t full[NSTAGE], empty[NSTAGE];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t c = ((int64_t)blockIdx.x * WARPS + warp) * 32 + lane;
    const bool live = c < a.nchains;
    if (!live) c = a.nchains - 1;                   // idle lanes shadow the last chain

    const double x0 = a.x[0];
    const double dx = (a.x[a.n - 1] - x0) / (double)(a.n - 1);
    const double* p = a.params + c * a.ldp;
    const double amp = p[0], k = 6.283185307179586476925287 / p[1], ph = p[2], c0 = p[3], sl = p[4];
    const double dth = k * dx;
    double cd1, sd1, cdT, sdT, sh, ch;
    fast_sincos_core(dth, sd1, cd1);
    fast_sincos_core(2.0 * dth, sh, ch);
    fast_sincos_core((double)(RESTART * TILE) * dth, sdT, cdT);
    const double sD = 2.0 * sh * ch;                // sin D, D = 4 dth
    const double hk = 2.0 * sh * sh;                // 1 - cos D
    double nkap = -2.0 * hk;                        // -4 sin^2(D/2)
    double dL4 = 4.0 * sl * dx;
    asm volatile("" : "+d"(dL4), "+d"(nkap));       // loop constants stay registers (no re-multiplication)
    // direct evaluation for this chain: |RESTART T dth| beyond the fast range, or D within ~0.14 rad of pi
    const int keybase = (sin_arg_key((double)(RESTART * TILE) * dth) >= MC3B_SIN_KEY_LIMIT || ch * ch < 0.005)
                            ? MC3B_SIN_KEY_LIMIT : 0;
    auto direct = [&](double x) { return fma(amp, sin(fma(x, k, ph)), fma(sl, x, c0)); };

    const int64_t nfull = a.n / TILE;
    int64_t tb, te;
    if (a.nsched > 0) { tb = a.tstart[blockIdx.y]; te = a.tstart[blockIdx.y + 1]; }
    else { tb = nfull * blockIdx.y / gridDim.y; te = nfull * (blockIdx.y + 1) / gridDim.y; }
    const int64_t nt = te - tb;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int64_t t, int s) {
        mbar_expect_tx(&full[s], (USIG ? 1u : 2u) * TILE * sizeof(double));
        bulk_g2s(sd[s], a.d + t * TILE, TILE * sizeof(double), &full[s]);
        if constexpr (!USIG) bulk_g2s(sw[s], a.w + t * TILE, TILE * sizeof(double), &full[s]);
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < NSTAGE && s < nt; s++) issue(tb + s, s);

    double acc = 0.0, S0 = 0.0, C0 = 0.0;
    double s[4], du[4], Lu[4], q[4];
    int rcount = 0, key = 0;
    for (int64_t it = 0; it < nt; it++) {
        const int st = (int)(it % NSTAGE);
        const uint32_t par = (uint32_t)((it / NSTAGE) & 1);
        const int tr = (int)(it % RESTART);             // tile within the restart interval
        const double xt = fma((double)((tb + it) * TILE), dx, x0);
        if (tr == 0) {
            // ---- restart: anchors of the four sequences (no data needed yet) ----
            const double th = fma(xt, k, ph);
            key = max(keybase, max(sin_arg_key(th), sin_arg_key(fma((double)(RESTART * TILE), dth, th))));
            if (rcount == 0) {
                fast_sincos_core(th, S0, C0);
                S0 *= amp; C0 *= amp;
            } else {
                const double sn = fma(C0, sdT, S0 * cdT);
                C0 = fma(-S0, sdT, C0 * cdT);
                S0 = sn;
            }
            rcount = (rcount + 1 == REANCHOR) ? 0 : rcount + 1;
            double cc[4];
            s[0] = S0; cc[0] = C0;
#pragma unroll
            for (int u = 1; u < 4; u++) {
                s[u] = fma(cc[u - 1], sd1, s[u - 1] * cd1);
                cc[u] = fma(-s[u - 1], sd1, cc[u - 1] * cd1);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                du[u] = fma(cc[u], sD, s[u] * hk);       // A sin(th_u) - A sin(th_u - D)
                Lu[u] = fma(sl, fma((double)u, dx, xt), c0);   // line at the first point of each sequence
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) q[u] = 0.0;
        mbar_wait(&full[st], par);
        const double* dt = sd[st];                  // what the residual subtracts
        const double* wt = sw[USIG ? 0 : st];
        if constexpr (!USIG) {
            __syncwarp();                           // the warp is done with its previous tile
#pragma unroll
            for (int i = lane; i < TILE; i += 32) pw[warp][i] = sd[st][i] * sw[st][i];
            __syncwarp();
            dt = pw[warp];
        }
        if (key < MC3B_SIN_KEY_LIMIT) {
            // 32 steps of 4 points per tile, in blocks of 8 steps: the line of step m is
            // Lu + m dL4 with ONE rounding (block base by FMA, step offset an immediate),
            // never a running sum -- offsets of 1e5 sigma keep 1e-10 on chi-squared
            double mb = (double)(tr * (TILE / 4));
#pragma unroll 1
            for (int o = 0; o < TILE; o += 32) {
                double Lb[4];
#pragma unroll
                for (int u = 0; u < 4; u++) Lb[u] = fma(dL4, mb, Lu[u]);
                mb += 8.0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int i = o + 4 * j;
                    const double2 d01 = *reinterpret_cast<const double2*>(&dt[i]);
                    const double2 d23 = *reinterpret_cast<const double2*>(&dt[i + 2]);
                    const double d[4] = {d01.x, d01.y, d23.x, d23.y};
                    double r[4];
                    if constexpr (USIG) {
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            r[u] = (j == 0 ? Lb[u] + s[u] : fma(dL4, (double)j, Lb[u] + s[u])) - d[u];
                    } else {
                        const double2 w01 = *reinterpret_cast<const double2*>(&wt[i]);
                        const double2 w23 = *reinterpret_cast<const double2*>(&wt[i + 2]);
                        const double w[4] = {w01.x, w01.y, w23.x, w23.y};
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            r[u] = fma(j == 0 ? Lb[u] + s[u] : fma(dL4, (double)j, Lb[u] + s[u]), w[u], -d[u]);
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) q[u] = fma(r[u], r[u], q[u]);
#pragma unroll
                    for (int u = 0; u < 4; u++) du[u] = fma(nkap, s[u], du[u]);
#pragma unroll
                    for (int u = 0; u < 4; u++) s[u] += du[u];
                }
            }
        } else {                                    // guarded chains: library sine per point
            double qd = 0.0;
            for (int i = 0; i < TILE; i++) {
                const double r = USIG ? direct(fma((double)i, dx, xt)) - sd[st][i]
                                      : (direct(fma((double)i, dx, xt)) - sd[st][i]) * sw[USIG ? 0 : st][USIG ? 0 : i];
                qd = fma(r, r, qd);
            }
            q[0] = qd;
        }
        acc += (q[0] + q[1]) + (q[2] + q[3]);
        // ---- stage hand-back: this warp is done with tile `it`; the producer refills
        // the stage of the PREVIOUS tile (every warp has long left it) -------------
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (threadIdx.x == 0 && it >= 1 && it - 1 + NSTAGE < nt) {
            const int sp = (int)((it - 1) % NSTAGE);
            mbar_wait(&empty[sp], (uint32_t)(((it - 1) / NSTAGE) & 1));
            issue(tb + it - 1 + NSTAGE, sp);
        }
    }

    if (blockIdx.y == gridDim.y - 1) {              // ragged tail, straight from global memory
        double t = 0.0;
        for (int64_t i = nfull * TILE; i < a.n; i++) {
            const double r = USIG ? direct(a.x[i]) - a.d[i] : (direct(a.x[i]) - a.d[i]) * a.w[i];
            t = fma(r, r, t);
        }
        acc += t;
    }
    const double w0 = USIG ? a.w[0] : 1.0;
    if (live) a.partial[(int64_t)blockIdx.y * a.ldpartial + c] = USIG ? acc * (w0 * w0) : acc;
#ifndef MC3B_NO_FUSE_CODE
    if (a.f.on) fused_metropolis(a.f, a.partial, a.ldpartial, a.nchains, WARPS * 32);
#endif
}

}  // namespace

int mc3b_launch_sinegrid(const ChisqArgs<double>& a, bool usig, unsigned groups, unsigned nsplit, cudaStream_t st) {
    if (usig) k_sinegrid<true><<<dim3(groups, nsplit), WARPS * 32, 0, st>>>(a);
    else k_sinegrid<false><<<dim3(groups, nsplit), WARPS * 32, 0, st>>>(a);
    MC3B_CHECK_LAUNCH("k_sinegrid");
    return MC3B_OK;
}
