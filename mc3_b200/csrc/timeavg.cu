// mc3_b200 -- time-series diagnostics (replace src_c/_time_averaging.c and
// src_c/_binarray.c).
//
// binrms: the reference bins the series once per bin size, O(N * nsizes)
// (_time_averaging.c:98-110).  Here every bin sum comes from block-local prefix
// sums: P[i] = sum of the series from the start of i's 1024-point block to i,
// plus an exclusive prefix T2 over the block totals.  A bin [s, e) is
//   inside one block:  P[e-1] - head(s)
//   across blocks:     (P[end of s's block] - head(s)) + (T2[be] - T2[bs+1]) + P[e-1]
// so the work is O(N log(maxbins)) bins in total and one 8N-byte streaming pass
// to build P.  Local prefixes keep the magnitudes small: bin sums agree with
// the reference's direct sums to ~1e-14 relative.  Reductions are fixed-order.
//
// binarray: HBM-bound streaming; one warp per bin with coalesced loads (measured
// faster than shared-memory staging, plain or TMA: see profiles/).
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int PB = 1024;            // prefix block (4 per thread, 256 threads)
constexpr int IGN = 10000;          // inverse-gamma grid (stats.h:140)
constexpr int YS = 64;              // splits of the bins of one bin size

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        for (int k = 0; k < 8; k++) t += sh[k];
        sh[8] = t;
    }
    __syncthreads();
    t = sh[8];
    __syncthreads();
    return t;
}

// P = block-local inclusive prefix (when WRITEP); tot[b] = block total.
template <bool WRITEP>
__global__ void __launch_bounds__(256) k_block_prefix(const double* x, int64_t n, double* P, double* tot) {
    __shared__ double wsum[8];
    const int64_t base = (int64_t)blockIdx.x * PB + threadIdx.x * 4;
    double v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (base + k < n) ? x[base + k] : 0.0;
    v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
    double run = v[3];                                   // inclusive scan of thread totals
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, run, o);
        if (lane >= o) run += t;
    }
    if (lane == 31) wsum[warp] = run;
    __syncthreads();
    double woff = 0.0;
    for (int k = 0; k < warp; k++) woff += wsum[k];
    const double excl = woff + run - v[3];
    if (WRITEP) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (base + k < n) P[base + k] = v[k] + excl;
    }
    if (threadIdx.x == 255) tot[blockIdx.x] = woff + run;
}

// Single CTA: T2 = exclusive prefix of tot (nblk+1 entries), stat[0] = mean.
__global__ void __launch_bounds__(256) k_totals_scan(const double* tot, int64_t nblk, int64_t n, double* T2,
                                                    double* stat) {
    __shared__ double sh[9];
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < nblk; b0 += 256) {
        const int64_t b = b0 + threadIdx.x;
        const double v = b < nblk ? tot[b] : 0.0;
        double run = v;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, run, o);
            if (lane >= o) run += t;
        }
        if (lane == 31) sh[warp] = run;
        __syncthreads();
        double woff = 0.0;
        for (int k = 0; k < warp; k++) woff += sh[k];
        const double c = carry;
        if (b < nblk) T2[b] = c + woff + run - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = c + woff + run;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        T2[nblk] = carry;
        stat[0] = carry / (double)n;
    }
}

// dev[b] = sum over block b of (x - mean)^2      (stats.h:60-72, population std)
__global__ void __launch_bounds__(256) k_dev2(const double* x, int64_t n, const double* stat, double* dev) {
    __shared__ double sh[9];
    const double mu = stat[0];
    const int64_t base = (int64_t)blockIdx.x * PB;
    double a = 0.0;
    for (int k = threadIdx.x; k < PB; k += 256)
        if (base + k < n) { const double d = x[base + k] - mu; a = fma(d, d, a); }
    a = block_sum_256(a, sh);
    if (threadIdx.x == 0) dev[blockIdx.x] = a;
}

// Single CTA: stat[1] = sqrt(sum(dev)/n)
__global__ void __launch_bounds__(256) k_std(const double* dev, int64_t nblk, int64_t n, double* stat) {
    __shared__ double sh[9];
    double a = 0.0;
    for (int64_t b = threadIdx.x; b < nblk; b += 256) a += dev[b];
    a = block_sum_256(a, sh);
    if (threadIdx.x == 0) stat[1] = sqrt(a / (double)n);
}

// One streaming pass for the mean and the population standard deviation when no
// prefix is needed (every bin size fits the tile kernel): per block of PB points its
// sum and its squared deviations about ITS OWN mean (two sweeps over registers), then
// k_moments_finish combines the blocks exactly,
//   mean = sum_b s_b / n,   M2 = sum_b [ M2_b + n_b (mean_b - mean)^2 ],
// which is as accurate as the reference's two passes over the data (stats.h:60-72)
// at half the traffic.
__global__ void __launch_bounds__(256) k_moments(const double* __restrict__ x, int64_t n, double* tot, double* dev) {
    __shared__ double sh[9];
    const int64_t base = (int64_t)blockIdx.x * PB + threadIdx.x * 4;
    double v[4];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const bool in = base + k < n; v[k] = in ? x[base + k] : 0.0; cnt += in; }
    const double s = block_sum_256((v[0] + v[1]) + (v[2] + v[3]), sh);
    const int64_t nb = (n - (int64_t)blockIdx.x * PB) < PB ? (n - (int64_t)blockIdx.x * PB) : PB;
    const double mu = s / (double)nb;
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < cnt) { const double d = v[k] - mu; a = fma(d, d, a); }
    a = block_sum_256(a, sh);
    if (threadIdx.x == 0) { tot[blockIdx.x] = s; dev[blockIdx.x] = a; }
}

__global__ void __launch_bounds__(256) k_moments_finish(const double* tot, const double* dev, int64_t nblk, int64_t n,
                                                       double* stat) {
    __shared__ double sh[9];
    double a = 0.0;
    for (int64_t b = threadIdx.x; b < nblk; b += 256) a += tot[b];
    const double mean = block_sum_256(a, sh) / (double)n;
    double m2 = 0.0;
    for (int64_t b = threadIdx.x; b < nblk; b += 256) {
        const int64_t nb = (n - b * PB) < PB ? (n - b * PB) : PB;
        const double d = tot[b] / (double)nb - mean;
        m2 += dev[b] + (double)nb * d * d;
    }
    m2 = block_sum_256(m2, sh);
    if (threadIdx.x == 0) { stat[0] = mean; stat[1] = sqrt(m2 / (double)n); }
}

__device__ __forceinline__ double range_sum(const double* P, const double* T2, int64_t s, int64_t e) {
    const int64_t bs = s / PB, be = (e - 1) / PB;
    const double head = (s % PB) ? P[s - 1] : 0.0;
    if (bs == be) return P[e - 1] - head;
    return (P[bs * PB + PB - 1] - head) + (T2[be] - T2[bs + 1]) + P[e - 1];
}

// grid (bin-size index, split): partial[y, i] = sum over this split's bins of mean^2.
__global__ void __launch_bounds__(256) k_binrms_main(const double* P, const double* T2, int64_t n, int64_t i_begin,
                                                    int64_t nout, int64_t binstep, double* partial) {
    __shared__ double sh[9];
    for (int64_t i = i_begin + blockIdx.x; i < nout; i += gridDim.x) {
        const int64_t b = 1 + i * binstep, M = n / b;
        const int64_t j0 = M * blockIdx.y / gridDim.y, j1 = M * (blockIdx.y + 1) / gridDim.y;
        const double inv = 1.0 / (double)b;
        double a = 0.0;
        for (int64_t j = j0 + threadIdx.x; j < j1; j += 256) {
            const double m = range_sum(P, T2, j * b, (j + 1) * b) * inv;
            a = fma(m, m, a);
        }
        a = block_sum_256(a, sh);
        if (threadIdx.x == 0) partial[(int64_t)blockIdx.y * nout + i] = a;
    }
}

// Tile version for the small bin sizes (b <= BMAX): a persistent CTA stages a
// tile of TT owned points + a halo of (largest small bin - 1) points in shared
// memory with one TMA bulk copy, turns it into an inclusive prefix in place, and
// evaluates EVERY small bin size from it (a bin belongs to the tile its first
// point is in).  HBM traffic is one pass over the series (+ halo) instead of 2-4
// sectors per bin.  acc[i] (shared) collects sum of mean^2 per bin size over
// the CTA's tiles in a fixed order; partial[cta, i] leaves at the end.
#ifndef MC3B_TT
#define MC3B_TT 8192
#endif
constexpr int TT = MC3B_TT;
constexpr int BMAX = 4096;
constexpr int BWARP = 256;
#ifndef MC3B_TNW
#define MC3B_TNW 16
#endif
constexpr int TNW = MC3B_TNW;            // warps per CTA of the tile kernel          // bin sizes below this: one warp per size; above: one lane per size

// Skewed shared-memory index: a bin size b makes the lanes of a warp read the
// prefix with stride b; one padding slot per 16 entries makes every stride that is
// a multiple of 16 odd in units of 16 (17, 51, ...): two instructions per access.
// (Round 1 padded at 16/256/4096: six integer instructions per access in a loop
// that ncu showed bound by instruction issue, profiles/r2_binrms_tile.md.)
__device__ __forceinline__ int pidx(int k) { return k + (k >> 4); }

// t0 mod b for 64-bit t0 with the reciprocal at hand (one correction step).
__device__ __forceinline__ int mod_small(int64_t t0, int b, double inv) {
    const int64_t q = (int64_t)((double)t0 * inv);
    int r = (int)(t0 - q * b);
    if (r < 0) r += b; else if (r >= b) r -= b;
    return r;
}

// Q is the EXCLUSIVE prefix of the tile: Q[k] = x[t0] + ... + x[t0+k-1], Q[0] = 0, so
// the sum of a bin [s, s+b) is Q[s+b] - Q[s]: two shared-memory loads, no shuffles, no
// first-lane special case.  Work items (one bin size for a warp while b < BWARP, then
// groups of 32 sizes with a lane each) are handed out through a shared counter in
// decreasing order of cost, so the warps of a CTA reach the end-of-tile barrier together.
__global__ void __launch_bounds__(TNW * 32) k_binrms_tile(const double* __restrict__ x, int64_t n, int64_t nsmall,
                                                    int64_t binstep, int halo, int64_t nout, double* partial) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double wsum[TNW], wsq[TNW];
    __shared__ int next_item;
    const int len_full = TT + halo;
    double* Q = sm;                                   // [pidx(len_full) + 1] skewed exclusive prefix
    double* acc = sm + pidx(len_full) + 1;            // [nsmall] sum of mean^2 per bin size
    double* invb = acc + nsmall;                      // [nsmall] 1/b
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int step = (int)binstep, ns = (int)nsmall;
    for (int i = threadIdx.x; i < ns; i += TNW * 32) { acc[i] = 0.0; invb[i] = 1.0 / (double)(1 + i * step); }
    // first index whose bin size reaches BWARP
    int nA = ns;
    if (1 + (nsmall - 1) * binstep >= BWARP) nA = (BWARP - 1 + step - 1) / step;
    const int nB = (ns - nA + 31) / 32;               // lane-per-size groups
    const int nitems = (nA > 1 ? nA - 1 : 0) + nB;    // bin size 1 comes from the scan
    const int seg = ((len_full + TNW - 1) / TNW + 127) & ~127;   // points per warp in the scan (multiple of 128)
    const int64_t ntiles = (n + TT - 1) / TT;
    __syncthreads();
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t t0 = t * TT;
        const int len = (int)((n - t0) < len_full ? (n - t0) : len_full);
        const int owned = (int)((n - t0) < TT ? (n - t0) : TT);
        // ---- scan: warp w owns points [w*seg, (w+1)*seg); a lane takes 4 consecutive
        // points of every 128 (two 16-byte loads), adds them up serially and the warp
        // scans the 32 lane totals once per 128 points
        {
            const int k0 = warp * seg;
            double carry = 0.0, sq = 0.0;
            for (int c = 0; c < seg; c += 128) {
                const int k = k0 + c + 4 * lane;
                double v[4] = {0.0, 0.0, 0.0, 0.0};
                if (k + 3 < len) {
                    const double2 a = *reinterpret_cast<const double2*>(x + t0 + k);
                    const double2 b2 = *reinterpret_cast<const double2*>(x + t0 + k + 2);
                    v[0] = a.x; v[1] = a.y; v[2] = b2.x; v[3] = b2.y;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; u++) if (k + u < len) v[u] = x[t0 + k + u];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) if (k + u < owned) sq = fma(v[u], v[u], sq);
                v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
                double r = v[3];
                for (int o = 1; o < 32; o <<= 1) {
                    const double up = __shfl_up_sync(0xffffffffu, r, o);
                    if (lane >= o) r += up;
                }
                const double off = carry + (r - v[3]);   // everything before this lane's 4 points
#pragma unroll
                for (int u = 0; u < 4; u++) if (k + u < len) Q[pidx(k + u + 1)] = off + v[u];
                carry += __shfl_sync(0xffffffffu, r, 31);
            }
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            if (lane == 0) { wsum[warp] = carry; wsq[warp] = sq; }
        }
        if (threadIdx.x == 0) { Q[0] = 0.0; next_item = 0; }
        __syncthreads();
        {
            double off = 0.0;
            for (int w = 0; w < warp; w++) off += wsum[w];
            if (warp > 0) {
                const int k0 = warp * seg;
                for (int k = k0 + lane; k < k0 + seg && k < len; k += 32) Q[pidx(k + 1)] += off;
            }
            if (threadIdx.x == 0) {                    // bin size 1: sum of squares of the owned points
                double s2 = 0.0;
                for (int w = 0; w < TNW; w++) s2 += wsq[w];
                acc[0] += s2;
            }
        }
        __syncthreads();
        const bool tail = len < len_full;              // last tiles: a bin must end inside the data
        for (;;) {
            int item = 0;
            if (lane == 0) item = atomicAdd(&next_item, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= nitems) break;
            if (item < nA - 1) {
                // ---- A: one warp per bin size 1 < b < BWARP, lanes over its bins
                const int i = item + 1;
                const int b = 1 + i * step;
                const double inv = invb[i];
                const int r = mod_small(t0, b, inv);
                const int ls0 = r ? b - r : 0;         // first bin starting in the tile
                int cnt = 0;
                if (ls0 < owned) {
                    cnt = (int)((double)(owned - ls0 - 1) * inv + 1e-9) + 1;       // bins starting in [ls0, owned)
                    if (tail) { const int fit = (int)((double)(len - ls0) * inv + 1e-9); cnt = cnt < fit ? cnt : fit; }
                }
                double a = 0.0;
                for (int m = lane; m < cnt; m += 32) {
                    const int s0 = ls0 + m * b;
                    const double mean = (Q[pidx(s0 + b)] - Q[pidx(s0)]) * inv;
                    a = fma(mean, mean, a);
                }
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) acc[i] += a;            // one warp per size and tile: no race
            } else {
                // ---- B: bin sizes >= BWARP, one lane per size (few bins each)
                const int i = nA + (item - (nA - 1 > 0 ? nA - 1 : 0)) * 32 + lane;
                if (i < ns) {
                    const int b = 1 + i * step;
                    const double inv = invb[i];
                    const int r = mod_small(t0, b, inv);
                    int ls = r ? b - r : 0;
                    double a = 0.0;
                    double prev = Q[pidx(ls < len ? ls : 0)];
                    while (ls < owned && ls + b <= len) {
                        const double e = Q[pidx(ls + b)];
                        const double mean = (e - prev) * inv;
                        a = fma(mean, mean, a);
                        prev = e;
                        ls += b;
                    }
                    acc[i] += a;
                }
            }
        }
        __syncthreads();                               // tile buffer free again
    }
    for (int i = threadIdx.x; i < ns; i += TNW * 32) partial[(int64_t)blockIdx.x * nout + i] = acc[i];
}

// One thread per bin size: rms, asymptotic errors, Gaussian extrapolation, and
// the table lead[M] = first output index whose bin count is M (M <= 35).
__global__ void k_binrms_finish(const double* partial, int ys, int64_t n, int64_t nout, int64_t binstep,
                                const double* stat, double* rms, double* rmslo, double* rmshi, double* err,
                                double* binsz, int* lead) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nout) return;
    const int64_t b = 1 + i * binstep, M = n / b;
    double a = 0.0;
    int y = 0;
    for (; y + 16 <= ys; y += 16) {                  // 16 loads in flight, added in row order
        double v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = partial[(int64_t)(y + k) * nout + i];
#pragma unroll
        for (int k = 0; k < 16; k++) a += v[k];
    }
    for (; y < ys; y++) a += partial[(int64_t)y * nout + i];
    const double r = sqrt(a / (double)M);
    rms[i] = r;
    rmslo[i] = rmshi[i] = r / sqrt(2.0 * (double)M);
    err[i] = stat[1] * sqrt((double)M / ((double)b * ((double)M - 1.0)));
    binsz[i] = (double)b;
    if (M <= 35 && M >= 0) {
        const int64_t Mprev = i > 0 ? n / (1 + (i - 1) * binstep) : -1;
        if (Mprev != M) lead[M] = (int)i;
    }
}

__device__ __forceinline__ double ig_pdf(double x, int M, double s) {
    return pow(x, -(double)M) * exp(-(double)M * s * s / (2.0 * x * x));
}

// One CTA per bin count M = blockIdx.x (<= 35): 68.3% credible region of the
// inverse-gamma posterior (stats.h:139-224).  The grid densities are evaluated
// in parallel; the outward merge from the mode, the running sum and the
// boundary walk are the reference's sequential algorithm on one thread.
__global__ void __launch_bounds__(256) k_invgamma(const int* lead, const double* err, double* igws, double* lohi) {
    const int M = blockIdx.x;
    const int li = lead[M];
    if (li < 0) return;
    double* pdf = igws + (size_t)M * 3 * IGN;    // grid order
    double* xs = pdf + IGN;                        // x in descending-density order
    double* ps = xs + IGN;                         // density in that order
    const double s = err[li], ds = s / sqrt(2.0 * (double)M);
    const double xmax = s + 50.0 * ds;
    double xmin = s - 4.0 * ds;
    if (xmin < 0.01 * s) xmin = 0.01 * s;
    const double dx = (xmax - xmin) / (IGN - 1.0);
    for (int k = threadIdx.x; k < IGN; k += 256) pdf[k] = ig_pdf(xmin + k * dx, M, s);
    __syncthreads();
    if (threadIdx.x != 0) return;
    auto pd = [&](int k) { return (k >= 0 && k < IGN) ? pdf[k] : ig_pdf(xmin + k * dx, M, s); };
    int ilo = (int)((s - xmin) / dx), ihi = ilo + 1, i;
    double plo = pd(ilo), phi = pd(ihi), psum = 0.0;
    for (i = 0; i < IGN; i++) {                    // merge outward from the mode
        if (ilo < 0 || ihi >= IGN) break;
        if (plo > phi) { xs[i] = xmin + ilo * dx; ps[i] = plo; --ilo; plo = pd(ilo); }
        else           { xs[i] = xmin + ihi * dx; ps[i] = phi; ++ihi; phi = pd(ihi); }
        psum += ps[i];
    }
    for (; i < IGN; i++) {                         // one side exhausted: finish the other
        const int k = (ilo < 0) ? ihi++ : ilo--;
        xs[i] = xmin + k * dx;
        ps[i] = pd(k);
    }
    double cdf = 0.0;
    i = 0;
    while (cdf < 0.683) cdf += ps[i++] / psum;
    double low = xs[i], high = xs[--i], tmp = high;
    if (low > high) { high = low; low = tmp; }
    for (;;) {
        tmp = xs[--i];
        if (low < tmp && tmp < high) break;
        else if (tmp < low) low = tmp;
        else high = tmp;
    }
    lohi[2 * M] = s - low;
    lohi[2 * M + 1] = high - s;
}

// _time_averaging.c:124-134 -- renormalise the error bars where M <= 35.
__global__ void k_binrms_small(int64_t n, int64_t nout, int64_t binstep, const double* lohi, const double* rms,
                               const double* err, double* rmslo, double* rmshi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nout) return;
    const int64_t M = n / (1 + i * binstep);
    if (M <= 35 && M >= 0) {
        rmslo[i] = lohi[2 * M] * rms[i] / err[i];
        rmshi[i] = lohi[2 * M + 1] * rms[i] / err[i];
    }
}

__global__ void k_fill_int(int* p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- binarray ---------------------------------------------------------------
// 1/sigma^2.  (A single-precision reciprocal seed + two fp64 Newton steps measured
// SLOWER than the quotient at config 4, 0.642 vs 0.547 ms: the kernel is bound by
// the loads it keeps in flight, not by the division.)
__device__ __forceinline__ double inv_square(double s) { return 1.0 / (s * s); }

template <bool W>
__global__ void __launch_bounds__(256) k_binarray_big(const double* d, const double* u, int64_t binsize, double* bd,
                                                     double* bs) {
    __shared__ double sh[9];
    const int64_t e0 = (int64_t)blockIdx.x * binsize;
    double a = 0.0, w = 0.0;
    for (int64_t k = threadIdx.x; k < binsize; k += 256) {
        if (W) {
            const double ww = inv_square(u[e0 + k]);
            w += ww;
            a += d[e0 + k] * ww;
        } else {
            a += d[e0 + k];
        }
    }
    a = block_sum_256(a, sh);
    if (W) w = block_sum_256(w, sh);
    if (threadIdx.x == 0) {
        if (W) {
            const double sdv = sqrt(1.0 / w);
            bs[blockIdx.x] = sdv;
            bd[blockIdx.x] = a * sdv * sdv;
        } else {
            bd[blockIdx.x] = a / (double)binsize;
        }
    }
}

// Direct version (no staging): one warp per bin, lanes read the bin with
// coalesced 8-byte loads (4 independent loads in flight per lane at binsize 100,
// 64 resident warps per SM keep ~64 KB in flight), fixed-order lane tree.  Short
// bins (< 32 points) take one thread per bin.
template <bool W, int NB, int NJ>
__global__ void __launch_bounds__(256) k_binarray_direct(const double* __restrict__ d, const double* __restrict__ u,
                                                        int64_t nbins, int64_t binsize, double* bd, double* bs) {
    // NB bins per warp and pass; every lane first issues all its loads of a pass
    // (up to 8 per bin, independent), then adds: 8*NB loads in flight per lane.
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b0 = warp * NB; b0 < nbins; b0 += nwarps * NB) {
        double a[NB], w[NB];
#pragma unroll
        for (int q = 0; q < NB; q++) { a[q] = 0.0; w[q] = 0.0; }
        for (int64_t base = 0; base < binsize; base += 32 * NJ) {
            double v[NB][NJ], sg[NB][NJ];
#pragma unroll
            for (int q = 0; q < NB; q++) {
                const int64_t b = b0 + q;
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    const int64_t k = base + lane + 32 * j;
                    const bool ok = (k < binsize) && (b < nbins);
                    v[q][j] = ok ? d[b * binsize + k] : 0.0;
                    if (W) sg[q][j] = ok ? u[b * binsize + k] : 1.0;
                }
            }
#pragma unroll
            for (int q = 0; q < NB; q++) {
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    if (W) {
                        const int64_t k = base + lane + 32 * j;
                        const double ww = (k < binsize) ? inv_square(sg[q][j]) : 0.0;
                        w[q] += ww;
                        a[q] = fma(v[q][j], ww, a[q]);
                    } else {
                        a[q] += v[q][j];
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NB; q++) {
            double aa = a[q], ww = w[q];
            for (int o = 16; o > 0; o >>= 1) {
                aa += __shfl_xor_sync(0xffffffffu, aa, o);
                if (W) ww += __shfl_xor_sync(0xffffffffu, ww, o);
            }
            const int64_t b = b0 + q;
            if (lane == 0 && b < nbins) {
                if (W) {
                    const double sdv = sqrt(1.0 / ww);
                    bs[b] = sdv;
                    bd[b] = aa * sdv * sdv;
                } else {
                    bd[b] = aa / (double)binsize;
                }
            }
        }
    }
}

template <bool W>
__global__ void __launch_bounds__(256) k_binarray_short(const double* __restrict__ d, const double* __restrict__ u,
                                                       int64_t nbins, int binsize, double* bd, double* bs) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    const double* x = d + b * binsize;
    double a = 0.0, w = 0.0;
    for (int k = 0; k < binsize; k++) {
        if (W) { const double q0 = inv_square(u[b * binsize + k]); w += q0; a = fma(x[k], q0, a); }
        else a += x[k];
    }
    if (W) {
        const double sdv = sqrt(1.0 / w);
        bs[b] = sdv;
        bd[b] = a * sdv * sdv;
    } else {
        bd[b] = a / (double)binsize;
    }
}

static size_t tile_smem_bytes(int halo, int64_t nsmall) {
    const int lf = TT + halo + 1;                     // exclusive prefix: one more entry
    return (size_t)((lf + (lf >> 4)) + 2 + 2 * nsmall) * 8;
}

struct RmsLayout { int64_t nblk, nout, nsmall; int ys, rows, tile_ctas, halo; bool use_tile, need_prefix;
                   int64_t oP, oTot, oT2, oDev, oPart, oIg, oLohi, oStat, oLead, words; };

RmsLayout rms_layout(int64_t n, int64_t maxbins, int64_t binstep, bool allow_tile = true) {
    RmsLayout L;
    L.nblk = ceil_div64(n, PB);
    L.nout = (maxbins - 1) / binstep + 1;
    L.ys = (int)(n / 65536 < 1 ? 1 : (n / 65536 > YS ? YS : n / 65536));
    // bin sizes b = 1 + i*binstep <= BMAX go to the tile kernel (large series only)
    L.use_tile = allow_tile && n >= 262144;
    L.nsmall = 0;
    if (L.use_tile) {
        const int64_t bcap = maxbins < BMAX ? maxbins : BMAX;
        L.nsmall = (bcap - 1) / binstep + 1;
        if (L.nsmall > L.nout) L.nsmall = L.nout;
    }
    const int64_t bsmall = L.nsmall > 0 ? 1 + (L.nsmall - 1) * binstep : 1;
    L.halo = (int)(((bsmall - 1) + 1) & ~1LL);                 // even: whole 16-byte units
    int sms = mc3b_sm_count();
    if (sms <= 0) sms = 148;
    const size_t tile_smem = tile_smem_bytes(L.halo, L.nsmall);
    int per_sm = (int)((227 * 1024) / (tile_smem + 1024));          // resident CTAs: shared memory ...
    if (per_sm > 2048 / (TNW * 32)) per_sm = 2048 / (TNW * 32);       // ... and threads
    if (per_sm < 1) per_sm = 1;
    if (const char* e = getenv("MC3B_TILE_CTAS")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;
    L.tile_ctas = sms * per_sm;
    const int64_t ntiles = ceil_div64(n, TT);
    if (L.tile_ctas > ntiles) L.tile_ctas = (int)ntiles;
    L.need_prefix = L.nsmall < L.nout;
    L.rows = L.use_tile ? (L.tile_ctas > L.ys ? L.tile_ctas : L.ys) : L.ys;
    int64_t o = 0;
    L.oP = o; o += n;                 // always reserved: the launch may fall back to the prefix path
    L.oTot = o; o += L.nblk;
    L.oT2 = o; o += L.nblk + 1;
    L.oDev = o; o += L.nblk;
    L.oPart = o; o += (int64_t)L.rows * L.nout;
    L.oIg = o; o += (int64_t)36 * 3 * IGN;
    L.oLohi = o; o += 72;
    L.oStat = o; o += 2;
    L.oLead = o; o += 18;          // 36 ints
    L.words = o;
    return L;
}

}  // namespace

extern "C" int64_t mc3b_binrms_workspace(int64_t n, int64_t maxbins, int64_t binstep) {
    if (n <= 0 || binstep <= 0) return 0;
    if (maxbins < 0) maxbins = n / 2;
    if (maxbins < 1) return 0;
    return rms_layout(n, maxbins, binstep).words * 8;
}

extern "C" int mc3b_binrms(const double* data, int64_t n, int64_t maxbins, int64_t binstep, void* workspace,
                           double* rms, double* rmslo, double* rmshi, double* stderr_, double* binsz, void* stream) {
    MC3B_CHECK_ARG(data && workspace && rms && rmslo && rmshi && stderr_ && binsz, "null pointer");
    if (maxbins < 0) maxbins = n / 2;
    MC3B_CHECK_ARG(n > 0 && binstep > 0 && maxbins >= 1 && maxbins <= n, "bad sizes (n=%lld maxbins=%lld binstep=%lld)",
                   (long long)n, (long long)maxbins, (long long)binstep);
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = ((uintptr_t)data & 15) == 0;     // the tile kernel's bulk copies need it
    const RmsLayout L = rms_layout(n, maxbins, binstep, aligned);
    double* w = (double*)workspace;
    double *P = w + L.oP, *tot = w + L.oTot, *T2 = w + L.oT2, *dev = w + L.oDev, *part = w + L.oPart;
    double *ig = w + L.oIg, *lohi = w + L.oLohi, *stat = w + L.oStat;
    int* lead = (int*)(w + L.oLead);
    if (L.need_prefix) {
        k_block_prefix<true><<<(unsigned)L.nblk, 256, 0, st>>>(data, n, P, tot);
        MC3B_CHECK_LAUNCH("k_block_prefix");
        k_totals_scan<<<1, 256, 0, st>>>(tot, L.nblk, n, T2, stat);
        MC3B_CHECK_LAUNCH("k_totals_scan");
        k_dev2<<<(unsigned)L.nblk, 256, 0, st>>>(data, n, stat, dev);
        MC3B_CHECK_LAUNCH("k_dev2");
        k_std<<<1, 256, 0, st>>>(dev, L.nblk, n, stat);
        MC3B_CHECK_LAUNCH("k_std");
    } else {                                             // mean and std in one streaming pass
        k_moments<<<(unsigned)L.nblk, 256, 0, st>>>(data, n, tot, dev);
        MC3B_CHECK_LAUNCH("k_moments");
        k_moments_finish<<<1, 256, 0, st>>>(tot, dev, L.nblk, n, stat);
        MC3B_CHECK_LAUNCH("k_moments_finish");
    }
    MC3B_CUDA(cudaMemsetAsync(part, 0, sizeof(double) * (size_t)L.rows * L.nout, st));
    if (L.nsmall > 0) {
        const size_t smem = tile_smem_bytes(L.halo, L.nsmall);
        MC3B_CUDA(cudaFuncSetAttribute(k_binrms_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_binrms_tile<<<(unsigned)L.tile_ctas, TNW * 32, smem, st>>>(data, n, L.nsmall, binstep, L.halo, L.nout, part);
        MC3B_CHECK_LAUNCH("k_binrms_tile");
    }
    if (L.need_prefix) {
        const int64_t nbig = L.nout - L.nsmall;
        const unsigned gx = (unsigned)(nbig < (1 << 20) ? nbig : (1 << 20));
        k_binrms_main<<<dim3(gx, (unsigned)L.ys), 256, 0, st>>>(P, T2, n, L.nsmall, L.nout, binstep, part);
        MC3B_CHECK_LAUNCH("k_binrms_main");
    }
    k_fill_int<<<1, 64, 0, st>>>(lead, 36, -1);
    MC3B_CHECK_LAUNCH("k_fill_int");
    k_binrms_finish<<<(unsigned)ceil_div64(L.nout, 128), 128, 0, st>>>(part, L.rows, n, L.nout, binstep, stat, rms,
                                                                        rmslo, rmshi, stderr_, binsz, lead);
    MC3B_CHECK_LAUNCH("k_binrms_finish");
    if (n / (1 + (L.nout - 1) * binstep) <= 35) {        // some bin size has <= 35 bins
        k_invgamma<<<36, 256, 0, st>>>(lead, stderr_, ig, lohi);
        MC3B_CHECK_LAUNCH("k_invgamma");
        k_binrms_small<<<(unsigned)ceil_div64(L.nout, 128), 128, 0, st>>>(n, L.nout, binstep, lohi, rms, stderr_,
                                                                           rmslo, rmshi);
        MC3B_CHECK_LAUNCH("k_binrms_small");
    }
    return MC3B_OK;
}

extern "C" int mc3b_binarray(const double* data, int64_t n, int64_t binsize, const double* uncert, double* bindata,
                             double* binstd, void* stream) {
    MC3B_CHECK_ARG(data && bindata && n > 0 && binsize > 0, "bad arguments");
    MC3B_CHECK_ARG(uncert == nullptr || binstd != nullptr, "weighted binning needs binstd");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nbins = n / binsize;
    if (nbins == 0) return MC3B_OK;
    if (binsize < 32) {
        const unsigned grid = (unsigned)ceil_div64(nbins, 256);
        if (uncert) k_binarray_short<true><<<grid, 256, 0, st>>>(data, uncert, nbins, (int)binsize, bindata, binstd);
        else k_binarray_short<false><<<grid, 256, 0, st>>>(data, uncert, nbins, (int)binsize, bindata, binstd);
        MC3B_CHECK_LAUNCH("k_binarray_short");
        return MC3B_OK;
    }
    if (binsize <= 8192) {
        int sms = mc3b_sm_count();
        int64_t grid = ceil_div64(nbins, 8);               // 8 warps (bins) per CTA
        const int64_t cap = (int64_t)sms * 8 * 16;        // grid-stride beyond 16 waves
        if (grid > cap) grid = cap;
        if (binsize <= 128) {
            // short bins: one pass of 4 loads per lane covers a bin; several bins per warp
            // and pass keep the loads in flight, at a register count (occupancy) that lets
            // other warps load while this one adds (the weighted kernel at 8 loads x 2 bins
            // x 2 arrays held 100 registers: 16 warps per SM, 45% of HBM)
            if (uncert) {
                grid = ceil_div64(nbins, 16);
                if (grid > cap) grid = cap;
                const char* e = getenv("MC3B_BA_MODE");           // A/B switch: bins per warp and pass
                const int mode = e ? atoi(e) : 1;                 // 1 bin per warp and pass: 0.299 ms at config 4 (2: 0.355, 4: 0.70)
                if (mode == 1) {
                    grid = ceil_div64(nbins, 8);
                    if (grid > cap) grid = cap;
                    k_binarray_direct<true, 1, 4><<<(unsigned)grid, 256, 0, st>>>(data, uncert, nbins, binsize, bindata, binstd);
                } else if (mode == 2) {
                    grid = ceil_div64(nbins, 32);
                    if (grid > cap) grid = cap;
                    k_binarray_direct<true, 4, 4><<<(unsigned)grid, 256, 0, st>>>(data, uncert, nbins, binsize, bindata, binstd);
                } else
                k_binarray_direct<true, 2, 4><<<(unsigned)grid, 256, 0, st>>>(data, uncert, nbins, binsize, bindata, binstd);
            } else {
                grid = ceil_div64(nbins, 32);
                if (grid > cap) grid = cap;
                k_binarray_direct<false, 4, 4><<<(unsigned)grid, 256, 0, st>>>(data, uncert, nbins, binsize, bindata, binstd);
            }
        } else {
            if (uncert) k_binarray_direct<true, 1, 8><<<(unsigned)grid, 256, 0, st>>>(data, uncert, nbins, binsize, bindata, binstd);
            else k_binarray_direct<false, 1, 8><<<(unsigned)grid, 256, 0, st>>>(data, uncert, nbins, binsize, bindata, binstd);
        }
        MC3B_CHECK_LAUNCH("k_binarray_direct");
        return MC3B_OK;
    }
    if (uncert) k_binarray_big<true><<<(unsigned)nbins, 256, 0, st>>>(data, uncert, binsize, bindata, binstd);
    else k_binarray_big<false><<<(unsigned)nbins, 256, 0, st>>>(data, uncert, binsize, bindata, binstd);
    MC3B_CHECK_LAUNCH("k_binarray");
    return MC3B_OK;
}
