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
// binarray: HBM-bound streaming; a CTA stages a contiguous run of whole bins in
// shared memory with coalesced loads, one warp reduces each bin.
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

// P = block-local inclusive prefix; tot[b] = block total.
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
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (base + k < n) P[base + k] = v[k] + excl;
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

__device__ __forceinline__ double range_sum(const double* P, const double* T2, int64_t s, int64_t e) {
    const int64_t bs = s / PB, be = (e - 1) / PB;
    const double head = (s % PB) ? P[s - 1] : 0.0;
    if (bs == be) return P[e - 1] - head;
    return (P[bs * PB + PB - 1] - head) + (T2[be] - T2[bs + 1]) + P[e - 1];
}

// grid (bin-size index, split): partial[y, i] = sum over this split's bins of mean^2.
__global__ void __launch_bounds__(256) k_binrms_main(const double* P, const double* T2, int64_t n, int64_t nout,
                                                    int64_t binstep, double* partial) {
    __shared__ double sh[9];
    for (int64_t i = blockIdx.x; i < nout; i += gridDim.x) {
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

// One thread per bin size: rms, asymptotic errors, Gaussian extrapolation, and
// the table lead[M] = first output index whose bin count is M (M <= 35).
__global__ void k_binrms_finish(const double* partial, int ys, int64_t n, int64_t nout, int64_t binstep,
                                const double* stat, double* rms, double* rmslo, double* rmshi, double* err,
                                double* binsz, int* lead) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nout) return;
    const int64_t b = 1 + i * binstep, M = n / b;
    double a = 0.0;
    for (int y = 0; y < ys; y++) a += partial[(int64_t)y * nout + i];
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
constexpr int BA_CHUNK = 4096;      // elements staged per CTA

template <bool W>
__global__ void __launch_bounds__(256) k_binarray_small(const double* d, const double* u, int64_t nbins,
                                                       int64_t binsize, int bpc, double* bd, double* bs) {
    constexpr int CHK = W ? BA_CHUNK / 2 : BA_CHUNK;
    __shared__ double sd[CHK];
    __shared__ double sw[W ? CHK : 1];
    const int64_t bin0 = (int64_t)blockIdx.x * bpc;
    const int nb = (int)((nbins - bin0) < bpc ? (nbins - bin0) : bpc);
    const int64_t e0 = bin0 * binsize;
    const int ne = nb * (int)binsize;
    for (int k = threadIdx.x; k < ne; k += 256) {
        if (W) {
            const double s = u[e0 + k], w = 1.0 / (s * s);
            sw[k] = w;
            sd[k] = d[e0 + k] * w;
        } else {
            sd[k] = d[e0 + k];
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = warp; b < nb; b += 8) {
        double a = 0.0, w = 0.0;
        for (int k = lane; k < (int)binsize; k += 32) {
            a += sd[b * (int)binsize + k];
            if (W) w += sw[b * (int)binsize + k];
        }
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            if (W) w += __shfl_xor_sync(0xffffffffu, w, o);
        }
        if (lane == 0) {
            if (W) {
                const double sdv = sqrt(1.0 / w);
                bs[bin0 + b] = sdv;
                bd[bin0 + b] = a * sdv * sdv;
            } else {
                bd[bin0 + b] = a / (double)binsize;
            }
        }
    }
}

template <bool W>
__global__ void __launch_bounds__(256) k_binarray_big(const double* d, const double* u, int64_t binsize, double* bd,
                                                     double* bs) {
    __shared__ double sh[9];
    const int64_t e0 = (int64_t)blockIdx.x * binsize;
    double a = 0.0, w = 0.0;
    for (int64_t k = threadIdx.x; k < binsize; k += 256) {
        if (W) {
            const double s = u[e0 + k], ww = 1.0 / (s * s);
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

struct RmsLayout { int64_t nblk, nout; int ys; int64_t oP, oTot, oT2, oDev, oPart, oIg, oLohi, oStat, oLead, words; };

RmsLayout rms_layout(int64_t n, int64_t maxbins, int64_t binstep) {
    RmsLayout L;
    L.nblk = ceil_div64(n, PB);
    L.nout = (maxbins - 1) / binstep + 1;
    L.ys = (int)(n / 65536 < 1 ? 1 : (n / 65536 > YS ? YS : n / 65536));
    int64_t o = 0;
    L.oP = o; o += n;
    L.oTot = o; o += L.nblk;
    L.oT2 = o; o += L.nblk + 1;
    L.oDev = o; o += L.nblk;
    L.oPart = o; o += (int64_t)L.ys * L.nout;
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
    const RmsLayout L = rms_layout(n, maxbins, binstep);
    double* w = (double*)workspace;
    double *P = w + L.oP, *tot = w + L.oTot, *T2 = w + L.oT2, *dev = w + L.oDev, *part = w + L.oPart;
    double *ig = w + L.oIg, *lohi = w + L.oLohi, *stat = w + L.oStat;
    int* lead = (int*)(w + L.oLead);
    k_block_prefix<<<(unsigned)L.nblk, 256, 0, st>>>(data, n, P, tot);
    MC3B_CHECK_LAUNCH("k_block_prefix");
    k_totals_scan<<<1, 256, 0, st>>>(tot, L.nblk, n, T2, stat);
    MC3B_CHECK_LAUNCH("k_totals_scan");
    k_dev2<<<(unsigned)L.nblk, 256, 0, st>>>(data, n, stat, dev);
    MC3B_CHECK_LAUNCH("k_dev2");
    k_std<<<1, 256, 0, st>>>(dev, L.nblk, n, stat);
    MC3B_CHECK_LAUNCH("k_std");
    const unsigned gx = (unsigned)(L.nout < (1 << 20) ? L.nout : (1 << 20));
    k_binrms_main<<<dim3(gx, (unsigned)L.ys), 256, 0, st>>>(P, T2, n, L.nout, binstep, part);
    MC3B_CHECK_LAUNCH("k_binrms_main");
    k_fill_int<<<1, 64, 0, st>>>(lead, 36, -1);
    MC3B_CHECK_LAUNCH("k_fill_int");
    k_binrms_finish<<<(unsigned)ceil_div64(L.nout, 128), 128, 0, st>>>(part, L.ys, n, L.nout, binstep, stat, rms,
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
    const int64_t chunk = uncert ? BA_CHUNK / 2 : BA_CHUNK;
    if (binsize <= chunk / 2) {
        const int bpc = (int)(chunk / binsize);
        const unsigned grid = (unsigned)ceil_div64(nbins, bpc);
        if (uncert) k_binarray_small<true><<<grid, 256, 0, st>>>(data, uncert, nbins, binsize, bpc, bindata, binstd);
        else k_binarray_small<false><<<grid, 256, 0, st>>>(data, uncert, nbins, binsize, bpc, bindata, binstd);
    } else {
        if (uncert) k_binarray_big<true><<<(unsigned)nbins, 256, 0, st>>>(data, uncert, binsize, bindata, binstd);
        else k_binarray_big<false><<<(unsigned)nbins, 256, 0, st>>>(data, uncert, binsize, bindata, binstd);
    }
    MC3B_CHECK_LAUNCH("k_binarray");
    return MC3B_OK;
}
