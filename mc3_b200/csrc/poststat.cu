// mc3_b200 -- highest-posterior-density statistics of the marginals on the device
// (replaces the host scipy path of mc3/stats/stats.py:433-467 cred_region and
// :764-802 marginal_statistics 'max_like', the step right after the sampling loop).
//
//   k_col_stats   one CTA per parameter: mean, population std, min, max of the
//                 column and the unbiased variance of its KDE subsample
//   k_kde_grid    grid (100, nfree): Gaussian kernel density (Scott's bandwidth, as
//                 scipy.stats.gaussian_kde) at 100 points inside +-6 sigma
//   k_hpd_finish  one CTA per parameter: linear resample to 3000 points, descending
//                 sort (bitonic, shared memory), running sum, density threshold at
//                 `quantile`, mode and the outermost points above the threshold
//
// Bound: latency (2e6 exponentials per parameter for a 20 000-row sample).
#include <math.h>
#include "common.cuh"

namespace {

constexpr int NGRID = 100, NFINE = 3000, NSORT = 4096;
// per parameter in `work`: [0] mean [1] std0 [2] min [3] max [4] kde variance (cov * factor^2)
// [5] lo [6] hi [7] n_kde ; then NGRID densities ; then the NFINE resampled densities
constexpr int WSTRIDE = 8 + NGRID + NFINE;

__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (t < o) sh[t] += sh[t + o];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256) k_col_stats(const double* post, int64_t n, int nfree, int64_t thin,
                                                   double* work) {
    __shared__ double sh[256];
    const int p = blockIdx.x, t = threadIdx.x;
    double s = 0.0, mn = INFINITY, mx = -INFINITY, sk = 0.0;
    for (int64_t i = t; i < n; i += 256) {
        const double v = post[i * nfree + p];
        s += v; mn = fmin(mn, v); mx = fmax(mx, v);
        if (i % thin == 0) sk += v;
    }
    const int64_t nk = (n + thin - 1) / thin;
    const double mean = block_sum(s, sh) / (double)n;
    const double meank = block_sum(sk, sh) / (double)nk;
    double d2 = 0.0, dk = 0.0;
    for (int64_t i = t; i < n; i += 256) {
        const double v = post[i * nfree + p];
        d2 += (v - mean) * (v - mean);
        if (i % thin == 0) dk += (v - meank) * (v - meank);
    }
    const double var0 = block_sum(d2, sh) / (double)n;
    const double var1 = block_sum(dk, sh) / (double)(nk - 1);
    sh[t] = mn; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (t < o) sh[t] = fmin(sh[t], sh[t + o]); __syncthreads(); }
    mn = sh[0]; __syncthreads();
    sh[t] = mx; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (t < o) sh[t] = fmax(sh[t], sh[t + o]); __syncthreads(); }
    mx = sh[0];
    if (t == 0) {
        double* w = work + (int64_t)p * WSTRIDE;
        const double sd = sqrt(var0);
        const double factor = pow((double)nk, -0.2);          // Scott: n^(-1/(d+4)), d = 1
        w[0] = mean; w[1] = sd; w[2] = mn; w[3] = mx;
        w[4] = var1 * factor * factor;
        w[5] = fmax(mean - 6.0 * sd, mn);                       // stats.py:447-450
        w[6] = fmin(mean + 6.0 * sd, mx);
        w[7] = (double)nk;
    }
}

// numpy.linspace(lo, hi, num)[i]
__device__ __forceinline__ double linspace_at(double lo, double hi, int num, int i) {
    if (i == num - 1) return hi;
    const double step = (hi - lo) / (double)(num - 1);
    return __dadd_rn(__dmul_rn((double)i, step), lo);
}

__global__ void __launch_bounds__(256) k_kde_grid(const double* post, int64_t n, int nfree, int64_t thin,
                                                  double* work) {
    __shared__ double sh[256];
    const int g = blockIdx.x, p = blockIdx.y;
    double* w = work + (int64_t)p * WSTRIDE;
    const double x = linspace_at(w[5], w[6], NGRID, g);
    const double hinv = -0.5 / w[4];
    double acc = 0.0;
    for (int64_t i = (int64_t)threadIdx.x * thin; i < n; i += 256 * thin) {
        const double d = x - post[i * nfree + p];
        acc += exp(d * d * hinv);
    }
    const double tot = block_sum(acc, sh);
    if (threadIdx.x == 0) w[8 + g] = tot / (sqrt(2.0 * M_PI * w[4]) * w[7]);
}

__global__ void __launch_bounds__(1024) k_hpd_finish(double* work, int nfree, double quantile, double* out) {
    __shared__ double srt[NSORT];
    __shared__ double s_hmin;
    const int p = blockIdx.x, t = threadIdx.x;
    double* w = work + (int64_t)p * WSTRIDE;
    double* pdf = w + 8 + NGRID;
    const double lo = w[5], hi = w[6];
    const double* y = w + 8;
    // scipy.interpolate.interp1d (linear): interval by searchsorted on the 100-point grid
    for (int i = t; i < NSORT; i += 1024) {
        double v = -INFINITY;
        if (i < NFINE) {
            const double xn = linspace_at(lo, hi, NFINE, i);
            int k = 1;
            while (k < NGRID - 1 && linspace_at(lo, hi, NGRID, k) < xn) k++;      // first grid point >= xn
            const double x0 = linspace_at(lo, hi, NGRID, k - 1), x1 = linspace_at(lo, hi, NGRID, k);
            const double slope = (y[k] - y[k - 1]) / (x1 - x0);
            v = slope * (xn - x0) + y[k - 1];
        }
        if (i < NFINE) pdf[i] = v;
        srt[i] = v;
    }
    __syncthreads();
    for (int k = 2; k <= NSORT; k <<= 1) {                     // bitonic sort, descending
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < NSORT; i += 1024) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const double a = srt[i], b = srt[l];
                    if (up ? a < b : a > b) { srt[i] = b; srt[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (t == 0) {                                              // stats.py:459-466
        double tot = 0.0;
        for (int i = 0; i < NFINE; i++) tot += srt[i];
        const double thr = quantile * tot;
        double c = 0.0;
        int ih = NFINE - 1;
        for (int i = 0; i < NFINE; i++) { c += srt[i]; if (c >= thr) { ih = i; break; } }
        s_hmin = ih > 0 ? srt[ih - 1] : INFINITY;
    }
    __syncthreads();
    const double hmin = s_hmin;
    // mode = first maximum; bounds = outermost fine-grid points with pdf > hmin
    int* s_lo = reinterpret_cast<int*>(srt);                  // the sorted copy is no longer needed
    int* s_hi = s_lo + 1024;
    int* s_am = s_hi + 1024;
    int ilo = NFINE, ihi = -1, iam = -1;
    double best = -INFINITY;
    for (int i = t; i < NFINE; i += 1024) {
        if (pdf[i] > hmin) { ilo = min(ilo, i); ihi = max(ihi, i); }
        if (pdf[i] > best) { best = pdf[i]; iam = i; }
    }
    s_lo[t] = ilo; s_hi[t] = ihi; s_am[t] = iam;
    __syncthreads();
    if (t == 0) {
        for (int k = 1; k < 1024; k++) {
            ilo = min(ilo, s_lo[k]); ihi = max(ihi, s_hi[k]);
            const int a = s_am[k];
            if (a >= 0 && (iam < 0 || pdf[a] > pdf[iam] || (pdf[a] == pdf[iam] && a < iam))) iam = a;
        }
        out[p] = linspace_at(lo, hi, NFINE, iam);
        out[nfree + p] = ihi >= 0 ? linspace_at(lo, hi, NFINE, ilo) : NAN;
        out[2 * nfree + p] = ihi >= 0 ? linspace_at(lo, hi, NFINE, ihi) : NAN;
    }
}

}  // namespace

extern "C" int64_t mc3b_hpd_workspace(int nfree) { return (int64_t)nfree * WSTRIDE * 8; }

extern "C" int mc3b_hpd(const double* posterior, int64_t n, int nfree, double quantile, void* workspace,
                        double* out, void* stream) {
    MC3B_CHECK_ARG(posterior && workspace && out && n >= 2 && nfree > 0, "bad arguments");
    MC3B_CHECK_ARG(quantile > 0.0 && quantile < 1.0, "quantile must be in (0, 1)");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t thin = n / 120000;                                  // stats.py:441
    if (thin < 1) thin = 1;
    double* w = (double*)workspace;
    k_col_stats<<<nfree, 256, 0, st>>>(posterior, n, nfree, thin, w);
    MC3B_CHECK_LAUNCH("k_col_stats");
    k_kde_grid<<<dim3(NGRID, nfree), 256, 0, st>>>(posterior, n, nfree, thin, w);
    MC3B_CHECK_LAUNCH("k_kde_grid");
    k_hpd_finish<<<nfree, 1024, 0, st>>>(w, nfree, quantile, out);
    MC3B_CHECK_LAUNCH("k_hpd_finish");
    return MC3B_OK;
}
