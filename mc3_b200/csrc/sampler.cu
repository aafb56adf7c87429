// mc3_b200 -- the per-generation sampler kernels (replace the body of
// Chain.run(), mc3/chain.py:183-299, for the whole population at once).
//
//   k_propose      one thread per chain: draws, jump, bounds, shared fill
//   k_metropolis   one thread per chain: partial sums + priors, accept/reject,
//                  counters, per-chain best, thinned history write
//   k_init_trials  initial-population trial points (mcmc_driver.py:229-262)
//   k_gr_*         Gelman-Rubin on device (gelman.py:36-92)
//
// Jump arithmetic uses explicit round-to-nearest mul/add/sub intrinsics (no FMA
// contraction) so that, fed the reference's recorded draws (replay mode), the
// proposed points equal numpy's elementwise results.
#include "common.cuh"

namespace {

constexpr int MAXP = MC3B_MAX_PARS;

struct Draws {                      // what one chain consumes in one generation
    int64_t a, b, iz;
    double usj, gs, u;
};

__device__ __forceinline__ void box_muller(double u1, double u2, double& n0, double& n1) {
    const double r = sqrt(-2.0 * log(1.0 - u1));     // 1-u1 in (0,1]
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    n0 = r * c;
    n1 = r * s;
}

template <bool REPLAY>
__global__ void __launch_bounds__(128) k_propose(mc3b_sampler_t S, mc3b_draws_t D, int64_t gen, int64_t zsize,
                                                 int64_t c_begin, int64_t c_end) {
    const int64_t c = c_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_end) return;
    if (gen < 0) {                                   // device-driven generation (graph mode)
        gen = *S.gen_dev;
        zsize = S.M0 + (gen / S.thinning) * S.nchains;
    }
    const int nfree = S.nfree, npars = S.npars;
    const double* x = S.X + c * nfree;
    double jump[MAXP], nrm[MAXP];
    Draws dr;
    dr.iz = -1; dr.usj = 1.0; dr.gs = 0.0;

    if (REPLAY) {
        for (int j = 0; j < nfree; j++) nrm[j] = D.normal[j];
        dr.a = dr.b = 0;
        if (S.sampler != MC3B_MRW) { dr.a = D.a[c]; dr.b = D.b[c]; }
        if (S.sampler == MC3B_SNOOKER) { dr.iz = D.iz[c]; dr.usj = D.usj[c]; dr.gs = D.gs[c]; }
        dr.u = D.u[c];
    } else {
        const Philox ph(S.seed);
        const uint32_t cid = (uint32_t)c, g0 = (uint32_t)gen, g1 = (uint32_t)((uint64_t)gen >> 32);
        const uint4 w0 = ph(cid, 0u, g0, g1), w1 = ph(cid, 1u, g0, g1), w2 = ph(cid, 2u, g0, g1);
        if (S.sampler == MC3B_DEMC) {               // chain.py:223-229
            int64_t r1 = 1 + ubelow(w0.x, w0.y, S.nchains - 1);
            if (r1 == c) r1 = 0;
            int64_t r2 = (r1 + 2 + ubelow(w0.z, w0.w, S.nchains - 2)) % S.nchains;
            if (r2 == c) r2 = (r1 + 1) % S.nchains;
            dr.a = r1; dr.b = r2;
        } else if (S.sampler == MC3B_SNOOKER) {     // chain.py:197-203
            int64_t i1 = ubelow(w0.x, w0.y, zsize);
            int64_t i2 = 1 + ubelow(w0.z, w0.w, zsize - 1);
            if (i2 == i1) i2 = 0;
            dr.a = i1; dr.b = i2;
            dr.usj = u01(w1.x, w1.y);
            dr.gs = 1.2 + u01(w1.z, w1.w);
            dr.iz = ubelow(w2.x, w2.y, zsize);
        } else {
            dr.a = dr.b = 0;
        }
        dr.u = u01(w2.z, w2.w);
        for (int j = 0; j < nfree; j += 2) {        // per-chain support draw
            const uint4 w = ph(cid, 3u + (uint32_t)(j >> 1), g0, g1);
            double n0, n1;
            box_muller(u01(w.x, w.y), u01(w.z, w.w), n0, n1);
            nrm[j] = n0 * S.pstep[S.ifree[j]];
            if (j + 1 < nfree) nrm[j + 1] = n1 * S.pstep[S.ifree[j + 1]];
        }
    }

    double mrfactor = 1.0;
    bool sjump = false;
    const double* zrow = nullptr;
    if (S.sampler == MC3B_SNOOKER) {
        const double* z1 = S.Z + dr.a * nfree;
        const double* z2 = S.Z + dr.b * nfree;
        sjump = dr.usj < 0.1;
        if (sjump) {                                 // chain.py:202-213
            zrow = S.Z + dr.iz * nfree;
            bool same = true;
            for (int j = 0; j < nfree; j++) same = same && (zrow[j] == x[j]);
            if (same) {
                for (int j = 0; j < nfree; j++) jump[j] = __dmul_rn(dr.gs, __dsub_rn(z2[j], z1[j]));
            } else {
                double zp1 = 0.0, zp2 = 0.0, dd = 0.0;
                for (int j = 0; j < nfree; j++) {
                    const double dz = __dsub_rn(x[j], zrow[j]);
                    zp1 = __dadd_rn(zp1, __dmul_rn(z1[j], dz));
                    zp2 = __dadd_rn(zp2, __dmul_rn(z2[j], dz));
                    dd = __dadd_rn(dd, __dmul_rn(dz, dz));
                }
                const double f = __dmul_rn(dr.gs, __dsub_rn(zp1, zp2));
                for (int j = 0; j < nfree; j++)
                    jump[j] = __ddiv_rn(__dmul_rn(f, __dsub_rn(x[j], zrow[j])), dd);
            }
        } else {                                     // chain.py:214-217
            for (int j = 0; j < nfree; j++)
                jump[j] = __dadd_rn(__dmul_rn(S.gamma, __dsub_rn(z1[j], z2[j])), __dmul_rn(S.fepsilon, nrm[j]));
        }
    } else if (S.sampler == MC3B_DEMC) {             // chain.py:230-232
        const double* x1 = S.X + dr.a * nfree;
        const double* x2 = S.X + dr.b * nfree;
        for (int j = 0; j < nfree; j++)
            jump[j] = __dadd_rn(__dmul_rn(S.gamma, __dsub_rn(x1[j], x2[j])), __dmul_rn(S.fepsilon, nrm[j]));
    } else {                                         // mrw, chain.py:219-220
        for (int j = 0; j < nfree; j++) jump[j] = nrm[j];
    }

    // chain.py:235-247 -- propose, bounds, shared parameters
    double* np_ = S.nextp + c * npars;
    for (int k = 0; k < npars; k++) np_[k] = S.params0[k];
    int inb = 1;
    for (int j = 0; j < nfree; j++) {
        const int k = S.ifree[j];
        double v = __dadd_rn(x[j], jump[j]);
        if (S.reflect) {                             // opt-in, non-reference behaviour
            const double lo = S.pmin[k], hi = S.pmax[k];
            for (int it = 0; it < 8 && (v < lo || v > hi); it++) v = v < lo ? 2.0 * lo - v : 2.0 * hi - v;
        }
        np_[k] = v;
        if (v < S.pmin[k] || v > S.pmax[k]) {
            inb = 0;
            atomicAdd(&S.outbounds[j], 1);
        }
    }
    for (int k = 0; k < npars; k++)
        if (S.pstep[k] < 0.0) np_[k] = np_[-(int)S.pstep[k] - 1];
    if (sjump && inb) {                              // chain.py:251-255
        double cn = 0.0, nn = 0.0;
        for (int j = 0; j < nfree; j++) {
            const double dc = __dsub_rn(x[j], zrow[j]), dn = __dsub_rn(np_[S.ifree[j]], zrow[j]);
            cn = __dadd_rn(cn, __dmul_rn(dc, dc));
            nn = __dadd_rn(nn, __dmul_rn(dn, dn));
        }
        mrfactor = pow(nn / cn, 0.5 * (nfree - 1));
    }
    S.mrfactor[c] = mrfactor;
    S.u[c] = dr.u;
    S.inb[c] = inb;
}

__global__ void __launch_bounds__(128) k_metropolis(mc3b_sampler_t S, const double* partial, int64_t ldpartial, int nsplit,
                                                    int64_t c_off, int64_t gen, int64_t zrow0, int64_t c_begin,
                                                    int64_t c_end) {
    const int64_t c = c_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_end) return;
    if (gen < 0) {                                   // device-driven generation (graph mode)
        gen = *S.gen_dev;
        zrow0 = ((gen + 1) % S.thinning == 0) ? S.M0 + ((gen + 1) / S.thinning - 1) * S.nchains : -1;
    }
    const int nfree = S.nfree, npars = S.npars;
    double* x = S.X + c * nfree;
    double cur = S.chisq_cur[c];
    if (S.inb[c]) {
        const double* np_ = S.nextp + c * npars;
        double nxt = 0.0;
        for (int s = 0; s < nsplit; s++) nxt += partial[(int64_t)s * ldpartial + (c - c_off)];
        if (S.prior != nullptr) {                    // stats.py:208-216 + stats.h:90-109
            double pr = 0.0;
            for (int k = 0; k < npars; k++) {
                const double lo = S.priorlow[k], up = S.priorup[k];
                if (lo > 0.0 && up > 0.0) {
                    const double off = np_[k] - S.prior[k];
                    const double t = off / (off > 0.0 ? up : lo);
                    pr += t * t;
                }
            }
            nxt += pr;
        }
        const double ratio = exp(0.5 * (cur - nxt)) * S.mrfactor[c];
        if (ratio > S.u[c]) {                        // chain.py:257-274 (NaN rejects)
            for (int j = 0; j < nfree; j++) x[j] = np_[S.ifree[j]];
            cur = nxt;
            S.chisq_cur[c] = nxt;
            S.naccept[c] += 1;
            if (nxt < S.best_chisq[c]) {
                S.best_chisq[c] = nxt;
                S.best_gen[c] = gen;
                for (int j = 0; j < nfree; j++) S.best_x[c * nfree + j] = x[j];
            }
        }
    }
    if (zrow0 >= 0) {                                // chain.py:276-289
        const int64_t row = zrow0 + c;
        if (row < S.zlen) {
            for (int j = 0; j < nfree; j++) S.Z[row * nfree + j] = x[j];
            S.log_post[row] = -0.5 * cur;
            S.zchain[row] = (int32_t)c;
        }
    }
}

__global__ void __launch_bounds__(128) k_init_trials(mc3b_sampler_t S, int kickoff, int64_t ntrials, int64_t round,
                                                     double* trial, int32_t* ok) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntrials) return;
    const Philox ph(S.seed);
    double* v = trial + t * S.npars;
    for (int k = 0; k < S.npars; k++) v[k] = S.params0[k];
    int good = 1;
    for (int j = 0; j < S.nfree; j += 2) {
        const uint4 w = ph((uint32_t)t, (uint32_t)(j >> 1), (uint32_t)round, 0xFFFFFFFFu);
        double d0, d1;
        if (kickoff == 0) box_muller(u01(w.x, w.y), u01(w.z, w.w), d0, d1);
        else { d0 = u01(w.x, w.y); d1 = u01(w.z, w.w); }
        for (int q = 0; q < 2 && j + q < S.nfree; q++) {
            const int k = S.ifree[j + q];
            const double d = q ? d1 : d0;
            v[k] = kickoff == 0 ? S.params0[k] + S.pstep[k] * d : S.pmin[k] + (S.pmax[k] - S.pmin[k]) * d;
        }
    }
    for (int k = 0; k < S.npars; k++)
        if (S.pstep[k] < 0.0) v[k] = v[-(int)S.pstep[k] - 1];
    for (int k = 0; k < S.npars; k++)
        if (v[k] > S.pmax[k] || v[k] < S.pmin[k]) good = 0;     // mcmc_driver.py:252
    ok[t] = good;
}

__global__ void k_advance(int64_t* gen_dev) { *gen_dev += 1; }

// log_prior of history rows (mc3/stats/stats.py:367-392) and the data chi-squared
// it implies: lpr = -0.5 sum_j t_j^2 with t_j = (z-prior)/low|up for Gaussian
// priors (low, up > 0), t_j = 2 log z where priorlow < 0; chisq = -2 (log_post - lpr).
__global__ void __launch_bounds__(128) k_log_prior(const double* Z, int64_t nrows, int nfree, const int32_t* ifree,
                                                   const double* prior, const double* plo, const double* pup,
                                                   const double* log_post, double* lpr, double* chisq) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    double acc = 0.0;
    for (int j = 0; j < nfree; j++) {
        const int k = ifree[j];
        const double z = Z[r * nfree + j], lo = plo[k], up = pup[k];
        double t = 0.0;
        if (lo > 0.0 && up > 0.0) {
            const double d = z - prior[k];
            t = d < 0.0 ? d / lo : (d > 0.0 ? d / up : d);
        } else if (lo < 0.0) {
            t = 2.0 * log(z);
        }
        acc += t * t;
    }
    const double v = -0.5 * acc;
    if (lpr) lpr[r] = v;
    if (chisq) chisq[r] = -2.0 * (log_post[r] - v);
}

// ---- Gelman-Rubin ----------------------------------------------------------
// Stage 1: one thread per (chain, parameter): mean and population variance of
// its niter samples (two passes, as numpy's var).  Stage 2: one CTA, fixed-order
// sums over chains -> W, B, V, sqrt(V/W)   (gelman.py:75-92).
__global__ void __launch_bounds__(128) k_gr_stats(const double* Z, int64_t nfree, int64_t nchains, int64_t M0,
                                                  const int64_t* rows, int64_t ldr, int64_t burnin, int64_t niter,
                                                  double* work) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nchains * nfree) return;
    const int64_t c = t / nfree, p = t % nfree;
    auto at = [&](int64_t k) -> double {
        const int64_t row = rows ? rows[c * ldr + k] : M0 + k * nchains + c;
        return Z[row * nfree + p];
    };
    double s = 0.0;
    for (int64_t k = burnin; k < burnin + niter; k++) s += at(k);
    const double mu = s / (double)niter;
    double v = 0.0;
    for (int64_t k = burnin; k < burnin + niter; k++) { const double d = at(k) - mu; v += d * d; }
    work[t] = mu;
    work[nchains * nfree + t] = v / (double)niter;
}

__global__ void __launch_bounds__(256) k_gr_psrf(const double* work, int64_t nfree, int64_t nchains, int64_t niter,
                                                 double* psrf) {
    __shared__ double sh[256];
    const int p = blockIdx.x;
    const double* mu = work;
    const double* var = work + nchains * nfree;
    auto block_sum = [&](double v) -> double {
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        const double r = sh[0];
        __syncthreads();
        return r;
    };
    double a = 0.0, b = 0.0;
    for (int64_t c = threadIdx.x; c < nchains; c += 256) { a += var[c * nfree + p]; b += mu[c * nfree + p]; }
    const double W = block_sum(a) / (double)nchains;
    const double mm = block_sum(b) / (double)nchains;
    double d2 = 0.0;
    for (int64_t c = threadIdx.x; c < nchains; c += 256) { const double d = mu[c * nfree + p] - mm; d2 += d * d; }
    const double B = (double)niter / ((double)nchains - 1.0) * block_sum(d2);
    if (threadIdx.x == 0) {
        const double V = W * (((double)niter - 1.0) / (double)niter) +
                         B * (((double)nchains + 1.0) / ((double)niter * (double)nchains));
        psrf[p] = sqrt(V / W);
    }
}

int check_sampler(const mc3b_sampler_t* s, int64_t c_begin, int64_t c_end) {
    MC3B_CHECK_ARG(s != nullptr, "null sampler");
    MC3B_CHECK_ARG(s->nfree > 0 && s->nfree <= MAXP && s->npars >= s->nfree && s->npars <= MAXP,
                   "nfree/npars out of range (max %d)", MAXP);
    MC3B_CHECK_ARG(c_begin >= s->chain0 && c_end <= s->chain0 + s->nlocal && c_begin <= c_end,
                   "chain range outside this device's slice");
    MC3B_CHECK_ARG(s->sampler == MC3B_MRW || s->sampler == MC3B_DEMC || s->sampler == MC3B_SNOOKER,
                   "unknown sampler %d", s->sampler);
    return MC3B_OK;
}

}  // namespace

extern "C" int mc3b_propose(const mc3b_sampler_t* s, int64_t gen, int64_t zsize, int64_t c_begin, int64_t c_end,
                            void* stream) {
    if (int rc = check_sampler(s, c_begin, c_end)) return rc;
    MC3B_CHECK_ARG(s->sampler != MC3B_DEMC || s->nchains >= 3, "demc needs at least 3 chains");
    MC3B_CHECK_ARG(gen >= 0 || (s->gen_dev && s->thinning > 0), "graph mode needs gen_dev and thinning");
    MC3B_CHECK_ARG(s->sampler != MC3B_SNOOKER || gen < 0 || zsize >= 2, "snooker needs at least 2 history rows");
    if (c_end == c_begin) return MC3B_OK;
    mc3b_draws_t none = {};
    k_propose<false><<<(unsigned)ceil_div64(c_end - c_begin, 128), 128, 0, (cudaStream_t)stream>>>(
        *s, none, gen, zsize, c_begin, c_end);
    MC3B_CHECK_LAUNCH("k_propose");
    return MC3B_OK;
}

extern "C" int mc3b_propose_replay(const mc3b_sampler_t* s, const mc3b_draws_t* d, int64_t c_begin, int64_t c_end,
                                   void* stream) {
    if (int rc = check_sampler(s, c_begin, c_end)) return rc;
    MC3B_CHECK_ARG(d && d->normal && d->u, "replay draws missing");
    MC3B_CHECK_ARG(s->sampler == MC3B_MRW || (d->a && d->b), "replay partner indices missing");
    MC3B_CHECK_ARG(s->sampler != MC3B_SNOOKER || (d->iz && d->usj && d->gs), "replay snooker draws missing");
    if (c_end == c_begin) return MC3B_OK;
    k_propose<true><<<(unsigned)ceil_div64(c_end - c_begin, 128), 128, 0, (cudaStream_t)stream>>>(
        *s, *d, 0, 0, c_begin, c_end);
    MC3B_CHECK_LAUNCH("k_propose<replay>");
    return MC3B_OK;
}

extern "C" int mc3b_metropolis(const mc3b_sampler_t* s, const double* partial, int64_t ldpartial, int nsplit,
                               int64_t c_off, int64_t gen, int64_t zrow0, int64_t c_begin, int64_t c_end,
                               void* stream) {
    if (int rc = check_sampler(s, c_begin, c_end)) return rc;
    MC3B_CHECK_ARG(partial && nsplit > 0 && c_off <= c_begin && ldpartial >= c_end - c_off, "bad partial workspace");
    MC3B_CHECK_ARG(gen >= 0 || (s->gen_dev && s->thinning > 0), "graph mode needs gen_dev and thinning");
    if (c_end == c_begin) return MC3B_OK;
    k_metropolis<<<(unsigned)ceil_div64(c_end - c_begin, 128), 128, 0, (cudaStream_t)stream>>>(
        *s, partial, ldpartial, nsplit, c_off, gen, zrow0, c_begin, c_end);
    MC3B_CHECK_LAUNCH("k_metropolis");
    return MC3B_OK;
}

extern "C" int mc3b_advance(const mc3b_sampler_t* s, void* stream) {
    MC3B_CHECK_ARG(s && s->gen_dev, "no device generation counter");
    k_advance<<<1, 1, 0, (cudaStream_t)stream>>>(s->gen_dev);
    MC3B_CHECK_LAUNCH("k_advance");
    return MC3B_OK;
}

extern "C" int mc3b_log_prior(const double* Z, int64_t nrows, int nfree, const int32_t* ifree, const double* prior,
                              const double* priorlow, const double* priorup, const double* log_post, double* lpr,
                              double* chisq, void* stream) {
    MC3B_CHECK_ARG(Z && ifree && prior && priorlow && priorup && nrows > 0 && nfree > 0, "bad arguments");
    MC3B_CHECK_ARG((lpr || chisq) && (!chisq || log_post), "need an output (chisq needs log_post)");
    k_log_prior<<<(unsigned)ceil_div64(nrows, 128), 128, 0, (cudaStream_t)stream>>>(Z, nrows, nfree, ifree, prior,
                                                                                     priorlow, priorup, log_post,
                                                                                     lpr, chisq);
    MC3B_CHECK_LAUNCH("k_log_prior");
    return MC3B_OK;
}

extern "C" int mc3b_init_trials(const mc3b_sampler_t* s, int kickoff, int64_t ntrials, int64_t round,
                                double* trial, int32_t* ok, void* stream) {
    MC3B_CHECK_ARG(s && trial && ok && ntrials > 0, "bad arguments");
    MC3B_CHECK_ARG(s->nfree > 0 && s->nfree <= MAXP && s->npars <= MAXP, "nfree/npars out of range");
    MC3B_CHECK_ARG(kickoff == 0 || kickoff == 1, "kickoff must be 0 (normal) or 1 (uniform)");
    k_init_trials<<<(unsigned)ceil_div64(ntrials, 128), 128, 0, (cudaStream_t)stream>>>(*s, kickoff, ntrials, round,
                                                                                        trial, ok);
    MC3B_CHECK_LAUNCH("k_init_trials");
    return MC3B_OK;
}

extern "C" int mc3b_gelman_rubin(const double* Z, int64_t nfree, int64_t nchains, int64_t M0, const int64_t* rows,
                                 int64_t ldr, int64_t burnin, int64_t niter, double* work, double* psrf,
                                 void* stream) {
    MC3B_CHECK_ARG(Z && work && psrf && nfree > 0 && nchains > 1 && niter > 0 && burnin >= 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    k_gr_stats<<<(unsigned)ceil_div64(nchains * nfree, 128), 128, 0, st>>>(Z, nfree, nchains, M0, rows, ldr, burnin,
                                                                           niter, work);
    MC3B_CHECK_LAUNCH("k_gr_stats");
    k_gr_psrf<<<(unsigned)nfree, 256, 0, st>>>(work, nfree, nchains, niter, psrf);
    MC3B_CHECK_LAUNCH("k_gr_psrf");
    return MC3B_OK;
}
