// mc3_b200 -- persistent kernel for small populations.
//
// The reference's everyday use (7-21 chains, 1e2-1e4 data points, ~1e5
// samples) is launch-latency bound when every generation is three kernel
// launches: the arithmetic of a generation takes well under a microsecond.
// k_run_small keeps one CTA resident and runs `ngen` generations inside it:
//
//     threads 0..nchains-1   propose_chain()            (sampler_dev.cuh)
//     __syncthreads
//     one warp per chain     model + chi-squared over the data (lanes stride the
//                            points, fixed-order lane tree)
//     __syncthreads
//     threads 0..nchains-1   metropolis_chain()         (priors, accept, write)
//     __syncthreads
//
// Same Philox streams, same proposal and acceptance code as the per-generation
// kernels; only the order in which a chain's chi-squared terms are added
// differs (rounding-level).  Bound: latency (one CTA by construction).
#include "models.cuh"
#include "sampler_dev.cuh"

namespace {

constexpr int SW = 8;       // warps

template <class M>
__global__ void __launch_bounds__(SW * 32) k_run_small(mc3b_sampler_t S, const double* __restrict__ x,
                                                       const double* __restrict__ d,
                                                       const double* __restrict__ w, int64_t n,
                                                       double* chisq_new, int64_t gen0, int64_t ngen) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nch = (int)S.nchains;
    const mc3b_draws_t none = {};
    for (int64_t g = gen0; g < gen0 + ngen; g++) {
        const int64_t zsize = S.M0 + (g / S.thinning) * S.nchains;
        const int64_t zrow0 = ((g + 1) % S.thinning == 0) ? S.M0 + ((g + 1) / S.thinning - 1) * S.nchains : -1;
        if ((int)threadIdx.x < nch) propose_chain<false>(S, none, g, zsize, threadIdx.x);
        __syncthreads();
        for (int c = warp; c < nch; c += SW) {
            if (S.inb[c]) {                              // chain.py:241: no evaluation out of bounds
                M m;
                m.load(S.nextp + (int64_t)c * S.npars);
                double a0 = 0.0, a1 = 0.0;
                int64_t i = lane;
                for (; i + 32 < n; i += 64) {
                    const double r0 = (m.eval_safe(x[i]) - d[i]) * w[i];
                    const double r1 = (m.eval_safe(x[i + 32]) - d[i + 32]) * w[i + 32];
                    a0 = fma(r0, r0, a0);
                    a1 = fma(r1, r1, a1);
                }
                if (i < n) { const double r0 = (m.eval_safe(x[i]) - d[i]) * w[i]; a0 = fma(r0, r0, a0); }
                double a = a0 + a1;
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) chisq_new[c] = a;
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < nch) metropolis_chain(S, chisq_new, nch, 1, 0, g, zrow0, threadIdx.x);
        __syncthreads();
    }
    if (threadIdx.x == 0 && S.gen_dev) *S.gen_dev = gen0 + ngen;
}

}  // namespace

extern "C" int mc3b_run_small(const mc3b_sampler_t* s, int model_id, int nmodel, const double* x, const double* data,
                              const double* invsig, int64_t n, double* scratch, int64_t gen0, int64_t ngen,
                              void* stream) {
    MC3B_CHECK_ARG(s && x && data && invsig && scratch && n > 0 && gen0 >= 0 && ngen > 0, "bad arguments");
    MC3B_CHECK_ARG(s->nchains <= SW * 32 && s->chain0 == 0 && s->nlocal == s->nchains,
                   "run_small handles one device and at most %d chains", SW * 32);
    MC3B_CHECK_ARG(s->nfree > 0 && s->nfree <= MC3B_MAX_PARS && s->npars <= MC3B_MAX_PARS && s->thinning > 0,
                   "nfree/npars/thinning out of range");
    MC3B_CHECK_ARG(s->sampler != MC3B_DEMC || s->nchains >= 3, "demc needs at least 3 chains");
    MC3B_CHECK_ARG(s->sampler != MC3B_SNOOKER || s->M0 >= 2, "snooker needs at least 2 history rows");
    if (model_id == MC3B_MODEL_SINUSOID_GRID) model_id = MC3B_MODEL_SINUSOID;
    MC3B_CHECK_ARG(mc3b_model_nparams(model_id, nmodel) == nmodel, "model %d does not take %d parameters", model_id,
                   nmodel);
    cudaStream_t st = (cudaStream_t)stream;
    MC3B_DISPATCH_MODEL(double, model_id, nmodel,
                        (k_run_small<M><<<1, SW * 32, 0, st>>>(*s, x, data, invsig, n, scratch, gen0, ngen)));
    MC3B_CHECK_LAUNCH("k_run_small");
    return MC3B_OK;
}
