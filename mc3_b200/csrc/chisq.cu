// mc3_b200 -- chi-squared family (replaces src_c/_chisq.c and the model +
// chi-squared half of Chain.eval_model, mc3/chain.py:302-340).
//
// k_model_chisq is the hot kernel of the sampler: one launch evaluates the
// built-in model and the data chi-squared of EVERY chain of the population.
//
//   grid  = (chain groups, data splits); block = 4 warps.
//   warp  = LC chains x LP = 32/LC point lanes.  LC = 32: one chain per lane,
//           every lane walks all points of a tile (shared-memory broadcast);
//           LC = 8 / 1: fewer chains, lanes share the points of the tile.
//   data  = x, data, 1/sigma tiles of TILE points, staged into shared memory
//           by 1-D TMA bulk copies (cp.async.bulk + mbarrier, 3 stages), so a
//           tile is fetched once per CTA and serves all its chains; the
//           per-chain constants live in registers for the whole launch.
//   sum   = per lane in registers, then a fixed-order xor-shuffle over the
//           point lanes, one partial per (split, chain): deterministic.
//
// Bound: FP64 (or FP32) pipe -- nchains*F flops per 24 bytes of data.
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <type_traits>
#include "models.cuh"
#include "sampler_dev.cuh"

#include "chisq_args.cuh"

namespace {

// USIG: one uncertainty for all points (w[0]): residuals are plain differences
// (a two-register DADD instead of a multiply or a three-register FMA) and the sum
// is scaled by w[0]^2 once per partial.
template <class M, typename T, int LC, int CPT, bool USIG>
__global__ void __launch_bounds__(WARPS * 32, (CPT == 2 ? MC3B_RESIDENT2 : RESIDENT)) k_model_chisq(ChisqArgs<T> a) {
    constexpr int TILE = tilecfg<T>::TILE;
    constexpr int LP = 32 / LC;
    asm volatile("griddepcontrol.launch_dependents;");   // the next proposal kernel may start its draws
    __shared__ __align__(128) T sx[STAGES][TILE];
    __shared__ __align__(128) T sd[STAGES][TILE];
    __shared__ __align__(128) T sw[STAGES][TILE];
    __shared__ __align__(8) uint64_t bar[STAGES];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lc = lane % LC, lp = lane / LC;
    const int64_t cbase = ((int64_t)blockIdx.x * WARPS + warp) * (LC * CPT) + lc;

    M mdl[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        int64_t c = cbase + (int64_t)k * LC;
        if (c >= a.nchains) c = a.nchains - 1;      // idle lanes shadow the last chain
        if constexpr (M::TILE_STATE) {
            const double dx = ((double)a.x[a.n - 1] - (double)a.x[0]) / (double)(a.n - 1);
            mdl[k].load(a.params + c * a.ldp, dx, TILE);
        } else {
            mdl[k].load(a.params + c * a.ldp);
        }
    }
    double acc[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) acc[k] = 0.0;

    // Tiles of this split: balanced contiguous ranges of FULL tiles; the
    // ragged tail (n % TILE points) belongs to the last split.
    const int64_t nfull = a.n / TILE;
    int64_t tb, te;
    if (a.nsched > 0) { tb = a.tstart[blockIdx.y]; te = a.tstart[blockIdx.y + 1]; }
    else { tb = nfull * blockIdx.y / gridDim.y; te = nfull * (blockIdx.y + 1) / gridDim.y; }
    const int64_t nt = te - tb;

    if (a.use_tma) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < STAGES; s++) mbar_init(&bar[s], 1);
            mbar_fence_init();
        }
        __syncthreads();
        auto issue = [&](int64_t t, int s) {
            const int64_t off = t * TILE;
            mbar_expect_tx(&bar[s], (USIG ? 2u : 3u) * TILE * sizeof(T));
            bulk_g2s(sx[s], a.x + off, TILE * sizeof(T), &bar[s]);
            bulk_g2s(sd[s], a.d + off, TILE * sizeof(T), &bar[s]);
            if constexpr (!USIG) bulk_g2s(sw[s], a.w + off, TILE * sizeof(T), &bar[s]);
        };
        if (threadIdx.x == 0)
            for (int s = 0; s < STAGES && s < nt; s++) issue(tb + s, s);
        constexpr bool PM = premul_of<M>::value && !USIG;
        // PM keeps d/sigma in a buffer of its own: TMA never writes it, so no generic-proxy
        // store ever meets an async-proxy refill of the same bytes (no proxy fence needed)
        __shared__ __align__(128) T sdw[PM ? STAGES : 1][PM ? TILE : 2];
        // PM: data tile <- data / sigma, once for all chains of the CTA.  Tile it+1 is
        // scaled at the end of tile it, so that the stage-release barrier publishes it.
        auto premul = [&](int64_t t) {
            const int s1 = (int)(t % STAGES);
            mbar_wait(&bar[s1], (uint32_t)((t / STAGES) & 1));
            for (int i = threadIdx.x; i < TILE; i += WARPS * 32) sdw[PM ? s1 : 0][PM ? i : 0] = sd[s1][i] * sw[s1][i];
        };
        if constexpr (PM) {
            if (nt > 0) premul(0);
            __syncthreads();
        }
        for (int64_t it = 0; it < nt; it++) {
            const int s = (int)(it % STAGES);
            if constexpr (!PM) mbar_wait(&bar[s], (uint32_t)((it / STAGES) & 1));
            constexpr int U = 4;                // points in flight per lane
            static_assert(TILE % (U * LP) == 0, "tile must hold whole groups");
            if constexpr (M::TILE_STATE) {
                static_assert(LP == 1, "stateful models walk a tile in order: one chain per lane");
#pragma unroll
                for (int k = 0; k < CPT; k++) mdl[k].begin_tile((double)sx[s][0], TILE);
            }
            T tacc[CPT], uacc[CPT][U];          // one accumulator per point slot: no serial tail
#pragma unroll
            for (int k = 0; k < CPT; k++) {
#pragma unroll
                for (int u = 0; u < U; u++) uacc[k][u] = (T)0;
            }
#pragma unroll(LOOP_UNROLL)
            for (int i = lp; i < TILE; i += U * LP) {
                T x[U], y[U];
#pragma unroll
                for (int u = 0; u < U; u++) x[u] = sx[s][i + u * LP];
#pragma unroll
                for (int k = 0; k < CPT; k++) {
                    eval_points<M, T, U>(mdl[k], x, y, 0);
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const T r = USIG ? y[u] - sd[s][i + u * LP]
                                    : PM ? fma(y[u], sw[s][i + u * LP], -sdw[PM ? s : 0][PM ? i + u * LP : 0])
                                         : (y[u] - sd[s][i + u * LP]) * sw[s][i + u * LP];
                        uacc[k][u] = fma(r, r, uacc[k][u]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < CPT; k++) tacc[k] = (uacc[k][0] + uacc[k][1]) + (uacc[k][2] + uacc[k][3]);
            if (M::GUARD) {                     // extreme model arguments: redo the tile safely
#pragma unroll
                for (int k = 0; k < CPT; k++) {
                    if (mdl[k].flagged()) {
                        mdl[k].clear();
                        tacc[k] = (T)0;
                        for (int i = lp; i < TILE; i += LP) {
                            const T r = USIG ? mdl[k].eval_safe(sx[s][i]) - sd[s][i]
                                        : PM ? fma(mdl[k].eval_safe(sx[s][i]), sw[s][i], -sdw[PM ? s : 0][PM ? i : 0])
                                             : (mdl[k].eval_safe(sx[s][i]) - sd[s][i]) * sw[s][i];
                            tacc[k] = fma(r, r, tacc[k]);
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < CPT; k++) acc[k] += (double)tacc[k];
            if constexpr (PM) {
                if (it + 1 < nt) premul(it + 1);
            }
            __syncthreads();                    // everyone is done reading stage s
            if (threadIdx.x == 0 && it + STAGES < nt) issue(tb + it + STAGES, s);
        }
    } else {
        for (int64_t it = 0; it < nt; it++) {   // unaligned inputs: plain staged loads
            const int64_t off = (tb + it) * TILE;
            __syncthreads();
            for (int i = threadIdx.x; i < TILE; i += WARPS * 32) {
                sx[0][i] = a.x[off + i]; sd[0][i] = a.d[off + i]; sw[0][i] = USIG ? (T)1 : a.w[off + i];
            }
            __syncthreads();
            T tacc[CPT];
#pragma unroll
            for (int k = 0; k < CPT; k++) tacc[k] = (T)0;
            for (int i = lp; i < TILE; i += LP) {
                const T x = sx[0][i], d = sd[0][i], w = sw[0][i];
#pragma unroll
                for (int k = 0; k < CPT; k++) {
                    const T r = (mdl[k].eval_safe(x) - d) * w;
                    tacc[k] = fma(r, r, tacc[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < CPT; k++) acc[k] += (double)tacc[k];
        }
    }

    // Ragged tail, straight from global memory (last split only).
    if (blockIdx.y == gridDim.y - 1) {
        T tacc[CPT];
#pragma unroll
        for (int k = 0; k < CPT; k++) tacc[k] = (T)0;
        for (int64_t i = nfull * TILE + lp; i < a.n; i += LP) {
            const T x = a.x[i], d = a.d[i], w = USIG ? (T)1 : a.w[i];
#pragma unroll
            for (int k = 0; k < CPT; k++) {
                const T r = (mdl[k].eval_safe(x) - d) * w;
                tacc[k] = fma(r, r, tacc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < CPT; k++) acc[k] += (double)tacc[k];
    }

    double w2 = 1.0;
    if constexpr (USIG) { const double w0 = (double)a.w[0]; w2 = w0 * w0; }
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o >= LC; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int64_t c = cbase + (int64_t)k * LC;
        if (lp == 0 && c < a.nchains) a.partial[(int64_t)blockIdx.y * a.ldpartial + c] = USIG ? v * w2 : v;
    }
    if (a.f.on) fused_metropolis(a.f, a.partial, a.ldpartial, a.nchains, WARPS * LC * CPT);
}

// Shape policy: a pure function of (plan_chains, n, dtype, SM count).  plan_chains
// is the chain count the shape is planned FOR -- by default the chains of the
// launch; a caller that wants identical chi-squared bits however the chains are
// spread over launches or devices plans for the whole population and launches any
// subset of it with that plan (the split boundaries, hence the order in which a
// chain's terms are added, then do not depend on the subset).
struct Shape {
    int lc, nsplit;
    int nsched; int32_t tstart[SCHED_MAX + 1];
};

// kind: MC3B_PLAN_GENERAL, or MC3B_PLAN_MOMENT for the sufficient-statistics kernel: it runs 4
// CTAs per SM and its CTAs are short, so its waves are planned for 4 and its splits shrink
// more slowly and stop at 8 tiles (measured at config 2: 0.0954 against 0.0985 ms; the same plan
// costs the other kernels 3-16 %, profiles/r2_plan_ab.txt).
Shape plan_shape(int64_t nchains, int64_t n, int dtype, int sms, int kind = MC3B_PLAN_GENERAL) {
    Shape s;
    int RESIDENT = kind == MC3B_PLAN_MOMENT ? 4 : mc3b_chisq::RESIDENT;    // MC3B_PLAN_RESIDENT: experiments
    if (const char* e = getenv("MC3B_PLAN_RESIDENT")) RESIDENT = atoi(e) > 0 ? atoi(e) : RESIDENT;
    s.lc = nchains >= 96 ? 32 : (nchains >= 12 ? 8 : 1);
    const int tile = dtype == MC3B_F32 ? tilecfg<float>::TILE : tilecfg<double>::TILE;
    const int64_t groups = ceil_div64(nchains, (int64_t)WARPS * s.lc);
    const int64_t nfull = n / tile;
    int waves = 2;
    if (const char* e = getenv("MC3B_WAVES")) waves = atoi(e) > 0 ? atoi(e) : 1;
    // whole waves of RESIDENT CTAs per SM (round to nearest: a few CTAs over one
    // wave cost a whole extra wave, so round down unless clearly closer to the next)
    int64_t want = ((int64_t)sms * RESIDENT * waves) / groups;
    int64_t ns = want < 1 ? 1 : want;
    if (ns > nfull) ns = nfull;
    if (ns > MC3B_MAX_SPLIT) ns = MC3B_MAX_SPLIT;
    if (ns < 1) ns = 1;
    s.nsplit = (int)ns;
    s.nsched = 0;
    // Decreasing split sizes ("factoring"): CTAs are dispatched in split order, so the
    // last ones to start are short and the SMs run dry together.  Each batch of
    // ~one wave of splits takes 1/f of the tiles that remain.
    // f = 2, at least 4 tiles per split measured best at config 2 (0.203 -> 0.186 ms;
    // profiles/r1_split_schedule.md); MC3B_SCHED="f,min" overrides, "0" gives equal splits.
    double f = kind == MC3B_PLAN_MOMENT ? 1.5 : 2.0; int mn = kind == MC3B_PLAN_MOMENT ? 8 : 4;
    if (const char* e = getenv("MC3B_SCHED")) { f = 0.0; sscanf(e, "%lf,%d", &f, &mn); }
    if (mn < 1) mn = 1;
    if (f >= 1.0 && s.lc == 32 && nfull < (1 << 30) && ns >= 8) {
        const double rows = (double)sms * RESIDENT / (double)groups;
        const int per = (int)(rows + 0.999);
        int64_t R = nfull; int k = 0; bool fits = true;
        s.tstart[0] = 0;
        while (R > 0 && fits) {
            int64_t sz = (int64_t)((double)R / (f * rows) + 0.5);
            if (sz < mn) sz = mn;
            for (int j = 0; j < per && R > 0; j++) {
                const int64_t t = sz < R ? sz : R;
                if (k >= SCHED_MAX) { fits = false; break; }
                s.tstart[k + 1] = (int32_t)(s.tstart[k] + t);
                k++; R -= t;
            }
        }
        if (fits && k >= 1) { s.nsched = k; s.nsplit = k; }
    }
    return s;
}

template <class M, typename T, bool USIG>
int launch_model_chisq(const Shape& sh, const ChisqArgs<T>& a, dim3 grid, cudaStream_t st) {
    dim3 block(WARPS * 32);
    if (sh.lc == 32) k_model_chisq<M, T, 32, 1, USIG><<<grid, block, 0, st>>>(a);
    else if (sh.lc == 8) k_model_chisq<M, T, 8, 1, USIG><<<grid, block, 0, st>>>(a);
    else k_model_chisq<M, T, 1, 1, USIG><<<grid, block, 0, st>>>(a);
    MC3B_CHECK_LAUNCH("k_model_chisq");
    return MC3B_OK;
}

template <typename T>
int model_chisq_t(int model_id, const double* params, int64_t ldp, int64_t nchains, int nmodel,
                  const void* x, const void* d, const void* w, int64_t n, double* partial, int64_t ldpartial,
                  int nsplit, int dtype, const mc3b_chisq_opts_t& o, cudaStream_t st) {
    ChisqArgs<T> A;
    A.params = params; A.ldp = ldp; A.nchains = nchains;
    A.x = (const T*)x; A.d = (const T*)d; A.w = (const T*)w; A.fold = (const T*)o.folded; A.consts = nullptr; A.ldc = 0; A.consts_wait = 0; A.m = MomentArgs{}; A.n = n;
    A.xt = nullptr; A.dxg = 0.0; A.ntiles = 0; A.partial = partial; A.ldpartial = ldpartial;
    const bool usig = o.uniform_sigma != 0;
    A.use_tma = ((((uintptr_t)x | (uintptr_t)d | (usig ? 0 : (uintptr_t)w)) & 15) == 0) ? 1 : 0;
    const int64_t plan_chains = o.plan_chains > 0 ? o.plan_chains : nchains;
    MC3B_CHECK_ARG(plan_chains >= nchains, "plan_chains (%lld) is smaller than the launch (%lld chains)",
                   (long long)plan_chains, (long long)nchains);
    const Shape sh = plan_shape(plan_chains, n, dtype, mc3b_sm_count(),
                                o.moment != nullptr && usig ? MC3B_PLAN_MOMENT : MC3B_PLAN_GENERAL);
    A.nsched = sh.nsched;
    if (sh.nsched > 0) memcpy(A.tstart, sh.tstart, sizeof(int32_t) * (sh.nsched + 1));
    MC3B_CHECK_ARG(nsplit == sh.nsplit, "nsplit %d does not match the plan (%d)", nsplit, sh.nsplit);
    const int64_t groups = ceil_div64(nchains, (int64_t)WARPS * sh.lc);
    MC3B_CHECK_ARG(groups <= 0x7fffffff, "too many chains for one launch");
    A.f.on = 0;
    if (o.fuse != nullptr) {
        MC3B_CHECK_ARG(o.fuse_done != nullptr, "the fused Metropolis epilogue needs its counters (fuse_done)");
        MC3B_CHECK_ARG(o.gen >= 0 || (o.fuse->gen_dev && o.fuse->thinning > 0),
                       "device-driven generations need gen_dev and thinning");
        MC3B_CHECK_ARG(!o.advance || o.fuse->gen_dev, "advance needs gen_dev");
        MC3B_CHECK_ARG(o.c_off >= o.fuse->chain0 && o.c_off + nchains <= o.fuse->chain0 + o.fuse->nlocal,
                       "chain range outside this device's slice");
        A.f.on = 1; A.f.advance = o.advance; A.f.done = o.fuse_done; A.f.c_off = o.c_off;
        A.f.gen = o.gen; A.f.zrow0 = o.zrow0; A.f.S = *o.fuse;
    }
    dim3 grid((unsigned)groups, (unsigned)nsplit), block(WARPS * 32);
    if (model_id == MC3B_MODEL_SINUSOID_GRID) {
        if constexpr (std::is_same<T, double>::value) {
            if (sh.lc == 32 && A.use_tma && n >= 2) {
                if (o.tile_x != nullptr) {
                    MC3B_CHECK_ARG(o.dx != 0.0 && o.ntiles >= 0 && o.ntiles * 128 <= n, "tile_x needs dx and ntiles <= n/128");
                    A.xt = o.tile_x; A.dxg = o.dx; A.ntiles = o.ntiles;
                }
                if (usig && o.moment != nullptr) {
                    MC3B_CHECK_ARG(o.work, "the moment form needs the constants workspace (work)");
                    MC3B_CHECK_ARG(o.moment->folded && o.moment->tiles && o.moment->amp_max > 0.0 &&
                                   (((uintptr_t)o.moment->folded | (uintptr_t)o.moment->tiles) & 31) == 0,
                                   "bad moment data");
                    A.m.folded = o.moment->folded; A.m.tiles = o.moment->tiles;
                    A.m.c0ref = o.moment->c0ref; A.m.slref = o.moment->slref; A.m.d2tot = o.moment->d2tot;
                    A.m.amp_max = o.moment->amp_max; A.m.guard_hits = o.moment->guard_hits;
                    A.m.xlo = o.moment->xlo; A.m.xhi = o.moment->xhi; A.m.layout = o.moment->layout;
                    MC3B_CHECK_ARG(A.m.layout == 0 || A.m.layout == 1, "bad moment layout %d", A.m.layout);
                    return mc3b_launch_sinefold(A, (double*)o.work, (unsigned)groups, (unsigned)nsplit, st);
                }
                if (usig && A.fold != nullptr && ((uintptr_t)A.fold & 15) == 0)
                    return mc3b_launch_sinefold(A, (double*)o.work, (unsigned)groups, (unsigned)nsplit, st);
                if (!getenv("MC3B_OLD_GRID"))
                    return mc3b_launch_sinegrid(A, usig, (unsigned)groups, (unsigned)nsplit, st);
                if (!usig) {                        // round-1 kernel, kept for A/B measurements
                    k_model_chisq<SineGridModel, double, 32, 1, false><<<grid, block, 0, st>>>(A);
                    MC3B_CHECK_LAUNCH("k_model_chisq<grid>");
                    return MC3B_OK;
                }
            }
        }
        model_id = MC3B_MODEL_SINUSOID;             // small populations, fp32, unaligned: plain model
    }
    if (usig) { MC3B_DISPATCH_MODEL(T, model_id, nmodel, return (launch_model_chisq<M, T, true>(sh, A, grid, st))); }
    else { MC3B_DISPATCH_MODEL(T, model_id, nmodel, return (launch_model_chisq<M, T, false>(sh, A, grid, st))); }
    return MC3B_OK;
}

// ---- model values (best_model, func(params) shape probe) -------------------
template <class M>
__global__ void k_model_eval(const double* params, int64_t ldp, const double* x, int64_t n, double* out) {
    M m;
    m.load(params + (int64_t)blockIdx.y * ldp);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[(int64_t)blockIdx.y * n + i] = m.eval_safe(x[i]);
}

// ---- sum of partials + priors ----------------------------------------------
__global__ void k_chisq_finish(const double* partial, int64_t ldpartial, int nsplit, int64_t nchains, const double* params,
                               int64_t ldp, int npars, const double* prior, const double* plo,
                               const double* pup, double* chisq) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchains) return;
    double acc = 0.0;
    for (int s = 0; s < nsplit; s++) acc += partial[(int64_t)s * ldpartial + c];
    if (prior != nullptr) {
        double pr = 0.0;
        for (int j = 0; j < npars; j++) {
            const double lo = plo[j], up = pup[j];
            if (lo > 0.0 && up > 0.0) {
                const double off = params[c * ldp + j] - prior[j];
                const double t = off / (off > 0.0 ? up : lo);
                pr += t * t;
            }
        }
        acc += pr;
    }
    chisq[c] = acc;
}

// ---- chi-squared of given models (user callables): HBM-bound ---------------
// One CTA per chain row; 256 threads stride the row (coalesced 8-byte loads),
// fixed-order reduction (lane tree, then warps in order).
__global__ void __launch_bounds__(256) k_chisq_rows(const double* model, int64_t ldm, const double* data,
                                                    const double* uncert, int64_t n, double* chisq) {
    const double* m = model + (int64_t)blockIdx.x * ldm;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) {
        const double r = (m[i] - data[i]) / uncert[i];
        acc = fma(r, r, acc);
    }
    __shared__ double ws[8];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; k++) t += ws[k];
        chisq[blockIdx.x] = t;
    }
}

__global__ void k_residuals(const double* model, const double* data, const double* uncert, int64_t n,
                            const double* off, const double* low, const double* up, int64_t np, double* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (model[i] - data[i]) / uncert[i];
    else if (i < n + np) {
        const int64_t k = i - n;
        out[i] = off[k] / (off[k] > 0.0 ? up[k] : low[k]);
    }
}

}  // namespace

extern "C" int mc3b_model_chisq_plan_kind(int kind, int64_t nchains, int64_t n, int dtype, int* nsplit) {
    MC3B_CHECK_ARG(nchains > 0 && n > 0 && nsplit != nullptr, "bad plan arguments");
    MC3B_CHECK_ARG(dtype == MC3B_F64 || dtype == MC3B_F32, "bad dtype %d", dtype);
    MC3B_CHECK_ARG(kind == MC3B_PLAN_GENERAL || kind == MC3B_PLAN_MOMENT, "bad plan kind %d", kind);
    int sms = mc3b_sm_count();
    if (sms <= 0) sms = 148;        // planning without a device (CPU-side sizing)
    *nsplit = plan_shape(nchains, n, dtype, sms, kind).nsplit;
    return MC3B_OK;
}

extern "C" int mc3b_model_chisq_plan(int64_t nchains, int64_t n, int dtype, int* nsplit) {
    return mc3b_model_chisq_plan_kind(MC3B_PLAN_GENERAL, nchains, n, dtype, nsplit);
}

extern "C" int mc3b_model_chisq_splits(int64_t nchains, int64_t n, int dtype, int64_t* point_start, int cap,
                                       int* nsplit) {
    MC3B_CHECK_ARG(nchains > 0 && n > 0 && nsplit != nullptr && point_start != nullptr, "bad plan arguments");
    MC3B_CHECK_ARG(dtype == MC3B_F64 || dtype == MC3B_F32, "bad dtype %d", dtype);
    int sms = mc3b_sm_count();
    if (sms <= 0) sms = 148;
    const Shape sh = plan_shape(nchains, n, dtype, sms);
    MC3B_CHECK_ARG(cap >= sh.nsplit + 1, "point_start holds %d entries, the plan needs %d", cap, sh.nsplit + 1);
    const int64_t tile = dtype == MC3B_F32 ? tilecfg<float>::TILE : tilecfg<double>::TILE;
    const int64_t nfull = n / tile;
    for (int y = 0; y <= sh.nsplit; y++)
        point_start[y] = tile * (sh.nsched > 0 ? (int64_t)sh.tstart[y] : nfull * y / sh.nsplit);
    point_start[sh.nsplit] = n;                       // the ragged tail belongs to the last split
    *nsplit = sh.nsplit;
    return MC3B_OK;
}

extern "C" int mc3b_model_chisq_ex(int model_id, int dtype, const double* params, int64_t ldp, int64_t nchains,
                                   int nmodel, const void* x, const void* data, const void* invsig, int64_t n,
                                   double* partial, int64_t ldpartial, int nsplit, const mc3b_chisq_opts_t* opts,
                                   void* stream) {
    MC3B_CHECK_ARG(params && x && data && invsig && partial, "null pointer");
    MC3B_CHECK_ARG(ldpartial >= nchains, "ldpartial < nchains");
    MC3B_CHECK_ARG(nchains > 0 && n > 0 && nmodel > 0 && nmodel <= ldp, "bad sizes");
    MC3B_CHECK_ARG(mc3b_model_nparams(model_id, nmodel) == nmodel, "model %d does not take %d parameters",
                   model_id, nmodel);
    mc3b_chisq_opts_t o = {};
    if (opts) o = *opts;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MC3B_F64)
        return model_chisq_t<double>(model_id, params, ldp, nchains, nmodel, x, data, invsig, n, partial, ldpartial,
                                     nsplit, dtype, o, st);
    if (dtype == MC3B_F32)
        return model_chisq_t<float>(model_id, params, ldp, nchains, nmodel, x, data, invsig, n, partial, ldpartial,
                                    nsplit, dtype, o, st);
    mc3b_set_error("bad dtype %d", dtype);
    return MC3B_ERR_ARG;
}

extern "C" int mc3b_fold_data(const double* data, int64_t n, double* out, void* stream) {
    MC3B_CHECK_ARG(data && out && n > 0, "bad fold arguments");
    return mc3b_launch_fold(data, n, out, (cudaStream_t)stream);
}

extern "C" int mc3b_moment_finish(const mc3b_moment_t* m, const double* partial, int64_t ldpartial, int nsplit,
                                  int64_t nchains, const double* params, int64_t ldp, int npars, const double* x,
                                  const double* data, const double* invsig, int64_t n, const double* prior,
                                  const double* priorlow, const double* priorup, double* chisq, void* stream) {
    MC3B_CHECK_ARG(m && partial && params && x && data && invsig && chisq, "null pointer");
    MC3B_CHECK_ARG(nchains > 0 && nsplit > 0 && n > 0 && ldpartial >= nchains && ldp >= 5 && m->amp_max > 0.0,
                   "bad moment_finish arguments");
    MC3B_CHECK_ARG(prior == nullptr || (priorlow && priorup && npars > 0 && npars <= ldp), "bad prior arguments");
    ChisqArgs<double> A = {};
    A.params = params; A.ldp = ldp; A.nchains = nchains; A.x = x; A.d = data; A.w = invsig; A.n = n;
    A.partial = const_cast<double*>(partial); A.ldpartial = ldpartial;
    A.m.folded = m->folded; A.m.tiles = m->tiles; A.m.c0ref = m->c0ref; A.m.slref = m->slref;
    A.m.d2tot = m->d2tot; A.m.amp_max = m->amp_max; A.m.xlo = m->xlo; A.m.xhi = m->xhi; A.m.guard_hits = m->guard_hits;
    return mc3b_launch_moment_finish(A, nsplit, npars, prior, priorlow, priorup, chisq, (cudaStream_t)stream);
}

extern "C" int mc3b_moment_prepare(const double* data, int64_t ntiles, double x0, double dx, const double* tile_x,
                                   double c0ref, double slref, int layout, double* folded, double* tiles, void* stream) {
    MC3B_CHECK_ARG(data && folded && tiles && ntiles >= 0, "bad moment_prepare arguments");
    MC3B_CHECK_ARG(layout == 0 || layout == 1, "bad moment layout %d", layout);
    MC3B_CHECK_ARG((((uintptr_t)tiles) & 31) == 0, "tiles must be 32-byte aligned");
    return mc3b_launch_moment_prepare(data, ntiles, x0, dx, tile_x, c0ref, slref, folded, tiles, layout, (cudaStream_t)stream);
}

extern "C" int mc3b_model_chisq(int model_id, int dtype, const double* params, int64_t ldp, int64_t nchains,
                                int nmodel, const void* x, const void* data, const void* invsig, int64_t n,
                                double* partial, int64_t ldpartial, int nsplit, void* stream) {
    return mc3b_model_chisq_ex(model_id, dtype, params, ldp, nchains, nmodel, x, data, invsig, n, partial,
                               ldpartial, nsplit, nullptr, stream);
}

extern "C" int mc3b_model_eval(int model_id, const double* params, int64_t ldp, int64_t nchains, int nmodel,
                               const double* x, int64_t n, double* out, void* stream) {
    if (model_id == MC3B_MODEL_SINUSOID_GRID) model_id = MC3B_MODEL_SINUSOID;
    MC3B_CHECK_ARG(params && x && out && nchains > 0 && n > 0, "bad arguments");
    MC3B_CHECK_ARG(mc3b_model_nparams(model_id, nmodel) == nmodel, "model %d does not take %d parameters",
                   model_id, nmodel);
    MC3B_CHECK_ARG(nchains <= 65535, "model_eval is for small batches (<= 65535 rows)");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t bx = ceil_div64(n, 256);
    if (bx > 1184) bx = 1184;
    dim3 grid((unsigned)bx, (unsigned)nchains);
    MC3B_DISPATCH_MODEL(double, model_id, nmodel, (k_model_eval<M><<<grid, 256, 0, st>>>(params, ldp, x, n, out)));
    MC3B_CHECK_LAUNCH("k_model_eval");
    return MC3B_OK;
}

extern "C" int mc3b_chisq_finish(const double* partial, int64_t ldpartial, int nsplit, int64_t nchains, const double* params,
                                 int64_t ldp, int npars, const double* prior, const double* priorlow,
                                 const double* priorup, double* chisq, void* stream) {
    MC3B_CHECK_ARG(partial && chisq && nsplit > 0 && nchains > 0, "bad arguments");
    MC3B_CHECK_ARG(prior == nullptr || (params && priorlow && priorup), "priors need params, priorlow, priorup");
    k_chisq_finish<<<(unsigned)ceil_div64(nchains, 128), 128, 0, (cudaStream_t)stream>>>(
        partial, ldpartial, nsplit, nchains, params, ldp, npars, prior, priorlow, priorup, chisq);
    MC3B_CHECK_LAUNCH("k_chisq_finish");
    return MC3B_OK;
}

extern "C" int mc3b_chisq_batch(const double* model, int64_t ldm, int64_t nchains, const double* data,
                                const double* uncert, int64_t n, double* chisq, void* stream) {
    MC3B_CHECK_ARG(model && data && uncert && chisq && nchains > 0 && n > 0 && ldm >= n, "bad arguments");
    k_chisq_rows<<<(unsigned)nchains, 256, 0, (cudaStream_t)stream>>>(model, ldm, data, uncert, n, chisq);
    MC3B_CHECK_LAUNCH("k_chisq_rows");
    return MC3B_OK;
}

extern "C" int mc3b_residuals(const double* model, const double* data, const double* uncert, int64_t n,
                              const double* off, const double* low, const double* up, int64_t np, double* out,
                              void* stream) {
    MC3B_CHECK_ARG(model && data && uncert && out && n > 0, "bad arguments");
    if (off == nullptr) np = 0;
    k_residuals<<<(unsigned)ceil_div64(n + np, 256), 256, 0, (cudaStream_t)stream>>>(model, data, uncert, n, off,
                                                                                      low, up, np, out);
    MC3B_CHECK_LAUNCH("k_residuals");
    return MC3B_OK;
}
