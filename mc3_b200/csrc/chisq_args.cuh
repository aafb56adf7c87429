// mc3_b200 -- launch parameters and the fused Metropolis epilogue shared by the
// model + chi-squared kernels (chisq.cu, chisq_grid.cu).
#pragma once
#include "models.cuh"
#include "sampler_dev.cuh"

namespace mc3b_chisq {

#ifndef MC3B_WARPS
#define MC3B_WARPS 4
#endif
#ifndef MC3B_RESIDENT
#define MC3B_RESIDENT 6
#endif
#ifndef MC3B_RESIDENT2
#define MC3B_RESIDENT2 4
#endif
#ifndef MC3B_LOOP_UNROLL
#define MC3B_LOOP_UNROLL 2
#endif
#ifndef MC3B_TILE_F64
#define MC3B_TILE_F64 128
#endif
constexpr int WARPS = MC3B_WARPS;
constexpr int STAGES = 3;
constexpr int LOOP_UNROLL = MC3B_LOOP_UNROLL;   // point groups of the inner loop unrolled together
constexpr int RESIDENT = MC3B_RESIDENT;   // CTAs per SM the register budget is tuned for (ILP over occupancy)
template <typename T> struct tilecfg { static constexpr int TILE = MC3B_TILE_F64; };   // fp64: 128 beat 256 by 6% (finer balance)
template <> struct tilecfg<float> { static constexpr int TILE = 512; };

constexpr int SCHED_MAX = 384;      // splits a size schedule can describe (else equal splits)
// Optional epilogue of the model kernels: the LAST CTA of a chain group (the one
// whose arrival completes the group's split count) adds the group's partial rows
// in split order and takes the Metropolis step of its chains -- what k_metropolis
// does, without the extra launch and without a second pass over `partial` from a
// cold kernel; the last group to finish bumps the device generation counter
// (replaces k_advance).  Same arithmetic, same order: identical bits.
struct FuseArgs {
    int on;
    int advance;                    // bump *S.gen_dev when every group is done
    int32_t* done;                  // [groups + 1] arrival counters; zero before the first launch, self-resetting
    int64_t c_off;                  // global id of the chain in row 0 of params
    int64_t gen, zrow0;             // as mc3b_metropolis (gen < 0: device-driven)
    mc3b_sampler_t S;
};

// k_sinefold<MOM>: data centred on a reference line, folded and scaled (mc3b_moment_prepare),
// the per-tile moments, and the guard of the expansion (include/mc3b200.h, mc3b_moment_t)
struct MomentArgs {
    const double* folded;           // nullptr: not the moment form
    const double* tiles;
    double c0ref, slref, d2tot, amp_max;
    double xlo, xhi;                // range of the abscissa (bound of the line term of the guard)
    int32_t* guard_hits;
    int layout;                     // of `folded`: 0 pairs interleaved (k_sinefold<MOM>), 1 mma fragments (k_sinemma)
};

template <typename T> struct ChisqArgs {
    int nsched;                     // > 0: split y covers tiles [tstart[y], tstart[y+1])
    int32_t tstart[SCHED_MAX + 1];
    const double* params;
    int64_t ldp, nchains;
    const T *x, *d, *w;
    const T* fold;                  // mirrored-pair copy of d (mc3b_fold_data) or nullptr
    // k_sinefold on a piecewise-uniform abscissa (include/mc3b200.h, tile_x): origin of every whole
    // 128-point tile, the common step, the number of whole tiles; x/d hold the tiles first, then
    // the points that fill no tile.  nullptr: one uniform grid, x_i = x[0] + i (x[n-1] - x[0])/(n-1)
    const double* xt;
    double dxg;
    int64_t ntiles;
    const double* consts;           // k_sinefold: per-chain constants [NCONST, ldc] (k_fold_consts) or nullptr
    int64_t ldc;
    int consts_wait;                // launched as a programmatic dependent of k_fold_consts
    MomentArgs m;
    int64_t n;
    double* partial;
    int64_t ldpartial;
    int use_tma;
    FuseArgs f;
};

// The Metropolis step itself, compiled once per translation unit: S points to the
// CTA's shared-memory copy of the sampler description.
static __device__ __noinline__ void fused_metropolis_step(const mc3b_sampler_t* S, double nxt, int64_t c, int64_t gen,
                                                          int64_t zrow0) {
    metropolis_chain(*S, nxt, gen, zrow0, c);
}

// Hook between the sum of a chain's partial rows and its Metropolis step; called by
// every thread of the reducer CTA (it may use CTA barriers).
struct NoFix { __device__ __forceinline__ void operator()(bool, int64_t, double&) const {} };

// chains_per_cta: chains a CTA covers (its thread t < chains_per_cta owns chain
// blockIdx.x * chains_per_cta + t of the launch).  `f` must be the kernel
// parameter itself (read from the constant bank, never copied to the stack).
//
// The reducer of a chain group is the CTA of its LAST split: CTAs are dispatched in
// block order, so it starts after every other split of the group is running or done,
// and with the decreasing split schedule it is also the shortest.  The other CTAs
// publish their row with one fire-and-forget reduction (no round trip: an atomic
// whose result every CTA waited for cost ~10 us per launch over the six waves) and
// leave; the reducer spins until all nsplit-1 arrivals are in.
template <class Fix = NoFix>
__device__ __forceinline__ void fused_metropolis(const FuseArgs& f, const double* partial, int64_t ldpartial,
                                                 int64_t nchains, int chains_per_cta, Fix fix = Fix()) {
    __shared__ __align__(16) mc3b_sampler_t sS;
    __shared__ double vbuf[STAGE_DOUBLES];
    __syncthreads();                               // the CTA's partial row is written
    const int others = (int)gridDim.y - 1;
    if ((int)blockIdx.y != others) {
        if (threadIdx.x == 0) {
            __threadfence();                       // cumulative over the barrier: the row before the count
            atomicAdd(&f.done[blockIdx.x], 1);
        }
        return;
    }
    if (threadIdx.x == 0) {
        sS = f.S;                                  // constant bank -> shared, static offsets only
        int seen;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(f.done + blockIdx.x) : "memory");
        } while (seen < others);
        f.done[blockIdx.x] = 0;                    // every arrival is in: ready for the next launch
    }
    __syncthreads();
    if (threadIdx.x < 32) {                        // one warp stages the per-parameter vectors
        mc3b_sampler_t T = sS;
        const int np = T.npars, nf = T.nfree;
        const double* src[7] = {T.pstep, T.pmin, T.pmax, T.params0, T.prior, T.priorlow, T.priorup};
        for (int i = threadIdx.x; i < 7 * np; i += 32) {
            const int v = i / np, k = i - v * np;
            if (src[v]) vbuf[v * MAXP + k] = src[v][k];
        }
        int32_t* ifr = reinterpret_cast<int32_t*>(vbuf + 7 * MAXP);
        for (int i = threadIdx.x; i < nf; i += 32) ifr[i] = T.ifree[i];
        stage_peers_load(T, vbuf, threadIdx.x, 32);
        __syncwarp();
        if (threadIdx.x == 0) {
            sS.pstep = vbuf; sS.pmin = vbuf + MAXP; sS.pmax = vbuf + 2 * MAXP; sS.params0 = vbuf + 3 * MAXP;
            if (T.prior) { sS.prior = vbuf + 4 * MAXP; sS.priorlow = vbuf + 5 * MAXP; sS.priorup = vbuf + 6 * MAXP; }
            sS.ifree = ifr;
            stage_peers_point(sS, vbuf);
        }
    }
    __syncthreads();
    const int64_t cl = (int64_t)blockIdx.x * chains_per_cta + threadIdx.x;
    const bool mine = (int)threadIdx.x < chains_per_cta && cl < nchains;
    int64_t gen = f.gen, zrow0 = f.zrow0;
    if (gen < 0) {
        gen = *sS.gen_dev;
        zrow0 = ((gen + 1) % sS.thinning == 0) ? sS.M0 + ((gen + 1) / sS.thinning - 1) * sS.nchains : -1;
    }
    double nxt = 0.0;
    bool inb = false;
    if (mine) {
        inb = sS.inb[f.c_off + cl] != 0;
        if (inb) nxt = sum_partials<true>(partial, ldpartial, (int)gridDim.y, cl);
    }
    fix(inb, cl, nxt);
    if (mine) fused_metropolis_step(&sS, nxt, f.c_off + cl, gen, zrow0);
    __syncthreads();
    if (threadIdx.x == 0 && f.advance) {
        // one fence for the CTA's stores (cumulative over the barrier): system scope when they
        // went to peer devices, so that the flag below needs no fence of its own
        if (sS.X_peers) __threadfence_system(); else __threadfence();
        if (atomicAdd(&f.done[gridDim.x], 1) == (int)gridDim.x - 1) {
            f.done[gridDim.x] = 0;
            __threadfence();
            const int64_t g = *sS.gen_dev + 1;
            *sS.gen_dev = g;
            if (sS.F_peers) flags_publish(sS, g);    // every group's peer stores are performed
        }
    }
}

}  // namespace mc3b_chisq
using namespace mc3b_chisq;

// chisq_grid.cu
int mc3b_launch_sinegrid(const ChisqArgs<double>& a, bool usig, unsigned groups, unsigned nsplit, cudaStream_t st);
int mc3b_launch_sinefold(const ChisqArgs<double>& a, double* work, unsigned groups, unsigned nsplit, cudaStream_t st);
int mc3b_launch_fold(const double* d, int64_t n, double* out, cudaStream_t st);
int mc3b_launch_moment_finish(const ChisqArgs<double>& a, int nsplit, int npars, const double* prior, const double* plo,
                              const double* pup, double* chisq, cudaStream_t st);
int mc3b_launch_moment_prepare(const double* d, int64_t ntiles, double x0, double dx, const double* tile_x,
                               double c0ref, double slref, double* folded, double* tiles, int layout, cudaStream_t st);
