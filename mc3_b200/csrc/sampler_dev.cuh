// mc3_b200 -- per-chain device functions of the sampler (shared by the
// per-generation kernels of sampler.cu and the persistent kernel of small.cu).
#pragma once
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

// Peer-memory exchange: generation flags (include/mc3b200.h, F_peers).
// Precondition: every CTA that stored into peers has executed a system fence (after a CTA
// barrier) before the event that let this thread know the generation is complete.
// ONE system fence here (it orders everything this thread has observed -- the other CTAs'
// fenced stores included -- before the flags), then relaxed stores: a release store per
// peer made every flag wait for the previous flag's round trip over NVLink (~3.3 us per
// peer: +9 / +15 / +29 us per generation at 2 / 4 / 8 GPUs, profiles/r2_bench_n*.json).
__device__ __forceinline__ void flags_publish(const mc3b_sampler_t& S, int64_t done) {
    __threadfence_system();
    for (int p = 0; p < S.world; p++)
        asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(S.F_peers[p] + S.rank), "l"((long long)done)
                     : "memory");
}
// every thread of the CTA calls it (one thread per peer polls, then a CTA barrier)
__device__ __forceinline__ void flags_wait(const mc3b_sampler_t& S, int64_t gen) {
    if ((int)threadIdx.x < S.world) {
        const int64_t* f = S.F_peers[S.rank] + threadIdx.x;
        long long v;
        do {
            asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        } while (v < (long long)gen);
    }
    __syncthreads();
}

// The per-parameter vectors a chain walks in loops (bounds, steps, priors, free
// indices) staged in shared memory by the whole CTA in one round of loads, and the
// sampler description repointed to them: a thread per chain otherwise pays a
// chain of dependent L2 latencies per parameter.  buf: STAGE_DOUBLES doubles.
// The peer pointer tables (X_peers, Z_peers, F_peers: [world] pointers in global memory)
// are staged too: the Metropolis step stores nfree values into every device, and a
// table read from global memory inside those loops cost one L2 round trip per (value,
// peer) -- +3.3 us per peer and generation at config 2 (profiles/r2_bench_n*.json).
constexpr int MAXW = 16;            // devices whose peer pointers are staged (more: read from global)
constexpr int STAGE_PEERS = 7 * MAXP + MAXP / 2;       // offset of the pointer slots in the staging buffer
constexpr int STAGE_DOUBLES = STAGE_PEERS + 3 * MAXW;
__device__ __forceinline__ void stage_peers_load(const mc3b_sampler_t& T, double* buf, int tid, int nthreads) {
    if (T.X_peers == nullptr || T.world > MAXW) return;
    void** slot = reinterpret_cast<void**>(buf + STAGE_PEERS);
    for (int i = tid; i < 3 * T.world; i += nthreads) {
        const int v = i / T.world, p = i - v * T.world;
        if (v == 0) slot[p] = T.X_peers[p];
        else if (v == 1) { if (T.Z_peers) slot[MAXW + p] = T.Z_peers[p]; }
        else if (T.F_peers) slot[2 * MAXW + p] = T.F_peers[p];
    }
}
__device__ __forceinline__ void stage_peers_point(mc3b_sampler_t& S, double* buf) {
    if (S.X_peers == nullptr || S.world > MAXW) return;
    void** slot = reinterpret_cast<void**>(buf + STAGE_PEERS);
    S.X_peers = reinterpret_cast<double* const*>(slot);
    if (S.Z_peers) S.Z_peers = reinterpret_cast<double* const*>(slot + MAXW);
    if (S.F_peers) S.F_peers = reinterpret_cast<int64_t* const*>(slot + 2 * MAXW);
}
__device__ __forceinline__ void stage_vectors(mc3b_sampler_t& S, double* buf) {
    const int np = S.npars, nf = S.nfree;
    const double* src[7] = {S.pstep, S.pmin, S.pmax, S.params0, S.prior, S.priorlow, S.priorup};
    for (int i = threadIdx.x; i < 7 * np; i += blockDim.x) {
        const int v = i / np, k = i - v * np;
        if (src[v]) buf[v * MAXP + k] = src[v][k];
    }
    int32_t* ifr = reinterpret_cast<int32_t*>(buf + 7 * MAXP);
    for (int i = threadIdx.x; i < nf; i += blockDim.x) ifr[i] = S.ifree[i];
    stage_peers_load(S, buf, threadIdx.x, blockDim.x);
    __syncthreads();
    S.pstep = buf; S.pmin = buf + MAXP; S.pmax = buf + 2 * MAXP; S.params0 = buf + 3 * MAXP;
    if (S.prior) { S.prior = buf + 4 * MAXP; S.priorlow = buf + 5 * MAXP; S.priorup = buf + 6 * MAXP; }
    S.ifree = ifr;
    stage_peers_point(S, buf);
}

// The random numbers chain c consumes in generation gen (chain.py:185, 197-203,
// 223-229, 257): a function of (seed, chain, generation) only -- not of the chains'
// states, so a proposal kernel launched as a programmatic dependent of the previous
// generation's model kernel draws them while that kernel is still running.
template <bool REPLAY>
__device__ __forceinline__ void chain_draws(const mc3b_sampler_t& S, const mc3b_draws_t& D, int64_t gen,
                                            int64_t zsize, int64_t c, Draws& dr, double* nrm) {
    const int nfree = S.nfree;
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
}

// Jump, bounds, shared parameters and the snooker factor from the draws and the
// population as of the start of generation gen (chain.py:195-255).
__device__ __forceinline__ void propose_apply(const mc3b_sampler_t& S, const Draws& dr, const double* nrm,
                                              int64_t gen, int64_t c) {
    const int nfree = S.nfree, npars = S.npars;
    // population as of the start of this generation (peer mode: half gen & 1)
    const double* Xg = S.X_peers ? S.X_peers[S.rank] + (gen & 1) * S.nchains * nfree : S.X;
    const double* x = Xg + c * nfree;
    double jump[MAXP];

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
        const double* x1 = Xg + dr.a * nfree;
        const double* x2 = Xg + dr.b * nfree;
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

// Proposal of chain c for generation gen (chain.py:185-247, 251-255).
template <bool REPLAY>
__device__ __forceinline__ void propose_chain(const mc3b_sampler_t& S, const mc3b_draws_t& D, int64_t gen,
                                              int64_t zsize, int64_t c) {
    Draws dr;
    double nrm[MAXP];
    chain_draws<REPLAY>(S, D, gen, zsize, c, dr, nrm);
    propose_apply(S, dr, nrm, gen, c);
}

// Data chi-squared of a proposal: the partial rows of the model kernel added in
// split order (fixed order = identical bits wherever the sum is taken).  CG: read
// through L2 (rows written by other SMs during the same kernel).
template <bool CG>
__device__ __forceinline__ double sum_partials(const double* partial, int64_t ldpartial, int nsplit, int64_t col) {
    double nxt = 0.0;
    const double* p = partial + col;
    int s = 0;
    for (; s + 32 <= nsplit; s += 32) {             // 32 loads in flight (the rows come from L2), added in order
        double v[32];
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] = CG ? __ldcg(p + (int64_t)(s + k) * ldpartial) : p[(int64_t)(s + k) * ldpartial];
#pragma unroll
        for (int k = 0; k < 32; k++) nxt += v[k];
    }
    for (; s + 8 <= nsplit; s += 8) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = CG ? __ldcg(p + (int64_t)(s + k) * ldpartial) : p[(int64_t)(s + k) * ldpartial];
#pragma unroll
        for (int k = 0; k < 8; k++) nxt += v[k];
    }
    for (; s < nsplit; s++) nxt += CG ? __ldcg(p + (int64_t)s * ldpartial) : p[(int64_t)s * ldpartial];
    return nxt;
}

// Metropolis step of chain c (chain.py:257-289); nxt_data = data chi-squared of its
// proposal (unused when the proposal was out of bounds).
__device__ __forceinline__ void metropolis_chain(const mc3b_sampler_t& S, double nxt_data, int64_t gen, int64_t zrow0,
                                                 int64_t c) {
    const int nfree = S.nfree, npars = S.npars;
    const bool peer = S.X_peers != nullptr;
    const int64_t half = S.nchains * nfree;
    double* x = peer ? S.X_peers[S.rank] + (gen & 1) * half + c * nfree : S.X + c * nfree;
    double cur = S.chisq_cur[c];
    bool accept = false;
    const double* np_ = S.nextp + c * npars;
    if (S.inb[c]) {
        double nxt = nxt_data;
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
            accept = true;
            cur = nxt;
            S.chisq_cur[c] = nxt;
            S.naccept[c] += 1;
            if (nxt < S.best_chisq[c]) {
                S.best_chisq[c] = nxt;
                S.best_gen[c] = gen;
                for (int j = 0; j < nfree; j++) S.best_x[c * nfree + j] = np_[S.ifree[j]];
            }
        }
    }
    const bool write = zrow0 >= 0 && zrow0 + c < S.zlen;
    const int64_t row = zrow0 + c;
    if (!peer) {
        if (accept)
            for (int j = 0; j < nfree; j++) x[j] = np_[S.ifree[j]];
        if (write)
            for (int j = 0; j < nfree; j++) S.Z[row * nfree + j] = x[j];      // chain.py:276-289
    } else {
        // next state of this chain into half (gen+1)&1 of EVERY device (NVLink
        // peer stores); thinned rows likewise when the history is shared
        const int64_t dst = ((gen + 1) & 1) * half + c * nfree;
        for (int j = 0; j < nfree; j++) {
            const double v = accept ? np_[S.ifree[j]] : x[j];
            for (int p = 0; p < S.world; p++) S.X_peers[p][dst + j] = v;
            if (write) {
                if (S.Z_peers) { for (int p = 0; p < S.world; p++) S.Z_peers[p][row * nfree + j] = v; }
                else S.Z[row * nfree + j] = v;
            }
        }
        // (no fence here: the caller orders the CTA's peer stores with ONE system fence after
        // a CTA barrier, before it counts the group as done -- see fused_metropolis / k_metropolis)
    }
    if (write) {
        S.log_post[row] = -0.5 * cur;
        S.zchain[row] = (int32_t)c;
    }
}


}  // namespace
