// mc3_b200 -- built-in model functions, evaluated per (chain, data point).
//
// The reference has no built-in models: its `func` is user Python evaluated
// once per chain-step (mc3/chain.py:316-319, ~90% of its run time).  These
// functors are the CUDA side of BASELINE.json's north_star item (2); their
// numpy twins for the oracle live in oracle/models.py (same formulas, same
// parameter order).  load() runs once per chain per launch and hoists every
// per-chain constant; eval() is the per-point cost.
#pragma once
#include "common.cuh"

// Branch-free fp64 sine: 14 FP64-pipe instructions against ~35 (plus a divergent
// sin/cos kernel choice) of the CUDA library's sin().  Reduction modulo pi with a
// two-constant Cody-Waite step (each FMA rounds once, so the reduced argument
// carries <= 4e-16 absolute error for any |a| < 1e9), odd polynomial of degree 17
// on [-pi/2, pi/2] (near-minimax fit, max error 3.6e-17 + rounding), sign from
// the parity of the quotient.  Measured against numpy.sin in
// tests/test_gpu_kernels.py::test_fast_sin_accuracy.  Huge or non-finite
// arguments take the library path.
__device__ __constant__ double kSinC[12] = {
    0.3183098861837907,          // 1/pi
    -3.141592653589793,          // -pi (high part)
    -1.2246467991473532e-16,     // -pi (low part)
    2.7314447669863995e-15, -7.643970296798572e-13, 1.6058977312464087e-10, -2.5052107616996182e-08,
    2.7557319219163234e-06, -0.00019841269841254974, 0.008333333333333316, -0.16666666666666666, 0.0};

__device__ __forceinline__ double fast_sin(double a) {
    const double MAGIC = 6755399441055744.0;               // 1.5 * 2^52
    const double t = fma(a, kSinC[0], MAGIC);               // low word = rint(a/pi)
    const double q = t - MAGIC;
    double r = fma(q, kSinC[1], a);
    r = fma(q, kSinC[2], r);
    const double s = r * r;
    double p = kSinC[3];
    p = fma(p, s, kSinC[4]);
    p = fma(p, s, kSinC[5]);
    p = fma(p, s, kSinC[6]);
    p = fma(p, s, kSinC[7]);
    p = fma(p, s, kSinC[8]);
    p = fma(p, s, kSinC[9]);
    p = fma(p, s, kSinC[10]);
    double v = fma(r * s, p, r);
    v = __hiloint2double(__double2hiint(v) ^ (__double2loint(t) << 31), __double2loint(v));
    if ((__double2hiint(a) & 0x7fffffff) >= 0x41cdcd65) v = sin(a);   // |a| >= 1e9, inf, nan
    return v;
}

template <typename T> struct mathx;
template <> struct mathx<double> {
    static __device__ __forceinline__ double sin_(double v) { return fast_sin(v); }
    static __device__ __forceinline__ double exp_(double v) { return exp(v); }
};
template <> struct mathx<float> {
    static __device__ __forceinline__ float sin_(float v) { return sinf(v); }
    static __device__ __forceinline__ float exp_(float v) { return expf(v); }
};

// y = sum_k p[k] x^k, NP coefficients (Horner).  get_started's quad() is NP=3.
template <typename T, int NP> struct PolyModel {
    T c[NP];
    __device__ __forceinline__ void load(const double* p) {
#pragma unroll
        for (int k = 0; k < NP; k++) c[k] = (T)p[k];
    }
    __device__ __forceinline__ T eval(T x) const {
        T y = c[NP - 1];
#pragma unroll
        for (int k = NP - 2; k >= 0; k--) y = fma(y, x, c[k]);
        return y;
    }
};

// y = p0 sin(2 pi x / p1 + p2) + p3 + p4 x        (BASELINE config 2)
template <typename T> struct SineModel {
    T a, k, ph, c, s;
    __device__ __forceinline__ void load(const double* p) {
        a = (T)p[0]; k = (T)(6.283185307179586476925287 / p[1]); ph = (T)p[2];
        c = (T)p[3]; s = (T)p[4];
    }
    __device__ __forceinline__ T eval(T x) const {
        return fma(a, mathx<T>::sin_(fma(x, k, ph)), fma(s, x, c));
    }
};

// y = p0 exp(-0.5 ((x - p1)/p2)^2) + p3           (Gaussian line)
template <typename T> struct GaussModel {
    T a, mu, is, c;
    __device__ __forceinline__ void load(const double* p) {
        a = (T)p[0]; mu = (T)p[1]; is = (T)(1.0 / p[2]); c = (T)p[3];
    }
    __device__ __forceinline__ T eval(T x) const {
        T d = (x - mu) * is;
        return fma(a, mathx<T>::exp_((T)(-0.5) * d * d), c);
    }
};

// y = p3 - p0 [ |x - p1| < p2/2 ]                 (transit-like box, config 3)
template <typename T> struct BoxModel {
    T lo, base, t0, h;
    __device__ __forceinline__ void load(const double* p) {
        base = (T)p[3]; lo = (T)(p[3] - p[0]); t0 = (T)p[1]; h = (T)(0.5 * p[2]);
    }
    __device__ __forceinline__ T eval(T x) const { return (fabs(x - t0) < h) ? lo : base; }
};

// Dispatch a generic lambda-like functor F<Model> over (model_id, nmodel).
#define MC3B_DISPATCH_MODEL(T, model_id, nmodel, CALL)                                   \
    switch (model_id) {                                                                  \
    case MC3B_MODEL_POLYNOMIAL:                                                          \
        switch (nmodel) {                                                                \
        case 1: { using M = PolyModel<T, 1>; CALL; } break;                              \
        case 2: { using M = PolyModel<T, 2>; CALL; } break;                              \
        case 3: { using M = PolyModel<T, 3>; CALL; } break;                              \
        case 4: { using M = PolyModel<T, 4>; CALL; } break;                              \
        case 5: { using M = PolyModel<T, 5>; CALL; } break;                              \
        case 6: { using M = PolyModel<T, 6>; CALL; } break;                              \
        case 7: { using M = PolyModel<T, 7>; CALL; } break;                              \
        case 8: { using M = PolyModel<T, 8>; CALL; } break;                              \
        default: mc3b_set_error("polynomial needs 1..8 coefficients, got %d", nmodel);   \
                 return MC3B_ERR_ARG;                                                    \
        } break;                                                                         \
    case MC3B_MODEL_SINUSOID: { using M = SineModel<T>; CALL; } break;                   \
    case MC3B_MODEL_GAUSSIAN: { using M = GaussModel<T>; CALL; } break;                  \
    case MC3B_MODEL_BOX: { using M = BoxModel<T>; CALL; } break;                         \
    default: mc3b_set_error("unknown model id %d", model_id); return MC3B_ERR_ARG;       \
    }

static inline int mc3b_model_nparams(int model_id, int nmodel) {
    switch (model_id) {
    case MC3B_MODEL_POLYNOMIAL: return nmodel;
    case MC3B_MODEL_SINUSOID: return 5;
    case MC3B_MODEL_GAUSSIAN: return 4;
    case MC3B_MODEL_BOX: return 4;
    }
    return -1;
}
