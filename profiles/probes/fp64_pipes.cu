// Probe: does DMMA.8x8x4 (FP64 tensor-core MMA) share the FP64 FMA pipe on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
// Prints, per mix of (DFMA, DMMA) instructions per loop iteration, the time and the
// rates; if the mixed loop takes max(t_dfma, t_dmma) the pipes are separate, if it
// takes the sum they are one.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NF, int NM>
__global__ void __launch_bounds__(256) k_mix(int64_t iters, double* sink) {
    double f[NF > 0 ? NF : 1], c[NM > 0 ? NM : 1][2];
#pragma unroll
    for (int k = 0; k < NF; k++) f[k] = (threadIdx.x + k) * 1e-3;
#pragma unroll
    for (int k = 0; k < NM; k++) { c[k][0] = threadIdx.x * 1e-3 + k; c[k][1] = k; }
    const double m = 0.999999 + 1e-9 * threadIdx.x, a = 0.5 + 1e-9 * threadIdx.x, b = 1e-3 * (threadIdx.x & 7);
    for (int64_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < NM; k++) dmma884(c[k][0], c[k][1], a, b);
#pragma unroll
            for (int k = 0; k < NF; k++) f[k] = fma(f[k], m, 1e-7);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NF; k++) s += f[k];
#pragma unroll
    for (int k = 0; k < NM; k++) s += c[k][0] + c[k][1];
    if (s == 123456.789) sink[0] = s;
}

template <int NF, int NM>
void run(const char* name, int sms, double* sink) {
    const int64_t iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_mix<NF, NM><<<sms * 4, 256>>>(iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double warps = (double)sms * 4 * 8;
    const double nf = warps * iters * 4 * NF, nm = warps * iters * 4 * NM;      // warp instructions
    printf("{\"mix\": \"%s\", \"dfma_per_iter\": %d, \"dmma_per_iter\": %d, \"ms\": %.4f, "
           "\"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"cycles_per_dfma_per_smsp\": %.3f, "
           "\"cycles_per_dmma_per_smsp\": %.3f}\n",
           name, NF, NM, best, nf * 64 / (best * 1e-3) / 1e12, nm * 512 / (best * 1e-3) / 1e12,
           NF ? best * 1e-3 * 1.965e9 / (nf / (sms * 4)) : 0.0, NM ? best * 1e-3 * 1.965e9 / (nm / (sms * 4)) : 0.0);
}

// ---- register-file operand cost of FP64 instructions --------------------------
// MODE 1: DADD, two fresh registers          a[k] += b[k]
// MODE 2: DFMA r*r+q (same register twice)   q[k] = fma(a[k], a[k], q[k])
// MODE 3: DFMA, three fresh registers        q[k] = fma(a[k], b[k], q[k])
// MODE 4: DFMA, one multiplier shared        q[k] = fma(n0, a[k], q[k])
// MODE 5: DADD with one shared addend        a[k] += n0
// MODE 6: MODE 1 plus one LDS.128 per four instructions
// MODE 7: the chain-point body of k_sinegrid_usig on registers (6 instructions per point)
template <int MODE>
__global__ void __launch_bounds__(128, 6) k_rf(int64_t iters, double* sink) {
    __shared__ double2 sh[128];
    double a[8], b[8], q[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        a[k] = (threadIdx.x + k) * 1e-3; b[k] = 1e-9 * (threadIdx.x + 3 * k + 1); q[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("" : "+d"(a[k]), "+d"(b[k]));      // opaque: no strength reduction
    sh[threadIdx.x] = make_double2(a[0], b[0]);
    double n0 = -1e-6 * (1 + threadIdx.x), dl = 1e-7 * threadIdx.x;
    asm volatile("" : "+d"(n0), "+d"(dl));
    __syncthreads();
    for (int64_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < 8; k++) a[k] += b[k];
            } else if (MODE == 2) {
#pragma unroll
                for (int k = 0; k < 8; k++) q[k] = fma(a[k], a[k], q[k]);
            } else if (MODE == 3) {
#pragma unroll
                for (int k = 0; k < 8; k++) q[k] = fma(a[k], b[k], q[k]);
            } else if (MODE == 4) {
#pragma unroll
                for (int k = 0; k < 8; k++) q[k] = fma(n0, a[k], q[k]);
            } else if (MODE == 5) {
#pragma unroll
                for (int k = 0; k < 8; k++) a[k] += n0;
            } else if (MODE == 6) {
#pragma unroll
                for (int k = 0; k < 8; k += 4) {
                    const double2 v = sh[(threadIdx.x + k + r + (int)i) & 127];
                    a[k] += b[k]; a[k + 1] += b[k + 1]; a[k + 2] += v.x; a[k + 3] += v.y;
                }
            } else {
                // a[0..3] = s, a[4..7] = du, b[0..3] = L + d stand-ins, q = accumulators
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const double rr = (b[u] + a[u]) - b[4 + u];
                    q[u] = fma(rr, rr, q[u]);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) a[4 + u] = fma(n0, a[u], a[4 + u]);
#pragma unroll
                for (int u = 0; u < 4; u++) a[u] += a[4 + u];
#pragma unroll
                for (int u = 0; u < 4; u++) b[u] += dl;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k] + q[k] + b[k];
    if (s == 123456.789) sink[0] = s;
}

template <int MODE>
void run_rf(const char* name, int per_iter, int sms, double* sink) {
    const int64_t iters = 6000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_rf<MODE><<<sms * 6, 128>>>(iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double instr_per_smsp = 6.0 * iters * 4 * per_iter;      // 6 CTAs x 4 warps / 4 SMSPs
    printf("{\"rf_probe\": \"%s\", \"fp64_instr_per_iter\": %d, \"ms\": %.4f, \"cycles_per_fp64_instr_per_smsp\": %.3f}\n",
           name, per_iter, best, best * 1e-3 * 1.965e9 / instr_per_smsp);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* sink; cudaMalloc(&sink, 64);
    run<8, 0>("dfma only", sms, sink);
    run<0, 4>("dmma only (4 acc)", sms, sink);
    run<0, 8>("dmma only (8 acc)", sms, sink);
    run<8, 1>("8 dfma + 1 dmma", sms, sink);
    run<8, 2>("8 dfma + 2 dmma", sms, sink);
    run<8, 4>("8 dfma + 4 dmma", sms, sink);
    run<4, 4>("4 dfma + 4 dmma", sms, sink);
    run_rf<1>("DADD a+=b (2 fresh)", 8, sms, sink);
    run_rf<2>("DFMA q=a*a+q (same reg twice)", 8, sms, sink);
    run_rf<3>("DFMA q=a*b+q (3 fresh)", 8, sms, sink);
    run_rf<4>("DFMA q=n0*a+q (shared multiplier)", 8, sms, sink);
    run_rf<5>("DADD a+=n0 (shared addend)", 8, sms, sink);
    run_rf<6>("DADD 2 fresh + LDS.128 per 4", 8, sms, sink);
    run_rf<7>("k_sinegrid_usig chain-point body (24 instr per 4 points)", 24, sms, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
