// bsw_intpeak.cu -- INT-pipe micro-benchmark: the roofline denominator for the extension kernels.
//
// MEASURED_PEAKS.json only holds HBM bandwidth and bf16 tensor throughput; the seed-extension DP is bound by
// the integer ALU pipe (13 integer ops per cell), so the denominator is measured here:
// dependency-free streams of 16 independent accumulators per thread, every SM fully occupied.
//   mode 0: IADD3              (1 op / instruction)
//   mode 1: IMNMX / VIMNMX     (1 op / instruction)
//   mode 2: VIADDMNMX (DPX)    (fused add+max, counted as 2 ops / instruction)
//   mode 3: the scalar cell body of the recurrence without memory (13 ops / cell, whatever SASS it becomes)
//   mode 4: IADD3 and IMAD interleaved (adds issued to both the ALU and the FMA pipe)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "bsw_kernels.h"

namespace bsw {

constexpr int PEAK_ACC = 16;
constexpr int PEAK_THREADS = 256;

template <int MODE>
__global__ void __launch_bounds__(PEAK_THREADS) int_peak_kernel(int iters, int a, int b, int* out, long long* clk)
{
    int r[PEAK_ACC];
#pragma unroll
    for (int k = 0; k < PEAK_ACC; ++k) r[k] = (int)threadIdx.x * (k + 1) + a;
    long long c0 = 0, t0 = 0;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        c0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < PEAK_ACC; ++k) {
            const int o = r[(k + 1) & (PEAK_ACC - 1)];
            if (MODE == 0) r[k] = r[k] + o + a;
            else if (MODE == 1) r[k] = max(r[k], o);
            else if (MODE == 2) r[k] = __viaddmax_s32(r[k], a, o);
            else if (MODE == 4) { if (k & 1) r[k] = r[k] * b + o; else r[k] = r[k] + o + a; }
            else if (MODE == 5) r[k] = (int)__umulhi((unsigned)r[k], (unsigned)(b + 0x10000)) + o;
            else if (MODE == 6) r[k] = (int)__byte_perm((unsigned)r[k], (unsigned)o, 0x7632);
        }
        if (MODE == 3) {
            // 4 independent cells per iteration: M,e,f,h1 state per chain; 13 ops each
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int M = r[4 * c], e = r[4 * c + 1], f = r[4 * c + 2], m = r[4 * c + 3];
                int h = M + a;            // add
                h = max(h, e);            // max
                h = max(h, f);            // max
                int mj = m > h ? b : it;  // select
                m = max(m, h);            // max
                int t = h - b;            // sub
                t = max(t, 0);            // max
                e = e - a;                // sub
                e = max(e, t);            // max
                int t2 = h - a;           // sub
                t2 = max(t2, 0);          // max
                f = f - b;                // sub
                f = max(f, t2);           // max
                r[4 * c] = h ^ mj; r[4 * c + 1] = e; r[4 * c + 2] = f; r[4 * c + 3] = m;
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < PEAK_ACC; ++k) s ^= r[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        long long c1 = clock64(), t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        clk[0] = c1 - c0; clk[1] = t1 - t0;
    }
}

template <int MODE>
static cudaError_t run_mode(int grid, int iters, int* d_out, long long* d_clk, cudaStream_t st, float* ms)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int_peak_kernel<MODE><<<grid, PEAK_THREADS, 0, st>>>(iters / 8, 1, 1, d_out, d_clk);     // warm-up
    cudaEventRecord(e0, st);
    int_peak_kernel<MODE><<<grid, PEAK_THREADS, 0, st>>>(iters, 1, 1, d_out, d_clk);
    cudaEventRecord(e1, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return e;
}

cudaError_t int_peak_run(double out_ops[5], double* sm_clock_mhz, int* sm_count, cudaStream_t st)
{
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const int grid = sms * 8;
    const int iters = 8192;
    int* d_out = nullptr; long long* d_clk = nullptr;
    if ((e = cudaMalloc((void**)&d_out, (size_t)grid * PEAK_THREADS * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&d_clk, 2 * sizeof(long long))) != cudaSuccess) { cudaFree(d_out); return e; }
    float ms[5] = { 0, 0, 0, 0, 0 };
    e = run_mode<0>(grid, iters, d_out, d_clk, st, &ms[0]);
    if (e == cudaSuccess) e = run_mode<1>(grid, iters, d_out, d_clk, st, &ms[1]);
    if (e == cudaSuccess) e = run_mode<2>(grid, iters, d_out, d_clk, st, &ms[2]);
    if (e == cudaSuccess) e = run_mode<3>(grid, iters, d_out, d_clk, st, &ms[3]);
    if (e == cudaSuccess) e = run_mode<4>(grid, iters, d_out, d_clk, st, &ms[4]);
    float ms5 = 0, ms6 = 0;
    if (e == cudaSuccess) e = run_mode<5>(grid, iters, d_out, d_clk, st, &ms5);
    if (e == cudaSuccess) e = run_mode<6>(grid, iters, d_out, d_clk, st, &ms6);
    if (getenv("BSW_TRACE")) fprintf(stderr, "[intpeak] IMAD.HI stream %.2f Tops/s, PRMT stream %.2f Tops/s\n", (double)grid * PEAK_THREADS * iters * PEAK_ACC / (ms5 * 1e-3) * 1e-12, (double)grid * PEAK_THREADS * iters * PEAK_ACC / (ms6 * 1e-3) * 1e-12);
    long long clk[2] = { 0, 1 };
    if (e == cudaSuccess) e = cudaMemcpy(clk, d_clk, sizeof(clk), cudaMemcpyDeviceToHost);
    cudaFree(d_out); cudaFree(d_clk);
    if (e != cudaSuccess) return e;
    const double threads = (double)grid * PEAK_THREADS;
    out_ops[0] = threads * iters * PEAK_ACC / (ms[0] * 1e-3);
    out_ops[1] = threads * iters * PEAK_ACC / (ms[1] * 1e-3);
    out_ops[2] = threads * iters * PEAK_ACC * 2.0 / (ms[2] * 1e-3);
    out_ops[3] = threads * iters * 4.0 * 13.0 / (ms[3] * 1e-3);
    out_ops[4] = threads * iters * PEAK_ACC / (ms[4] * 1e-3);
    *sm_clock_mhz = clk[1] > 0 ? (double)clk[0] / (double)clk[1] * 1e3 : 0.0;
    *sm_count = sms;
    return cudaSuccess;
}

}  // namespace bsw
