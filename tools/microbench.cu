// Microbenchmarks that give the non-tensor roofline denominators the kNN / gather kernels are judged against
// ("FP32 non-tensor FFMA peak, smem/L2 bandwidth: to be measured by the builder").
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench.cu -o /tmp/microbench && /tmp/microbench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE>
__global__ void fp32_kernel(float* out, int iters, float a, float b) {
    // 8 independent chains per thread
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {            // scalar FFMA x2 (two lanes of the float2)
                x[i].x = __fmaf_rn(x[i].x, a, b);
                x[i].y = __fmaf_rn(x[i].y, a, b);
            } else if (MODE == 1) {     // packed FFMA2
                x[i] = __ffma2_rn(x[i], a2, b2);
            } else if (MODE == 2) {     // scalar FMUL + FADD (separately rounded)
                x[i].x = __fadd_rn(__fmul_rn(x[i].x, a), b);
                x[i].y = __fadd_rn(__fmul_rn(x[i].y, a), b);
            } else if (MODE == 3) {     // packed FMUL2 then scalar FADDs (the MULADD kNN pattern)
                float2 m = __fmul2_rn(x[i], a2);
                x[i].x = __fadd_rn(m.x, b);
                x[i].y = __fadd_rn(m.y, b);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
double run_fp32(int sms, const char* name, double flop_per_iter_thread) {
    float* out;
    CK(cudaMalloc(&out, 4));
    const int iters = 4096, threads = 256, blocks = sms * 8;
    fp32_kernel<MODE><<<blocks, threads>>>(out, 64, 1.0001f, 1e-7f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    fp32_kernel<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 1e-7f);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = (double)blocks * threads * iters * 16.0;   // 16 element-ops (mul-add pairs) per iteration
    printf("fp32 %-28s %8.3f ms  %8.2f T element-madd/s  (%.1f TFLOP/s counting 2 flop)\n", name, ms, lane_ops / ms * 1e-9,
           lane_ops * flop_per_iter_thread / ms * 1e-9);
    cudaFree(out);
    return lane_ops / ms * 1e-9;
}

// random row gather: each warp sums `per_warp` random rows of `row_bytes` from a table of `rows` rows
__global__ void gather_kernel(const float2* __restrict__ tab, int rows, int row_f2, int per_warp, float* out, int window) {
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned h = warp * 2654435761u + 12345u;
    float2 acc = make_float2(0.f, 0.f);
    const int base = (int)((warp * 97u) % (unsigned)rows);
    for (int i = 0; i < per_warp; i += 4) {
        float2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            h = h * 1664525u + 1013904223u;
            int r = window ? ((base + (int)((h >> 8) & (unsigned)(window - 1))) & (rows - 1)) : (int)((h >> 8) & (unsigned)(rows - 1));   // rows, window: powers of two
            v[u] = (lane < row_f2) ? __ldg(tab + (size_t)r * row_f2 + lane) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
    }
    if (acc.x == 1234.5f) out[0] = acc.y;
}

void run_gather(int sms, int rows, int row_bytes, int window, const char* name) {
    float2* tab; float* out;
    size_t bytes = (size_t)rows * row_bytes;
    CK(cudaMalloc(&tab, bytes)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(tab, 0, bytes));
    const int row_f2 = row_bytes / 8, per_warp = 2048, threads = 256, blocks = sms * 16;
    gather_kernel<<<blocks, threads>>>(tab, rows, row_f2, 64, out, window);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    gather_kernel<<<blocks, threads>>>(tab, rows, row_f2, per_warp, out, window);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double got = (double)blocks * (threads / 32) * per_warp * row_bytes;
    printf("gather %-40s table %7.1f MiB  %8.3f ms  %8.1f GB/s\n", name, bytes / 1048576.0, ms, got / ms * 1e-6);
    cudaFree(tab); cudaFree(out);
}

__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) b[i] = a[i];
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s  SMs %d  clock %d MHz  L2 %d MiB  smem/SM %zu KB\n", p.name, sms, p.clockRate / 1000, p.l2CacheSize >> 20,
           p.sharedMemPerMultiprocessor >> 10);
    run_fp32<0>(sms, "FFMA scalar", 2.0);
    run_fp32<1>(sms, "FFMA2 packed", 2.0);
    run_fp32<2>(sms, "FMUL+FADD scalar", 2.0);
    run_fp32<3>(sms, "FMUL2 + scalar FADD", 2.0);
    // HBM copy for reference
    {
        size_t n = (size_t)1 << 28;   // 256 Mi floats4 = 4 GiB? keep 1 GiB: 64 Mi float4
        n = (size_t)64 << 20;
        float4 *a, *b;
        CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
        CK(cudaMemset(a, 1, n * 16));
        copy_kernel<<<sms * 16, 512>>>(a, b, n);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        copy_kernel<<<sms * 16, 512>>>(a, b, n);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("hbm copy 1 GiB read + 1 GiB write: %.3f ms  %.1f GB/s\n", ms, 2.0 * n * 16 / ms * 1e-6);
        cudaFree(a); cudaFree(b);
    }
    run_gather(sms, 4096 * 16, 256, 0, "256B rows, random over 16 clouds");
    run_gather(sms, 4096 * 64, 256, 0, "256B rows, random over 64 clouds");
    run_gather(sms, 4096 * 256, 256, 0, "256B rows, random over 256 clouds");
    run_gather(sms, 4096 * 64, 128, 0, "128B rows (fp16), random over 64 clouds");
    run_gather(sms, 4096 * 64, 256, 512, "256B rows, 512-row window (Morton-local)");
    run_gather(sms, 4096 * 64, 256, 128, "256B rows, 128-row window (Morton-local)");
    return 0;
}
