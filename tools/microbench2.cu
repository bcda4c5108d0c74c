// Microbenchmarks behind the round-2 kNN design (thread-per-row scan with candidates broadcast from shared memory and a
// register insertion list): broadcast LDS.128 rate, FMNMX / FMNMX3 rate, and their mix with FFMA.  SM-cycle timed.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench2.cu -o /tmp/microbench2 && /tmp/microbench2
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// MODE 0: every lane reads the SAME float4 (broadcast); 1: lane-consecutive float4 (conflict-free, 4 wavefronts);
// 2: broadcast LDS.64; 3: broadcast LDS.32
template <int MODE>
__global__ void lds_kernel(float* out, long long* cyc, int iters) {
    __shared__ float4 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int j = (it * 16 + u) & 1023;
            if (MODE == 0) {
                const float4 p = buf[j];
                acc += p.x + p.y * p.z + p.w;
            } else if (MODE == 1) {
                const float4 p = buf[j + lane];
                acc += p.x + p.y * p.z + p.w;
            } else if (MODE == 2) {
                const float2 p = reinterpret_cast<const float2*>(buf)[j];
                acc += p.x * p.y;
            } else {
                acc += reinterpret_cast<const float*>(buf)[j];
            }
        }
    }
    const long long t1 = clock64();
    if (acc == 12345.678f) out[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// MODE 0: FMNMX chain x8 independent; 1: FMNMX3; 2: insertion step L'[i] = min(L[i], max(L[i-1], c)) over 20 slots;
// 3: merge-2 step M[i] = min3(L[i], max(L[i-1],c1), max(L[i-2],c2)); 4: FFMA x8 (reference); 5: FFMA + FMNMX interleaved
template <int MODE>
__global__ void alu_kernel(float* out, long long* cyc, int iters, float a, float b) {
    float x[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) x[i] = threadIdx.x * 1e-3f + i;
    float c = a;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fminf(x[i], c + i);
            c += b;
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fminf(fminf(x[i], x[i + 8]), c);
            c += b;
        } else if (MODE == 2) {
            float prev = -1e30f;
#pragma unroll
            for (int i = 0; i < 20; ++i) {
                const float li = x[i];
                x[i] = fminf(li, fmaxf(prev, c));
                prev = li;
            }
            c = c * a + b;
        } else if (MODE == 3) {
            const float c1 = fminf(c, c * a), c2 = fmaxf(c, c * a);
            float p1 = -1e30f, p2 = -1e30f;
#pragma unroll
            for (int i = 0; i < 20; ++i) {
                const float li = x[i];
                x[i] = fminf(fminf(li, fmaxf(p1, c1)), fmaxf(p2, c2));
                p2 = p1;
                p1 = li;
            }
            c = c * a + b;
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __fmaf_rn(x[i], a, b);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = __fmaf_rn(x[i], a, b);
                x[i + 8] = fminf(x[i + 8], x[i]);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 20; ++i) s += x[i];
    if (s == 12345.678f) out[0] = s + c;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
double run(F launch, int blocks, long long* dcyc) {
    launch();
    CK(cudaDeviceSynchronize());
    launch();
    CK(cudaDeviceSynchronize());
    long long* h = (long long*)malloc(sizeof(long long) * blocks);
    CK(cudaMemcpy(h, dcyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? h[i] : mx;
    free(h);
    return (double)mx;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    float* out;
    long long* cyc;
    CK(cudaMalloc(&out, 4));
    CK(cudaMalloc(&cyc, sizeof(long long) * sms * 4));
    printf("device %s  SMs %d\n", p.name, sms);
    const int iters = 2048;
    for (int warps : {8, 16, 32}) {
        const int threads = warps * 32;
        const double lds_n = (double)warps * iters * 16;      // warp-level LDS instructions per SM
        double c;
        c = run([&] { lds_kernel<0><<<sms, threads>>>(out, cyc, iters); }, sms, cyc);
        printf("warps/SM %2d  LDS.128 broadcast      %6.2f cycles per warp-LDS  (%.2f candidates/cycle/SM)\n", warps, c / lds_n, lds_n / c);
        c = run([&] { lds_kernel<1><<<sms, threads>>>(out, cyc, iters); }, sms, cyc);
        printf("warps/SM %2d  LDS.128 lane-consecutive %6.2f cycles per warp-LDS\n", warps, c / lds_n);
        c = run([&] { lds_kernel<2><<<sms, threads>>>(out, cyc, iters); }, sms, cyc);
        printf("warps/SM %2d  LDS.64  broadcast      %6.2f cycles per warp-LDS\n", warps, c / lds_n);
        c = run([&] { lds_kernel<3><<<sms, threads>>>(out, cyc, iters); }, sms, cyc);
        printf("warps/SM %2d  LDS.32  broadcast      %6.2f cycles per warp-LDS\n", warps, c / lds_n);
    }
    for (int warps : {16, 32}) {
        const int threads = warps * 32;
        double c;
        c = run([&] { alu_kernel<0><<<sms, threads>>>(out, cyc, iters, 1.0001f, 1e-7f); }, sms, cyc);
        printf("warps/SM %2d  FMNMX  x8 indep        %6.3f warp-instr/cycle/SM\n", warps, (double)warps * iters * 8 / c);
        c = run([&] { alu_kernel<1><<<sms, threads>>>(out, cyc, iters, 1.0001f, 1e-7f); }, sms, cyc);
        printf("warps/SM %2d  FMNMX3 x8 indep        %6.3f warp-instr/cycle/SM\n", warps, (double)warps * iters * 8 / c);
        c = run([&] { alu_kernel<2><<<sms, threads>>>(out, cyc, iters, 1.0001f, 1e-7f); }, sms, cyc);
        printf("warps/SM %2d  insert-1 (20 slots)    %6.1f cycles per warp-insert per SM-share (%.1f cycles/SMSP)\n", warps, c / ((double)warps * iters), c / ((double)warps * iters) * 4);
        c = run([&] { alu_kernel<3><<<sms, threads>>>(out, cyc, iters, 1.0001f, 1e-7f); }, sms, cyc);
        printf("warps/SM %2d  merge-2  (20 slots)    %6.1f cycles per warp-merge(2 cand) per SM-share (%.1f cycles/SMSP)\n", warps, c / ((double)warps * iters), c / ((double)warps * iters) * 4);
        c = run([&] { alu_kernel<4><<<sms, threads>>>(out, cyc, iters, 1.0001f, 1e-7f); }, sms, cyc);
        printf("warps/SM %2d  FFMA   x8 indep        %6.3f warp-instr/cycle/SM\n", warps, (double)warps * iters * 8 / c);
        c = run([&] { alu_kernel<5><<<sms, threads>>>(out, cyc, iters, 1.0001f, 1e-7f); }, sms, cyc);
        printf("warps/SM %2d  FFMA+FMNMX mix        %6.3f warp-instr/cycle/SM\n", warps, (double)warps * iters * 16 / c);
    }
    return 0;
}
