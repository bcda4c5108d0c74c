// Per-SM issue rate of the instructions the conv5 fp8 epilogue is made of (one CTA of 16 warps on one SM, 8 independent
// chains per thread).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/alu_probe tools/alu_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void probe(uint32_t* out, long long* cyc, int iters) {
    uint32_t r[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { r[i] = threadIdx.x * 2654435761u + i * 40503u; f[i] = 1.0f + 0.001f * (threadIdx.x + i); }
    const float c = 0.5f + 1e-6f * threadIdx.x;
    unsigned long long d[8], cc;
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(d[i]) : "f"(f[i]), "f"(c));
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c));
            if (OP == 1) asm volatile("cvt.rs.satfinite.e4m3x4.f32 %0, {%1, %2, %3, %4}, %0;" : "+r"(r[i]) : "f"(f[i]), "f"(c), "f"(f[i]), "f"(c));
            if (OP == 2) { uint16_t h; asm volatile("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(h) : "f"(f[i]), "f"(__uint_as_float(r[i] & 0x3fffffffu))); r[i] = h; }
            if (OP == 3) asm volatile("xor.b32 %0, %0, %1;" : "+r"(r[i]) : "r"(0x9E3779B1u + i));
            if (OP == 4) asm volatile("shr.u32 %0, %0, 1;" : "+r"(r[i]));
            if (OP == 5) asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(0x9E3779B1u));
            if (OP == 6) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(c));
            if (OP == 7) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c));
            if (OP == 8) asm volatile("prmt.b32 %0, %0, %1, 0x7531;" : "+r"(r[i]) : "r"(0x12345678u));
            if (OP == 9) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(f[i]), "f"(__uint_as_float(r[i] & 0x3fffffffu)));
            if (OP == 10) { uint16_t h; asm volatile("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(h) : "r"(r[i] & 0x3bff3bffu)); r[i] = h; }
            if (OP == 11) asm volatile("cvt.rs.f16x2.f32 %0, %1, %2, %0;" : "+r"(r[i]) : "f"(f[i]), "f"(c));
            if (OP == 12) asm volatile("mul.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c));
            if (OP == 13) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(cc));
        }
    }
    const long long t1 = clock64();
    uint32_t a = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a ^= r[i] ^ __float_as_uint(f[i]) ^ (uint32_t)d[i] ^ (uint32_t)(d[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4 * 512 * 148); cudaMalloc(&cyc, 8 * 148);
    const int iters = 4096;
    for (int warps : {4, 8, 16}) {
        probe<OP><<<1, warps * 32>>>(out, cyc, iters);
        probe<OP><<<1, warps * 32>>>(out, cyc, iters);
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double inst = (double)iters * 8 * warps;
        printf("%-34s warps %2d  %.2f clk per warp-instruction per SMSP  (%.1f lanes/clk/SM)\n", name, warps, c / (inst / 4), inst * 32 / c);
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("max.f32 (FMNMX)");
    run<1>("cvt.rs.satfinite.e4m3x4.f32");
    run<2>("cvt.rn.satfinite.e4m3x2.f32");
    run<10>("cvt.rn.satfinite.e4m3x2.f16x2");
    run<11>("cvt.rs.f16x2.f32");
    run<9>("cvt.rn.bf16x2.f32");
    run<3>("xor.b32 (LOP3)");
    run<4>("shr.u32 (SHF)");
    run<8>("prmt.b32");
    run<5>("mul.lo.u32 (IMAD)");
    run<6>("fma.rn.f32");
    run<7>("add.f32");
    run<12>("mul.f32");
    run<13>("fma.rn.f32x2 (FFMA2)");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
