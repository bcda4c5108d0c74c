// byte order and rounding statistics of cvt.rs.satfinite.e4m3x4.f32 (stochastic rounding, sm_100a)
#include <cstdio>
#include <cstdint>
#include <cuda_fp8.h>
__global__ void k(uint32_t* y, float* mean) {
    uint32_t d;
    asm volatile("cvt.rs.satfinite.e4m3x4.f32 %0, {%1, %2, %3, %4}, %5;" : "=r"(d) : "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(4.0f), "r"(0x12345678u));
    y[0] = d;
    // 1.3 lies between e4m3 1.25 and 1.375: expected mean under stochastic rounding = 1.3
    float s = 0.f;
    for (uint32_t i = 0; i < 4096; ++i) {
        uint32_t r = i * 0x9E3779B1u; r ^= r >> 15; r *= 0x2C1B3C6Du; r ^= r >> 12;
        asm volatile("cvt.rs.satfinite.e4m3x4.f32 %0, {%1, %2, %3, %4}, %5;" : "=r"(d) : "f"(1.3f), "f"(1.3f), "f"(1.3f), "f"(1.3f), "r"(r));
        for (int b = 0; b < 4; ++b) {
            __nv_fp8_e4m3 v; v.__x = (d >> (8 * b)) & 0xff;
            s += float(v);
        }
    }
    mean[0] = s / (4096 * 4);
}
int main() {
    uint32_t* y; float* m;
    cudaMallocManaged(&y, 4); cudaMallocManaged(&m, 4);
    k<<<1, 1>>>(y, m);
    cudaDeviceSynchronize();
    printf("packed {1,2,3,4} = 0x%08x (e4m3: 1=0x38 2=0x40 3=0x44 4=0x48)  mean of SR(1.3) = %f  %s\n", y[0], m[0], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
