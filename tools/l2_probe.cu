// L2 retention probe (B200): a persistent grid streams a large buffer chunk by chunk; in step i every CTA first-reads its slice of
// chunk i and re-reads ANOTHER CTA's slice of chunk i - lag.  ncu's dram__bytes_read.sum over the launch tells how many of the
// re-reads the L2 served: ideal (all hit) = buffer size, none = 2 x buffer size.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bin/l2_probe l2_probe.cu ; ncu --metrics dram__bytes_read.sum ./l2_probe <chunk MB> <lag> <shift>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512) probe(const uint4* __restrict__ buf, size_t chunk_vec, int n_chunks, int lag, int shift, int first_hint,
                                             uint4* sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    unsigned long long pol_last, pol_first;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    const size_t slice = chunk_vec / gridDim.x;
    for (int i = 0; i < n_chunks + lag; ++i) {
        if (i < n_chunks) {
            const uint4* p = buf + (size_t)i * chunk_vec + (size_t)blockIdx.x * slice;
            for (size_t k = threadIdx.x; k < slice; k += blockDim.x) {
                uint4 v;
                if (first_hint) asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + k), "l"(pol_last));
                else asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + k));
                acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
            }
        }
        if (i >= lag) {
            const int b = (blockIdx.x + shift) % gridDim.x;
            const uint4* p = buf + (size_t)(i - lag) * chunk_vec + (size_t)b * slice;
            for (size_t k = threadIdx.x; k < slice; k += blockDim.x) {
                uint4 v;
                if (first_hint) asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + k), "l"(pol_first));
                else asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + k));
                acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
            }
        }
    }
    if (acc.x == 0x12345678u && acc.y == 1u) sink[0] = acc;
}
int main(int argc, char** argv) {
    const size_t chunk_mb = argc > 1 ? atoi(argv[1]) : 8;
    const int lag = argc > 2 ? atoi(argv[2]) : 2, shift = argc > 3 ? atoi(argv[3]) : 37, hint = argc > 4 ? atoi(argv[4]) : 0;
    const int n_chunks = (int)(1024 / chunk_mb);
    const size_t chunk_vec = chunk_mb * 1024 * 1024 / 16;
    uint4* buf;
    cudaMalloc(&buf, (size_t)n_chunks * chunk_vec * 16 + 1024);
    cudaMemset(buf, 1, (size_t)n_chunks * chunk_vec * 16);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<<<sms, 512>>>(buf, chunk_vec, n_chunks, lag, shift, hint, buf + (size_t)n_chunks * chunk_vec);
    cudaEventRecord(e0);
    probe<<<sms, 512>>>(buf, chunk_vec, n_chunks, lag, shift, hint, buf + (size_t)n_chunks * chunk_vec);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("chunk %zu MB lag %d (%zu MB) shift %d hint %d: %.3f ms, %.1f GB/s of loads, %s\n", chunk_mb, lag, chunk_mb * lag, shift, hint, ms,
           2.0 * n_chunks * chunk_mb / 1024.0 / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
    return 0;
}
