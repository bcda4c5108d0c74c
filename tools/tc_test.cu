// Stand-alone bring-up harness for the tcgen05 / TMEM / TMA building blocks (no torch):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I epc-net_b200/csrc tools/tc_test.cu -o /tmp/tc_test -lcuda
// Runs the TF32 GEMM kernels of epc-net_b200/csrc/tc_gemm.cuh on exactly-representable inputs and compares
// against a CPU reference bit for bit, then times them.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stdarg.h>
#include <vector>
#include "tc_gemm.cuh"

namespace epc {   // stand-alone stubs for what api.cu provides inside the library
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); printf("\n"); }
void count_launch(int) {}
ScopedStage::ScopedStage(int i, cudaStream_t s) : id(i), st(s), on(false) {}
ScopedStage::~ScopedStage() {}
}  // namespace epc

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static float rnd_exact(uint32_t& s, int range) {   // small multiples of 1/8: exact in TF32, sums exact in fp32
    s = s * 1664525u + 1013904223u;
    return (float)((int)((s >> 10) % (2 * range + 1)) - range) / 8.0f;
}

int test_nt(int M, int N, int K, int BN, bool relu, bool timing) {
    std::vector<float> A((size_t)M * K), B((size_t)N * K), bias(N), C((size_t)M * N), R((size_t)M * N);
    uint32_t s = 123;
    for (auto& v : A) v = rnd_exact(s, 8);
    for (auto& v : B) v = rnd_exact(s, 8);
    for (auto& v : bias) v = rnd_exact(s, 16);
    const int mstep = (M > 1024) ? 509 : 1;        // big problems: check a strided subset of rows
    for (int m = 0; m < M; m += mstep)
        for (int n = 0; n < N; ++n) {
            float acc = 0.f;
            for (int k = 0; k < K; ++k) acc += A[(size_t)m * K + k] * B[(size_t)n * K + k];
            acc += bias[n];
            R[(size_t)m * N + n] = relu ? fmaxf(acc, 0.f) : acc;
        }
    float *dA, *dB, *dbias, *dC;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dbias, N * 4)); CK(cudaMalloc(&dC, C.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbias, bias.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xff, C.size() * 4));
    epc::TcGemmNT g;
    g.A = dA; g.B = dB; g.C = dC; g.bias = dbias; g.M = M; g.N = N; g.K = K; g.ldc = N; g.relu = relu ? 1 : 0; g.BN = BN;
    int rc = epc::tc_gemm_nt(g, 0);
    if (rc) { printf("tc_gemm_nt launch failed rc=%d\n", rc); return 1; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0; double maxerr = 0;
    for (size_t i = 0; i < C.size(); ++i) {
        if ((i / N) % mstep) continue;
        double d = fabs((double)C[i] - (double)R[i]);
        if (!(d == 0)) { if (bad < 5) printf("  mismatch at (%zu,%zu): got %g want %g\n", i / N, i % N, C[i], R[i]); ++bad; }
        if (d > maxerr || d != d) maxerr = d;
    }
    printf("tc_gemm_nt M=%d N=%d K=%d BN=%d relu=%d : %s (%zu mismatches, max err %g)\n", M, N, K, BN, (int)relu,
           bad ? "FAIL" : "exact", bad, maxerr);
    if (timing && !bad) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 3; ++i) epc::tc_gemm_nt(g, 0);
        cudaEventRecord(e0);
        const int reps = 20;
        for (int i = 0; i < reps; ++i) epc::tc_gemm_nt(g, 0);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("   %.3f ms/launch  %.1f TFLOP/s (tf32)\n", ms / reps, 2.0 * M * N * K / (ms / reps) * 1e-9);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dbias); cudaFree(dC);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    int fails = 0;
    fails += test_nt(128, 128, 32, 128, false, false);
    fails += test_nt(128, 128, 128, 128, false, false);
    fails += test_nt(256, 256, 256, 128, true, false);
    fails += test_nt(256, 256, 256, 256, true, false);
    fails += test_nt(384, 64, 1024, 64, false, false);
    if (!fails) {
        test_nt(131072, 1024, 256, 256, true, true);     // conv5 of 32 clouds
        test_nt(131072, 1024, 256, 128, true, true);
        test_nt(131072, 64, 1024, 64, false, true);      // assignment logits of 32 clouds
    }
    printf(fails ? "SOME TESTS FAILED\n" : "ALL OK\n");
    return fails;
}
