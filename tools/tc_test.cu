// Stand-alone bring-up harness for the tcgen05 / TMEM / TMA GEMM template of epc-net_b200/csrc/tc_gemm.cuh (no torch):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I epc-net_b200/csrc tools/tc_test.cu -o /tmp/tc_test
// Every configuration the EPC-Net head uses is run on exactly-representable inputs and compared with a CPU
// reference (bit for bit where the arithmetic is exact), then timed at the production shape.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "tc_gemm.cuh"

namespace epc {   // stand-alone stubs for what api.cu provides inside the library
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); printf("\n"); }
void count_launch(int) {}
ScopedStage::ScopedStage(int i, cudaStream_t s) : id(i), st(s), on(false) {}
ScopedStage::~ScopedStage() {}
}  // namespace epc
using namespace epc;
typedef __nv_bfloat16 bf16;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static uint32_t g_seed = 123;
static float rnd_exact(int range) {   // multiples of 1/8: exact in TF32 and bf16, products/sums exact in fp32
    g_seed = g_seed * 1664525u + 1013904223u;
    return (float)((int)((g_seed >> 10) % (2 * range + 1)) - range) / 8.0f;
}
static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
template <typename T> T* to_dev(const std::vector<T>& h) { T* d; CK(cudaMalloc(&d, h.size() * sizeof(T))); CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); return d; }
static std::vector<bf16> to_bf16(const std::vector<float>& v) { std::vector<bf16> o(v.size()); for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16(v[i]); return o; }

template <typename F> float time_ms(F f, int reps = 20) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}
static int sync_ok(const char* what) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: kernel failed: %s\n", what, cudaGetErrorString(e)); return 0; }
    return 1;
}

// ---- T1: tf32, K-major x K-major, fp32 store (+bias, relu) ---------------------------------------------------------
template <int BN> int test_tf32_store(int M, int N, int K, bool timing) {
    std::vector<float> A((size_t)M * K), B((size_t)N * K), bias(N), C((size_t)M * N);
    for (auto& v : A) v = rnd_exact(8);
    for (auto& v : B) v = rnd_exact(8);
    for (auto& v : bias) v = rnd_exact(16);
    float *dA = to_dev(A), *dB = to_dev(B), *db = to_dev(bias), *dC; CK(cudaMalloc(&dC, C.size() * 4)); CK(cudaMemset(dC, 0xff, C.size() * 4));
    tc::GemmParams p = {}; p.M = M; p.N = N; p.K = K; p.splitk = 1; p.C = dC; p.ldc = N; p.bias = db; p.relu = 1;
    Operand<float> oa{dA, M, K, K}, ob{dB, N, K, K};
    auto run = [&]() { return tc_gemm_launch<float, BN, false, false, tc::EPI_STORE_F32>(oa, ob, p, 1, 0, 2); };
    if (run() || !sync_ok("tf32_store")) return 1;
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0; const int mstep = M > 1024 ? 509 : 1;
    for (int m = 0; m < M; m += mstep) for (int n = 0; n < N; ++n) {
        float acc = 0.f; for (int k = 0; k < K; ++k) acc += A[(size_t)m * K + k] * B[(size_t)n * K + k];
        acc = fmaxf(acc + bias[n], 0.f);
        if (C[(size_t)m * N + n] != acc) { if (bad < 3) printf("  (%d,%d) got %g want %g\n", m, n, C[(size_t)m * N + n], acc); ++bad; }
    }
    printf("T1 tf32 store   M=%d N=%d K=%d BN=%d : %s\n", M, N, K, BN, bad ? "FAIL" : "exact");
    if (timing && !bad) { float ms = time_ms(run); printf("    %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * M * N * K / ms * 1e-9); }
    cudaFree(dA); cudaFree(dB); cudaFree(db); cudaFree(dC);
    return bad != 0;
}

// ---- T2: bf16 conv5 epilogue: H = bf16(relu(A B^T + b)), rowss partials ---------------------------------------------
template <int BN, bool PERSIST = false> int test_bf16_conv5(int M, int N, int K, bool timing) {
    std::vector<float> A((size_t)M * K), B((size_t)N * K), bias(N);
    for (auto& v : A) v = rnd_exact(8);
    for (auto& v : B) v = rnd_exact(8);
    for (auto& v : bias) v = rnd_exact(16);
    bf16 *dA = to_dev(to_bf16(A)), *dB = to_dev(to_bf16(B)), *dH; float* db = to_dev(bias); float* dss;
    const int parts = (N / BN) * (PERSIST ? 2 : 1);
    CK(cudaMalloc(&dH, (size_t)M * N * 2)); CK(cudaMalloc(&dss, (size_t)M * parts * 4));
    tc::GemmParams p = {}; p.M = M; p.N = N; p.K = K; p.splitk = 1; p.C = dH; p.ldc = N; p.bias = db; p.relu = 1; p.aux = dss;
    Operand<bf16> oa{dA, M, K, K}, ob{dB, N, K, K};
    auto run = [&]() { return PERSIST ? tc_gemm_bres_launch<bf16, BN, tc::EPI_CONV5_BF16, 8>(oa, ob, p, 0)
                                      : tc_gemm_launch<bf16, BN, false, false, tc::EPI_CONV5_BF16>(oa, ob, p, 1, 0, 2); };
    if (run() || !sync_ok("bf16_conv5")) return 1;
    std::vector<bf16> H((size_t)M * N); std::vector<float> ss((size_t)M * parts);
    CK(cudaMemcpy(H.data(), dH, H.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ss.data(), dss, ss.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0; const int mstep = M > 1024 ? 509 : 1;
    for (int m = 0; m < M; m += mstep) {
        std::vector<double> s2(parts, 0.0);
        for (int n = 0; n < N; ++n) {
            float acc = 0.f; for (int k = 0; k < K; ++k) acc += A[(size_t)m * K + k] * B[(size_t)n * K + k];
            acc = fmaxf(acc + bias[n], 0.f);
            s2[PERSIST ? n / (BN / 2) : n / BN] += (double)acc * acc;
            if (__bfloat162float(H[(size_t)m * N + n]) != bf16_round(acc)) { if (bad < 3) printf("  H(%d,%d) got %g want %g\n", m, n, __bfloat162float(H[(size_t)m * N + n]), bf16_round(acc)); ++bad; }
        }
        for (int q = 0; q < parts; ++q) if (fabs(ss[(size_t)m * parts + q] - s2[q]) > 1e-5 * (1 + s2[q])) { if (bad < 3) printf("  rowss(%d,%d) got %g want %g\n", m, q, ss[(size_t)m * parts + q], s2[q]); ++bad; }
    }
    printf("T2 bf16 conv5%s M=%d N=%d K=%d BN=%d : %s\n", PERSIST ? "(P)" : "   ", M, N, K, BN, bad ? "FAIL" : "exact");
    if (timing && !bad) { float ms = time_ms(run); printf("    %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * M * N * K / ms * 1e-9); }
    cudaFree(dA); cudaFree(dB); cudaFree(db); cudaFree(dH); cudaFree(dss);
    return bad != 0;
}

// ---- T3: bf16 assignment epilogue (N = 64) -----------------------------------------------------------------------
int test_bf16_assign(int M, int K, bool timing, bool persist = false) {
    const int N = 64, parts = 4;
    std::vector<float> A((size_t)M * K), B((size_t)N * K), rowss((size_t)M * parts), sc(N), sh(N);
    for (auto& v : A) v = fabsf(rnd_exact(8));
    for (auto& v : B) v = rnd_exact(8);
    for (auto& v : rowss) v = 4.0f + fabsf(rnd_exact(64));
    for (auto& v : sc) v = 0.5f + fabsf(rnd_exact(8));
    for (auto& v : sh) v = rnd_exact(8);
    bf16 *dA = to_dev(to_bf16(A)), *dB = to_dev(to_bf16(B)), *dS; float *dr = to_dev(rowss), *dsc = to_dev(sc), *dsh = to_dev(sh), *dap;
    const int tiles = (M + 127) / 128;
    CK(cudaMalloc(&dS, (size_t)M * 64 * 2)); CK(cudaMalloc(&dap, (size_t)tiles * 64 * 4));
    tc::GemmParams p = {}; p.M = M; p.N = N; p.K = K; p.splitk = 1; p.C = dS; p.ldc = 64; p.aux = dap; p.rowss = dr; p.rowss_parts = parts;
    p.bn_scale = dsc; p.bn_shift = dsh;
    Operand<bf16> oa{dA, M, K, K}, ob{dB, N, K, K};
    auto run = [&]() { return persist ? tc_gemm_bres_launch<bf16, 64, tc::EPI_ASSIGN>(oa, ob, p, 0)
                                      : tc_gemm_launch<bf16, 64, false, false, tc::EPI_ASSIGN>(oa, ob, p, 1, 0, 1); };
    if (run() || !sync_ok("bf16_assign")) return 1;
    std::vector<bf16> S((size_t)M * 64); std::vector<float> ap((size_t)tiles * 64);
    CK(cudaMemcpy(S.data(), dS, S.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ap.data(), dap, ap.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0; const int tstep = tiles > 8 ? 97 : 1;
    for (int t = 0; t < tiles; t += tstep) {
        std::vector<double> colsum(64, 0.0);
        for (int r = 0; r < 128 && t * 128 + r < M; ++r) {
            const int m = t * 128 + r;
            double ssq = 0; for (int q = 0; q < parts; ++q) ssq += rowss[(size_t)m * parts + q];
            const double inv = 1.0 / sqrt(ssq);
            double l[64], mx = -1e30, den = 0;
            for (int n = 0; n < 64; ++n) {
                double acc = 0; for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
                l[n] = acc * inv * sc[n] + sh[n]; mx = fmax(mx, l[n]);
            }
            for (int n = 0; n < 64; ++n) { l[n] = exp(l[n] - mx); den += l[n]; }
            for (int n = 0; n < 64; ++n) {
                const double a = l[n] / den; colsum[n] += a;
                const double want = a * inv, got = __bfloat162float(S[(size_t)m * 64 + n]);
                if (fabs(got - want) > 0.01 * fabs(want) + 1e-7) { if (bad < 3) printf("  S(%d,%d) got %g want %g\n", m, n, got, want); ++bad; }
            }
        }
        for (int n = 0; n < 64; ++n) if (fabs(ap[(size_t)t * 64 + n] - colsum[n]) > 1e-4 * (1 + colsum[n])) { if (bad < 3) printf("  a_part(%d,%d) got %g want %g\n", t, n, ap[(size_t)t * 64 + n], colsum[n]); ++bad; }
    }
    printf("T3 bf16 assign%s M=%d K=%d : %s\n", persist ? "(P)" : "   ", M, K, bad ? "FAIL" : "ok");
    if (timing && !bad) { float ms = time_ms(run); printf("    %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * M * N * K / ms * 1e-9); }
    cudaFree(dA); cudaFree(dB); cudaFree(dS); cudaFree(dr); cudaFree(dsc); cudaFree(dsh); cudaFree(dap);
    return bad != 0;
}

// ---- T4: bf16 MN-major x MN-major, batched + split-K: V[b] = H[b]^T S[b] --------------------------------------------
int test_bf16_vlad(int batch, int Npts, int F, int splitk, bool timing) {
    const int C = 64;
    std::vector<float> H((size_t)batch * Npts * F), S((size_t)batch * Npts * C);
    for (auto& v : H) v = fabsf(rnd_exact(8));
    for (auto& v : S) v = fabsf(rnd_exact(4));
    bf16 *dH = to_dev(to_bf16(H)), *dS = to_dev(to_bf16(S)); float* dV;
    const size_t slab = (size_t)batch * F * C;
    CK(cudaMalloc(&dV, slab * splitk * 4)); CK(cudaMemset(dV, 0xff, slab * splitk * 4));
    tc::GemmParams p = {}; p.M = F; p.N = C; p.K = Npts / splitk; p.k_batch_rows = Npts; p.splitk = splitk; p.C = dV; p.ldc = C;
    p.c_batch = (long long)F * C; p.c_slab = (long long)slab;
    Operand<bf16> oa{dH, (long long)batch * Npts, F, F}, ob{dS, (long long)batch * Npts, C, C};
    auto run = [&]() { return tc_gemm_launch<bf16, 64, true, true, tc::EPI_STORE_F32>(oa, ob, p, batch, 0, 1); };
    if (run() || !sync_ok("bf16_vlad")) return 1;
    std::vector<float> V(slab * splitk);
    CK(cudaMemcpy(V.data(), dV, V.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0; const int fstep = F > 256 ? 37 : 1;
    for (int b = 0; b < batch; ++b) for (int f = 0; f < F; f += fstep) for (int c = 0; c < C; ++c) {
        float acc = 0.f; for (int n = 0; n < Npts; ++n) acc += H[((size_t)b * Npts + n) * F + f] * S[((size_t)b * Npts + n) * C + c];
        float got = 0.f; for (int s = 0; s < splitk; ++s) got += V[s * slab + ((size_t)b * F + f) * C + c];
        if (got != acc) { if (bad < 3) printf("  V(%d,%d,%d) got %g want %g\n", b, f, c, got, acc); ++bad; }
    }
    printf("T4 bf16 vlad    batch=%d N=%d F=%d splitk=%d : %s\n", batch, Npts, F, splitk, bad ? "FAIL" : "exact");
    if (timing && !bad) { float ms = time_ms(run); printf("    %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * batch * F * C * Npts / ms * 1e-9); }
    cudaFree(dH); cudaFree(dS); cudaFree(dV);
    return bad != 0;
}

// ---- T5: tf32 column-max epilogue ---------------------------------------------------------------------------------
template <int BN> int test_tf32_colmax(int clouds, int Npts, int N, int K, bool timing, bool persist = false) {
    const int M = clouds * Npts;
    std::vector<float> A((size_t)M * K), B((size_t)N * K), bias(N), G((size_t)clouds * N);
    for (auto& v : A) v = rnd_exact(8);
    for (auto& v : B) v = rnd_exact(8);
    for (auto& v : bias) v = rnd_exact(16);
    float *dA = to_dev(A), *dB = to_dev(B), *db = to_dev(bias), *dG; CK(cudaMalloc(&dG, G.size() * 4));
    tc::GemmParams p = {}; p.M = M; p.N = N; p.K = K; p.splitk = 1; p.bias = db; p.aux = dG; p.rows_per_cloud = Npts;
    Operand<float> oa{dA, M, K, K}, ob{dB, N, K, K};
    auto run = [&]() { cudaMemsetAsync(dG, 0, G.size() * 4, 0);
                       return persist ? tc_gemm_bres_launch<float, BN, tc::EPI_COLMAX, 8>(oa, ob, p, 0)
                                      : tc_gemm_launch<float, BN, false, false, tc::EPI_COLMAX>(oa, ob, p, 1, 0, 2); };
    if (run() || !sync_ok("tf32_colmax")) return 1;
    CK(cudaMemcpy(G.data(), dG, G.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0; const int nstep = M > 4096 ? 61 : 1;
    for (int b = 0; b < clouds; ++b) for (int n = 0; n < N; n += nstep) {
        float mx = 0.f;
        for (int r = 0; r < Npts; ++r) { float acc = 0.f; const float* a = &A[((size_t)b * Npts + r) * K]; for (int k = 0; k < K; ++k) acc += a[k] * B[(size_t)n * K + k]; mx = fmaxf(mx, acc + bias[n]); }
        if (G[(size_t)b * N + n] != mx) { if (bad < 3) printf("  g(%d,%d) got %g want %g\n", b, n, G[(size_t)b * N + n], mx); ++bad; }
    }
    printf("T5 tf32 colmax%s clouds=%d N=%d K=%d BN=%d : %s\n", persist ? "(P)" : "   ", clouds, N, K, BN, bad ? "FAIL" : "exact");
    if (timing && !bad) { float ms = time_ms(run); printf("    %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * M * N * K / ms * 1e-9); }
    cudaFree(dA); cudaFree(dB); cudaFree(db); cudaFree(dG);
    return bad != 0;
}

int main() {
    int fails = 0;
    fails += test_tf32_store<128>(256, 256, 256, false);
    fails += test_tf32_store<64>(384, 64, 1024, false);
    fails += test_bf16_conv5<256>(256, 512, 256, false);
    fails += test_bf16_conv5<128>(300, 256, 128, false);
    fails += test_bf16_assign(300, 1024, false);
    fails += test_bf16_vlad(2, 256, 128, 1, false);
    fails += test_bf16_vlad(2, 512, 256, 2, false);
    fails += test_tf32_colmax<256>(2, 256, 256, 128, false);
    fails += test_bf16_conv5<256, true>(256, 512, 256, false);
    fails += test_bf16_conv5<256, true>(40000, 1024, 256, false);
    fails += test_bf16_assign(300, 1024, false, true);
    fails += test_bf16_assign(40000, 1024, false, true);
    fails += test_tf32_colmax<256>(2, 256, 256, 128, false, true);
    fails += test_tf32_colmax<256>(3, 4096, 1024, 128, false, true);
    printf(fails ? "SOME CORRECTNESS TESTS FAILED (%d)\n" : "ALL CORRECT\n", fails);
    // production shapes (32 clouds x 4096 points)
    test_tf32_store<256>(131072, 1024, 256, true);
    test_bf16_conv5<256>(131072, 1024, 256, true);
    test_bf16_conv5<128>(131072, 1024, 256, true);
    test_bf16_assign(131072, 1024, true);
    test_bf16_vlad(32, 4096, 1024, 1, true);
    test_bf16_vlad(32, 4096, 1024, 2, true);
    test_tf32_colmax<256>(32, 4096, 1024, 128, true);
    printf("-- persistent B-resident variants, 8 clouds (the head's L2-resident sub-batch) and 32 clouds\n");
    test_bf16_conv5<256, false>(32768, 1024, 256, true);
    test_bf16_conv5<256, true>(32768, 1024, 256, true);
    test_bf16_conv5<256, true>(131072, 1024, 256, true);
    test_bf16_assign(32768, 1024, true, false);
    test_bf16_assign(32768, 1024, true, true);
    test_tf32_colmax<256>(32, 4096, 1024, 128, true, true);
    return fails;
}
