// Exact-fp32 (FFMA) strided batched GEMM with fused bias/ReLU epilogue and split-K.
// Used for the small / bandwidth-bound contractions of the head (hidden FC, gating, FC of EPC-Net-L,
// retrieval scoring) and as the on-device fp32 reference for the tensor-core kernels.
#include "common.cuh"
#include "kernels.h"

namespace epc {

constexpr int GT = 64;     // C tile (GT x GT)
constexpr int GK = 16;     // K step
constexpr int GLD = GT + 4;

// A_K: A is k-contiguous (sAk == 1) else m-contiguous (sAm == 1); same for B (B_K: sBk == 1 else sBn == 1).
template <bool A_K, bool B_K>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[GK][GLD];
    __shared__ __align__(16) float Bs[GK][GLD];
    const int tid = threadIdx.x;
    const int z = blockIdx.z;
    const int batch = z / g.splitk, split = z % g.splitk;
    const float* A = g.A + (long long)batch * g.bA;
    const float* B = g.B + (long long)batch * g.bB;
    float* C = g.C + (long long)batch * g.bC + (long long)split * g.slab;
    const int m0 = blockIdx.x * GT, n0 = blockIdx.y * GT;
    const int ksteps = (g.K + GK - 1) / GK;
    const int per = (ksteps + g.splitk - 1) / g.splitk;
    const int ks0 = split * per, ks1 = min(ksteps, ks0 + per);

    const int tx = tid & 15, ty = tid >> 4;     // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4] = {};

    for (int ks = ks0; ks < ks1; ++ks) {
        const int k0 = ks * GK;
        // ---- A tile: GT(m) x GK(k) -> As[k][m]
        if (A_K) {
            const int m = tid >> 2, k4 = (tid & 3) * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m0 + m < g.M) {
                const float* src = A + (long long)(m0 + m) * g.sAm + (k0 + k4);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (k0 + k4 + i < g.K) v[i] = __ldg(src + i);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) As[k4 + i][m] = v[i];
        } else {
            const int k = tid >> 4, m4 = (tid & 15) * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (k0 + k < g.K) {
                const float* src = A + (long long)(k0 + k) * g.sAk + (m0 + m4);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (m0 + m4 + i < g.M) v[i] = __ldg(src + i);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) As[k][m4 + i] = v[i];
        }
        // ---- B tile: GK(k) x GT(n) -> Bs[k][n]
        if (B_K) {
            const int n = tid >> 2, k4 = (tid & 3) * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (n0 + n < g.N) {
                const float* src = B + (long long)(n0 + n) * g.sBn + (k0 + k4);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (k0 + k4 + i < g.K) v[i] = __ldg(src + i);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[k4 + i][n] = v[i];
        } else {
            const int k = tid >> 4, n4 = (tid & 15) * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (k0 + k < g.K) {
                const float* src = B + (long long)(k0 + k) * g.sBk + (n0 + n4);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (n0 + n4 + i < g.N) v[i] = __ldg(src + i);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[k][n4 + i] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            float* dst = C + (long long)m * g.ldc + n;
            if (g.splitk == 1) {
                if (g.bias) v += g.bias[n];
                if (g.relu) v = fmaxf(v, 0.f);
            }
            *dst = v;
        }
    }
}

int sgemm(const GemmArgs& g0, cudaStream_t st) {
    GemmArgs g = g0;
    if (g.batch <= 0) g.batch = 1;
    if (g.splitk <= 0) g.splitk = 1;
    EPC_CHECK_ARG(g.sAk == 1 || g.sAm == 1, "sgemm: A must be contiguous along m or k");
    EPC_CHECK_ARG(g.sBk == 1 || g.sBn == 1, "sgemm: B must be contiguous along k or n");
    EPC_CHECK_ARG(!(g.splitk > 1 && (g.bias || g.relu)), "sgemm: split-K cannot fuse bias/relu");
    if (g.M == 0 || g.N == 0) return EPC_OK;
    dim3 grid((g.M + GT - 1) / GT, (g.N + GT - 1) / GT, g.batch * g.splitk);
    EPC_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "sgemm: grid too large (N=%d batch=%d)", g.N, g.batch);
    const bool ak = (g.sAk == 1), bk = (g.sBk == 1);
    if (ak && bk)
        sgemm_kernel<true, true><<<grid, 256, 0, st>>>(g);
    else if (ak && !bk)
        sgemm_kernel<true, false><<<grid, 256, 0, st>>>(g);
    else if (!ak && bk)
        sgemm_kernel<false, true><<<grid, 256, 0, st>>>(g);
    else
        sgemm_kernel<false, false><<<grid, 256, 0, st>>>(g);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

}  // namespace epc
