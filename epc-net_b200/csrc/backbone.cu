// K2 -- ProxyConv backbone (models/epc-net.py:62-132, models/epc-net-l.py:44-80).
//
// The reference realises  m_i = (1/20) sum_{j in N(i)} x_j  as a dense (N x N mask) x (N x 64) batch
// matmul per block (2.15 GFLOP and a 64 MiB read each).  Here the neighbour lists of K1 are gathered
// directly (20 x 256 B rows per point, L2/L1 resident thanks to the Morton order), and the block body
//     t = m - x ; t = conv_a(t) ; t = conv_b(t) ; out = t + m ; x' = conv_{b+1}(out)
// runs on the tile while it is in shared memory.  Rows whose thresholded set has > 20 members (ties at
// the 20th distance, utils/tf_util.py:663-665) take an exact dense re-scan of the cloud.
#include "common.cuh"
#include "kernels.h"

namespace epc {

// x0 = relu(BN(p W1 + b1)), cin = 3  (models/epc-net.py:66-69)
__global__ void conv_in_kernel(const float4* __restrict__ sorted, long long R, const float* __restrict__ W,
                               const float* __restrict__ bias, float* __restrict__ x) {
    __shared__ float sW[3 * 64 + 64];
    for (int i = threadIdx.x; i < 3 * 64 + 64; i += blockDim.x) sW[i] = (i < 192) ? W[i] : bias[i - 192];
    __syncthreads();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t >> 4;
    const int c4 = (int)(t & 15) * 4;
    if (r >= R) return;
    const float4 p = sorted[r];
    float4 o;
    float* op = reinterpret_cast<float*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c4 + i;
        float acc = sW[192 + c];
        acc = fmaf(p.x, sW[c], acc);
        acc = fmaf(p.y, sW[64 + c], acc);
        acc = fmaf(p.z, sW[128 + c], acc);
        op[i] = fmaxf(acc, 0.f);
    }
    *reinterpret_cast<float4*>(x + r * 64 + c4) = o;
}

int conv_in(const float4* sorted, long long R, const DenseDev& L, float* x, cudaStream_t st) {
    EPC_CHECK_ARG(L.cin == 3 && L.cout == 64, "conv_in expects a 3->64 layer, got %d->%d", L.cin, L.cout);
    if (R == 0) return EPC_OK;
    const long long threads = R * 16;
    conv_in_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(sorted, R, L.W, L.b, x);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

constexpr int PB_TILE = 64;       // points per CTA
constexpr int PB_LD = 65;         // smem row stride (floats): conflict-free column walks
constexpr int PB_THREADS = 256;

// One 64->64 pointwise layer on the smem tile: out[p][c] = relu(b[c] + sum_k in[p][k] W[k][c]) (+ res[p][c]).
// Warp w: points (w&1)*32 + lane, output chunk (w>>1)*16 .. +16  => weight reads are warp broadcasts.
__device__ __forceinline__ void tile_dense64(const float* __restrict__ sIn, const float* __restrict__ sW,
                                             const float* __restrict__ sB, const float* __restrict__ sRes,
                                             float* __restrict__ sOut, int warp, int lane) {
    const int p = (warp & 1) * 32 + lane;
    const int c0 = (warp >> 1) * 16;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = sB[c0 + i];
#pragma unroll 8
    for (int k = 0; k < 64; ++k) {
        const float a = sIn[p * PB_LD + k];
        const float4* w4 = reinterpret_cast<const float4*>(sW + k * 64 + c0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 w = w4[q];
            acc[4 * q + 0] = fmaf(a, w.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(a, w.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(a, w.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(a, w.w, acc[4 * q + 3]);
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float v = fmaxf(acc[i], 0.f);
        if (sRes) v += sRes[p * PB_LD + c0 + i];
        sOut[p * PB_LD + c0 + i] = v;
    }
}

template <bool HAS_NEXT>
__global__ void __launch_bounds__(PB_THREADS)
proxy_block_kernel(const float* __restrict__ x, const uint16_t* __restrict__ nbr, const float* __restrict__ kthd,
                   const int* __restrict__ cnt, const float4* __restrict__ sorted, int N, int arith, float divisor,
                   const float* __restrict__ Wa, const float* __restrict__ ba, const float* __restrict__ Wb,
                   const float* __restrict__ bb, const float* __restrict__ Wn, const float* __restrict__ bn,
                   float* __restrict__ concat, __nv_bfloat16* __restrict__ concat16, int ctot, int coff,
                   float* __restrict__ xnext) {
    extern __shared__ __align__(16) float smem[];
    float* sWa = smem;                    // [64][64]
    float* sWb = sWa + 4096;
    float* sWn = sWb + 4096;
    float* sBias = sWn + 4096;            // [3][64]
    float* sT = sBias + 192;              // [64][65]
    float* sU = sT + PB_TILE * PB_LD;
    float* sM = sU + PB_TILE * PB_LD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int tile0 = blockIdx.x * PB_TILE;

    for (int i = tid; i < 4096; i += PB_THREADS) {
        sWa[i] = Wa[i];
        sWb[i] = Wb[i];
        if (HAS_NEXT) sWn[i] = Wn[i];
    }
    if (tid < 64) {
        sBias[tid] = ba[tid];
        sBias[64 + tid] = bb[tid];
        if (HAS_NEXT) sBias[128 + tid] = bn[tid];
    }

    // ---- gather-mean: warp per point, lane = channel pair ------------------------------------------
    const float* xb = x + (size_t)b * N * 64;
    for (int pl = warp; pl < PB_TILE; pl += PB_THREADS / 32) {
        const int pos = tile0 + pl;
        float2 m = make_float2(0.f, 0.f), t = make_float2(0.f, 0.f);
        if (pos < N) {
            const size_t row = (size_t)b * N + pos;
            const int c = cnt[row];
            float2 acc = make_float2(0.f, 0.f);
            if (c == KNN_K) {
                const int mine = (lane < KNN_K) ? (int)nbr[row * KNN_K + lane] : 0;
                float2 v[KNN_K];
#pragma unroll
                for (int q = 0; q < KNN_K; ++q) {
                    const int j = __shfl_sync(FULL, mine, q);
                    v[q] = __ldg(reinterpret_cast<const float2*>(xb + (size_t)j * 64) + lane);
                }
#pragma unroll
                for (int q = 0; q < KNN_K; ++q) {
                    acc.x += v[q].x;
                    acc.y += v[q].y;
                }
            } else {
                // ties at the 20th distance: the set is {j : d_ij <= kthd_i}; re-scan the cloud exactly
                const float thr = kthd[row];
                const float4 qp = sorted[row];
                const float4* sp = sorted + (size_t)b * N;
                for (int j0 = 0; j0 < N; j0 += 32) {
                    const float4 pj = sp[j0 + lane];
                    const float d = (arith == EPC_KNN_ARITH_MULADD)
                                        ? canon_dist<0>(qp.x, qp.y, qp.z, qp.w, pj.x, pj.y, pj.z, pj.w)
                                        : canon_dist<1>(qp.x, qp.y, qp.z, qp.w, pj.x, pj.y, pj.z, pj.w);
                    unsigned mk = __ballot_sync(FULL, d <= thr);
                    while (mk) {
                        const int j = j0 + __ffs(mk) - 1;
                        mk &= mk - 1;
                        const float2 v = __ldg(reinterpret_cast<const float2*>(xb + (size_t)j * 64) + lane);
                        acc.x += v.x;
                        acc.y += v.y;
                    }
                }
            }
            m.x = __fdiv_rn(acc.x, divisor);                      // x1 = matmul(dpist, x) / float(k)
            m.y = __fdiv_rn(acc.y, divisor);
            const float2 xi = __ldg(reinterpret_cast<const float2*>(xb + (size_t)pos * 64) + lane);
            t.x = m.x - xi.x;                                     // t1 = x1 - x
            t.y = m.y - xi.y;
        }
        sM[pl * PB_LD + 2 * lane] = m.x;
        sM[pl * PB_LD + 2 * lane + 1] = m.y;
        sT[pl * PB_LD + 2 * lane] = t.x;
        sT[pl * PB_LD + 2 * lane + 1] = t.y;
    }
    __syncthreads();
    tile_dense64(sT, sWa, sBias, nullptr, sU, warp, lane);        // conv_a
    __syncthreads();
    tile_dense64(sU, sWb, sBias + 64, sM, sT, warp, lane);        // conv_b, then  + m
    __syncthreads();
    for (int i = tid; i < PB_TILE * 64; i += PB_THREADS) {
        const int pl = i >> 6, c = i & 63;
        if (tile0 + pl < N) {
            const size_t o = ((size_t)b * N + tile0 + pl) * ctot + coff + c;
            const float val = sT[pl * PB_LD + c];
            if (concat) concat[o] = val;
            if (concat16) concat16[o] = __float2bfloat16(val);      // operand of the bf16 tensor-core conv5
        }
    }
    if (HAS_NEXT) {
        tile_dense64(sT, sWn, sBias + 128, nullptr, sU, warp, lane);   // conv of the next block
        __syncthreads();
        for (int i = tid; i < PB_TILE * 64; i += PB_THREADS) {
            const int pl = i >> 6, c = i & 63;
            if (tile0 + pl < N) xnext[((size_t)b * N + tile0 + pl) * 64 + c] = sU[pl * PB_LD + c];
        }
    }
}

int proxy_block(const float* x, const KnnState& g, int B, int N, int arith, float divisor, const DenseDev& conv_a,
                const DenseDev& conv_b, const DenseDev* conv_next, float* concat, __nv_bfloat16* concat16, int ctot,
                int coff, float* xnext, cudaStream_t st) {
    EPC_CHECK_ARG(conv_a.cin == 64 && conv_a.cout == 64 && conv_b.cin == 64 && conv_b.cout == 64,
                  "ProxyConv block layers must be 64->64");
    if (B == 0) return EPC_OK;
    const size_t smem = (size_t)(3 * 4096 + 192 + 3 * PB_TILE * PB_LD) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        EPC_CUDA(cudaFuncSetAttribute(proxy_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        EPC_CUDA(cudaFuncSetAttribute(proxy_block_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    dim3 grid((N + PB_TILE - 1) / PB_TILE, B);
    if (conv_next) {
        proxy_block_kernel<true><<<grid, PB_THREADS, smem, st>>>(x, g.nbr, g.kthd, g.cnt, g.sorted, N, arith, divisor,
                                                                 conv_a.W, conv_a.b, conv_b.W, conv_b.b, conv_next->W,
                                                                 conv_next->b, concat, concat16, ctot, coff, xnext);
    } else {
        proxy_block_kernel<false><<<grid, PB_THREADS, smem, st>>>(x, g.nbr, g.kthd, g.cnt, g.sorted, N, arith, divisor,
                                                                  conv_a.W, conv_a.b, conv_b.W, conv_b.b, nullptr,
                                                                  nullptr, concat, concat16, ctot, coff, nullptr);
    }
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

}  // namespace epc
