// K2 (range-safe variant) -- ProxyConv backbone in fp32 storage / TF32 tensor-core operands, for the clouds that the
// fp16 fast path (backbone.cu) flagged because an activation left the fp16 range: e.g. the all-zero "fake" clouds of
// evaluate.py:425-430, whose thresholded neighbour sets hold all N points so that activations grow by N/20 per block.
// Every CTA first checks its cloud's flag and exits at once when it is clear, so this pass costs a few microseconds
// when nothing is flagged.
//
// models/epc-net.py:62-132, models/epc-net-l.py:44-80.
//
// The reference realises  m_i = (1/20) sum_{j in N(i)} x_j  as a dense (N x N mask) x (N x 64) batch
// matmul per block (2.15 GFLOP and a 64 MiB read each).  Here the neighbour lists of K1 are gathered
// directly (20 x 256 B rows per point, L2/L1 resident thanks to the Morton order), and the block body
//     t = m - x ; t = conv_a(t) ; t = conv_b(t) ; out = t + m ; x' = conv_{b+1}(out)
// runs on the tile while it is in shared memory.  Rows whose thresholded set has > 20 members (ties at
// the 20th distance, utils/tf_util.py:663-665) take an exact dense re-scan of the cloud.
#include "tc_gemm.cuh"
#include "kernels.h"

namespace epc {

// x0 = relu(BN(p W1 + b1)), cin = 3  (models/epc-net.py:66-69)
constexpr int SAFE_SLOTS = 4;     // grid.y of the range-safe kernels: CTA (x, y) serves the flagged clouds b = y, y + 4, ...

__global__ void conv_in_f32_kernel(const float4* __restrict__ sorted, int B, int N, const float* __restrict__ W,
                                   const float* __restrict__ bias, float* __restrict__ x, const int* __restrict__ flags) {
    __shared__ float sW[3 * 64 + 64];
    for (int i = threadIdx.x; i < 3 * 64 + 64; i += blockDim.x) sW[i] = (i < 192) ? W[i] : bias[i - 192];
    __syncthreads();
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int c4 = (threadIdx.x & 15) * 4;
    if (n >= N) return;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        if (flags[b] == 0) continue;
        const size_t r = (size_t)b * N + n;
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
}

int conv_in_f32(const float4* sorted, int B, int N, const DenseDev& L, float* x, const int* flags, cudaStream_t st) {
    EPC_CHECK_ARG(L.cin == 3 && L.cout == 64, "conv_in expects a 3->64 layer, got %d->%d", L.cin, L.cout);
    if (B == 0) return EPC_OK;
    dim3 grid((N * 16 + 255) / 256, B < SAFE_SLOTS ? B : SAFE_SLOTS);
    conv_in_f32_kernel<<<grid, 256, 0, st>>>(sorted, B, N, L.W, L.b, x, flags);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// ------------------------------------------------------------------------------------------------
// ProxyConv block on tcgen05 tensor cores (TF32).  One CTA = 128 consecutive (Morton-ordered) points:
//   gather  : all 8 warps; warp per point, lane = channel pair; 20 x 256 B row loads in flight per warp
//   GEMM a/b/n : [128 x 64] . [64 x 64] on the tensor cores, accumulator in TMEM; the A operand tile lives in
//               shared memory in the UMMA K-major 128B-swizzle layout and is rewritten in place by the epilogue
//               warps (t -> relu(conv_a) -> x_b), so activations never leave the SM between the three layers.
// Weights arrive as pre-swizzled 16 KB shared-memory images (prepared once at model creation).
// ------------------------------------------------------------------------------------------------
constexpr int PB_TILE = 128;
constexpr int PB_THREADS = 256;
constexpr int PB_MLD = 68;                       // row stride (floats) of the fp32 neighbour-mean tile (16 B aligned rows)
constexpr uint32_t PB_A_BYTES = 2 * 128 * 128;   // A tile: 2 k-blocks x 128 rows x 128 B
constexpr uint32_t PB_W_BYTES = 2 * 64 * 128;    // weight image: 2 k-blocks x 64 rows x 128 B
constexpr size_t PB_SMEM = 1024 + PB_A_BYTES + 2 * PB_W_BYTES + PB_TILE * PB_MLD * 4 + 3 * 64 * 4 + 64 +
                           PB_TILE * KNN_K * 2 + PB_TILE * 4;

// byte offset of the 16-byte chunk holding channels [4*k4, 4*k4+4) of row r in a [rows x 64] fp32 K-major SW128 tile
__device__ __forceinline__ uint32_t sw128_chunk(int r, int k4, uint32_t kblock_bytes) {
    return (uint32_t)(k4 >> 3) * kblock_bytes + (uint32_t)r * 128u + (uint32_t)(((k4 & 7) ^ (r & 7)) << 4);
}

__device__ __forceinline__ void pb_issue_gemm(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, uint64_t* bar) {
    constexpr uint32_t idesc = tc::make_idesc(2 /*TF32*/, 128, 64, 0, 0);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = tc::smem_desc_sw128(a_addr + kb * (PB_A_BYTES / 2) + kk * 32, 16, 1024);
            const uint64_t db = tc::smem_desc_sw128(w_addr + kb * (PB_W_BYTES / 2) + kk * 32, 16, 1024);
            tc::mma_ss<false>(tmem_d, da, db, idesc, (kb | kk) != 0);
        }
    tc::mma_commit(bar);
}

template <bool HAS_NEXT>
__global__ void __launch_bounds__(PB_THREADS, 2)
proxy_block_f32_kernel(const int* __restrict__ flags, int B, const float* __restrict__ x, const uint16_t* __restrict__ nbr, const float* __restrict__ kthd,
                   const int* __restrict__ cnt, const float4* __restrict__ sorted, int N, int arith, float divisor,
                   const float* __restrict__ Wa_img, const float* __restrict__ ba, const float* __restrict__ Wb_img,
                   const float* __restrict__ bb, const float* __restrict__ Wn_img, const float* __restrict__ bn,
                   float* __restrict__ concat, __nv_bfloat16* __restrict__ concat16, int ctot, int coff,
                   float* __restrict__ xnext, float* __restrict__ cloud_absmax) {
    {                                                    // only the clouds the fp16 pass flagged: usually none
        bool any = false;
        for (int bb2 = blockIdx.y; bb2 < B; bb2 += gridDim.y) any |= (flags[bb2] != 0);
        if (!any) return;
    }
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base;                                  // A operand tile (t, then relu(conv_a), then x_b)
    uint8_t* sW0 = sA + PB_A_BYTES;                      // conv_a weights, later conv_next
    uint8_t* sW1 = sW0 + PB_W_BYTES;                     // conv_b weights
    float* sM = reinterpret_cast<float*>(sW1 + PB_W_BYTES);   // neighbour mean m, [128][66] fp32
    float* sBias = sM + PB_TILE * PB_MLD;                // [3][64]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sBias + 192);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    int* sCnt = reinterpret_cast<int*>(bar + 8);                         // [128] size of each point's thresholded set
    unsigned short* sNbr = reinterpret_cast<unsigned short*>(sCnt + PB_TILE);   // [128][20] neighbour positions

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile0 = blockIdx.x * PB_TILE;
    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(tmem_slot, 64);
        tc::tmem_relinquish();
    }
    uint32_t ph = 0;                                     // phase of the MMA-completion barrier, flips at every wait
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    if (flags[b] == 0) continue;

    // ---- weights (conv_a's slot is overwritten with the next block's first conv further down) ----------------------
    {
        const uint4* ga = reinterpret_cast<const uint4*>(Wa_img);
        const uint4* gb = reinterpret_cast<const uint4*>(Wb_img);
        uint4* s0 = reinterpret_cast<uint4*>(sW0);
        uint4* s1 = reinterpret_cast<uint4*>(sW1);
        for (int i = tid; i < (int)(PB_W_BYTES / 16); i += PB_THREADS) {
            s0[i] = __ldg(ga + i);
            s1[i] = __ldg(gb + i);
        }
        if (tid < 64) {
            sBias[tid] = ba[tid];
            sBias[64 + tid] = bb[tid];
            sBias[128 + tid] = HAS_NEXT ? bn[tid] : 0.f;
        }
    }

    // ---- gather-mean ---------------------------------------------------------------------------------------
    // warp per point; the two half-warps fetch two different neighbour rows per instruction (lane & 15 = which
    // float4 of the 256-byte row), so a point costs 10 LDG.128 instead of 20 LDG.64.
    const float* xb = x + (size_t)b * N * 64;
    const size_t row_tile = (size_t)b * N + tile0;
    for (int i = tid; i < PB_TILE * KNN_K; i += PB_THREADS) sNbr[i] = nbr[row_tile * KNN_K + i];
    if (tid < PB_TILE) sCnt[tid] = cnt[row_tile + tid] & 0xffffff;     // the top byte holds the fast path's out-of-tile count
    __syncthreads();
    const int half = lane >> 4, c4 = lane & 15;
    const float inv_div = 1.0f / divisor;               // x1 = matmul(dpist, x) / float(k): one rounding differs from a true division
#pragma unroll 1
    for (int pl = warp; pl < PB_TILE; pl += PB_THREADS / 32) {
        const int pos = tile0 + pl;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sCnt[pl] == KNN_K) {
            float4 v[KNN_K / 2];
#pragma unroll
            for (int q = 0; q < KNN_K / 2; ++q) {
                const int j = sNbr[pl * KNN_K + 2 * q + half];
                v[q] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)j * 64) + c4);
            }
#pragma unroll
            for (int q = 0; q < KNN_K / 2; ++q) {
                acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w;
            }
        } else {
            // ties at the 20th distance: the set is {j : d_ij <= kthd_i}; re-scan the cloud exactly (rare)
            const size_t row = row_tile + pl;
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
                    if (half == 0) {                     // one half-warp accumulates; the other contributes zeros
                        const float4 vv = __ldg(reinterpret_cast<const float4*>(xb + (size_t)j * 64) + c4);
                        acc.x += vv.x; acc.y += vv.y; acc.z += vv.z; acc.w += vv.w;
                    }
                }
            }
        }
        acc.x += __shfl_xor_sync(FULL, acc.x, 16);
        acc.y += __shfl_xor_sync(FULL, acc.y, 16);
        acc.z += __shfl_xor_sync(FULL, acc.z, 16);
        acc.w += __shfl_xor_sync(FULL, acc.w, 16);
        const float4 m = make_float4(acc.x * inv_div, acc.y * inv_div, acc.z * inv_div, acc.w * inv_div);
        if (half == 0) {
            *reinterpret_cast<float4*>(sM + pl * PB_MLD + 4 * c4) = m;
        } else {
            const float4 xi = __ldg(reinterpret_cast<const float4*>(xb + (size_t)pos * 64) + c4);
            *reinterpret_cast<float4*>(sA + sw128_chunk(pl, c4, PB_A_BYTES / 2)) =          // t1 = x1 - x (TF32 MMA operand)
                make_float4(round_tf32(m.x - xi.x), round_tf32(m.y - xi.y), round_tf32(m.z - xi.z), round_tf32(m.w - xi.w));
        }
    }
    tc::fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t a_addr = tc::smem_u32(sA), w0_addr = tc::smem_u32(sW0), w1_addr = tc::smem_u32(sW1);
    const bool epi = warp >= 4;                         // warps 4..7 own TMEM lane quarters 0..3
    const int erow = (warp & 3) * 32 + lane;            // this epilogue thread's point within the tile
    const uint32_t trow = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    const size_t grow = (size_t)b * N + tile0 + erow;

    // ---- conv_a -----------------------------------------------------------------------------------------------
    if (tid == 0) {
        pb_issue_gemm(tmem_d, a_addr, w0_addr, bar);
        tc::mbar_wait(bar, ph);         // one thread polls; everybody else sleeps on the hardware barrier
    }
    ph ^= 1u;
    __syncthreads();
    tc::tc_fence_after();
    if (epi) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float v[32];
            tc::tmem_ld32(trow + 32u * h, v);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const int k = 32 * h + i;
                float4 o;
                o.x = round_tf32(fmaxf(v[i] + sBias[k], 0.f));
                o.y = round_tf32(fmaxf(v[i + 1] + sBias[k + 1], 0.f));
                o.z = round_tf32(fmaxf(v[i + 2] + sBias[k + 2], 0.f));
                o.w = round_tf32(fmaxf(v[i + 3] + sBias[k + 3], 0.f));
                *reinterpret_cast<float4*>(sA + sw128_chunk(erow, k >> 2, PB_A_BYTES / 2)) = o;
            }
        }
    } else if (HAS_NEXT) {
        // conv_a's weights are dead now: bring in the next block's first conv while the epilogue runs
        const uint4* gn = reinterpret_cast<const uint4*>(Wn_img);
        uint4* s0 = reinterpret_cast<uint4*>(sW0);
        for (int i = tid; i < (int)(PB_W_BYTES / 16); i += 128) s0[i] = __ldg(gn + i);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    // ---- conv_b, residual, concat ---------------------------------------------------------------------------------
    if (tid == 0) {
        pb_issue_gemm(tmem_d, a_addr, w1_addr, bar);
        tc::mbar_wait(bar, ph);
    }
    ph ^= 1u;
    __syncthreads();
    tc::tc_fence_after();
    if (epi) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float v[32];
            tc::tmem_ld32(trow + 32u * h, v);
            float o[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const int k = 32 * h + i;
                const float4 mm = *reinterpret_cast<const float4*>(sM + erow * PB_MLD + k);
                o[i] = fmaxf(v[i] + sBias[64 + k], 0.f) + mm.x;                        // x_b = relu(conv_b) + m
                o[i + 1] = fmaxf(v[i + 1] + sBias[65 + k], 0.f) + mm.y;
                o[i + 2] = fmaxf(v[i + 2] + sBias[66 + k], 0.f) + mm.z;
                o[i + 3] = fmaxf(v[i + 3] + sBias[67 + k], 0.f) + mm.w;
            }
            if (HAS_NEXT) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(sA + sw128_chunk(erow, (32 * h + i) >> 2, PB_A_BYTES / 2)) =
                        make_float4(round_tf32(o[i]), round_tf32(o[i + 1]), round_tf32(o[i + 2]), round_tf32(o[i + 3]));
            }
            if (concat) {
                float4* dst = reinterpret_cast<float4*>(concat + grow * ctot + coff + 32 * h);
#pragma unroll
                for (int i = 0; i < 8; ++i)       // operand of the TF32 conv5 (EPC-Net-L, KD export): store it rounded
                    dst[i] = make_float4(round_tf32(o[4 * i]), round_tf32(o[4 * i + 1]), round_tf32(o[4 * i + 2]), round_tf32(o[4 * i + 3]));
            }
            if (concat16) {
                if (cloud_absmax) {                       // the fp8 head's range bound must cover the re-computed rows too
                    float mx = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fabsf(o[i]));
                    mx = fminf(mx * 1.01f, 3.0e38f);      // bf16 rounding may round up
                    if (mx > 0.f) atomicMax(reinterpret_cast<unsigned*>(cloud_absmax) + (int)(grow / N), __float_as_uint(mx));
                }
                uint4* dst = reinterpret_cast<uint4*>(concat16 + grow * ctot + coff + 32 * h);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __nv_bfloat162 hh = __floats2bfloat162_rn(o[8 * i + 2 * j], o[8 * i + 2 * j + 1]);
                        pk[j] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    dst[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
        }
    }
    if (HAS_NEXT) {
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        // ---- first conv of the next block ----------------------------------------------------------------------
        if (tid == 0) {
            pb_issue_gemm(tmem_d, a_addr, w0_addr, bar);
            tc::mbar_wait(bar, ph);
        }
        ph ^= 1u;
        __syncthreads();
        tc::tc_fence_after();
        if (epi) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[32];
                tc::tmem_ld32(trow + 32u * h, v);
                float4* dst = reinterpret_cast<float4*>(xnext + grow * 64 + 32 * h);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = 32 * h + 4 * i;
                    dst[i] = make_float4(fmaxf(v[4 * i] + sBias[128 + k], 0.f), fmaxf(v[4 * i + 1] + sBias[129 + k], 0.f),
                                         fmaxf(v[4 * i + 2] + sBias[130 + k], 0.f), fmaxf(v[4 * i + 3] + sBias[131 + k], 0.f));
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();                                     // the tile's shared memory and TMEM are free again
  }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(*tmem_slot, 64);
}

int proxy_block_f32(const int* flags, const float* x, const KnnState& g, int B, int N, int arith, float divisor, const DenseDev& conv_a,
                const DenseDev& conv_b, const DenseDev* conv_next, float* concat, __nv_bfloat16* concat16, int ctot,
                int coff, float* xnext, float* cloud_absmax, cudaStream_t st) {
    EPC_CHECK_ARG(conv_a.cin == 64 && conv_a.cout == 64 && conv_b.cin == 64 && conv_b.cout == 64,
                  "ProxyConv block layers must be 64->64");
    EPC_CHECK_ARG(N % PB_TILE == 0, "proxy_block: N=%d must be a multiple of %d", N, PB_TILE);
    EPC_CHECK_ARG(conv_a.Wimg32 && conv_b.Wimg32 && (!conv_next || conv_next->Wimg32), "proxy_block: missing swizzled weight images");
    if (B == 0) return EPC_OK;
    static PerDeviceSize attr_a, attr_b;
    EPC_CUDA(ensure_dyn_smem(proxy_block_f32_kernel<true>, PB_SMEM, attr_a));
    EPC_CUDA(ensure_dyn_smem(proxy_block_f32_kernel<false>, PB_SMEM, attr_b));
    dim3 grid(N / PB_TILE, B < SAFE_SLOTS ? B : SAFE_SLOTS);
    if (conv_next) {
        proxy_block_f32_kernel<true><<<grid, PB_THREADS, PB_SMEM, st>>>(flags, B, x, g.nbr, g.kthd, g.cnt, g.sorted, N, arith, divisor,
                                                                    conv_a.Wimg32, conv_a.b, conv_b.Wimg32, conv_b.b,
                                                                    conv_next->Wimg32, conv_next->b, concat, concat16, ctot,
                                                                    coff, xnext, cloud_absmax);
    } else {
        proxy_block_f32_kernel<false><<<grid, PB_THREADS, PB_SMEM, st>>>(flags, B, x, g.nbr, g.kthd, g.cnt, g.sorted, N, arith, divisor,
                                                                     conv_a.Wimg32, conv_a.b, conv_b.Wimg32, conv_b.b, nullptr,
                                                                     nullptr, concat, concat16, ctot, coff, nullptr, cloud_absmax);
    }
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// Host: [cin=64][cout=64] folded weights -> the shared-memory image of the K-major, 128B-swizzled B operand
// (element (n,k) = W[k][n]): 2 k-blocks x 64 rows x 128 B.
void make_w64_image_f32(const float* W, float* img) {
    for (int k = 0; k < 64; ++k)
        for (int n = 0; n < 64; ++n) {
            const int kb = k >> 5, chunk = (k & 31) >> 2, within = k & 3;
            const size_t off_bytes = (size_t)kb * 8192 + (size_t)n * 128 + (size_t)((chunk ^ (n & 7)) << 4) + within * 4;
            img[off_bytes / 4] = W[k * 64 + n];
        }
}

}  // namespace epc
