// The G_VLAD / NetVLAD head on an fp8 (e4m3) copy of the per-point features: conv5 -> H' (fp8), cluster assignment + VLAD
// accumulate in one launch (the two-role design of head_fused.cu) on tcgen05 kind::f8f6f4.
//
// Why fp8: the three head GEMMs are bound by moving H -- 8 MiB per cloud as bf16, written once (the HBM write stream bounds
// conv5) and read twice (bytes in flight per SM bound the assignment and VLAD).  As e4m3 it is 4 MiB; Wc^T shrinks to 64 KB of
// shared memory, which leaves the assignment a 160 KB ring instead of 96 KB, and a VLAD stage holds 128 points instead of 64.
// Why it is safe:
//   range     every stored tensor carries an exact POWER-OF-TWO scale that cancels downstream:
//             H'  = 2^e H with e per cloud from the bound |H[r,f]| <= absmax(x) * max_f sum_c |W5[c,f]| + max|b5| <= 2^E,
//                   e = 8 - E (|H'| <= 256 < 448: never saturates).  The row norm is taken from the scaled fp32 accumulators,
//                   and both GEMMs only see H'/|H'| = H/|H| (models/epc-net.py:147-148), so 2^e drops out exactly;
//             Wc' = 2^w Wc per model, 2^-w folded into the cluster-BN scale;
//             S'' = 2^t softmax/|H'| with t per cloud from the smallest row norm; V is multiplied by 2^-t when finalised.
//   precision e4m3 keeps 3 mantissa bits per element, but every output is a sum over 1024 features or 4096 points, and the
//             conversions round STOCHASTICALLY (cvt.rs with hashed bits, tc_gemm.cuh) so that the errors average out even
//             over identical points: measured
//             on the descriptors (tests/test_gpu_model.py, bench.py parity block) the head stays ~4x..10x inside the
//             north_star tolerance (max-abs <= 1e-3 after L2, cosine >= 0.9999).  EPC_HEAD_FP8=0 selects the bf16 head.
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <stdlib.h>

#include "tc_gemm.cuh"
#include "kernels.h"

namespace epc {
namespace h8 {

using namespace tc;

constexpr uint32_t TILE_BYTES = 128 * 128;              // 128 rows x 128 one-byte elements: every H' / S'' box
constexpr uint32_t WC_BYTES = 64 * 128;                 // one 128-feature k-block of Wc'^T [64 x 1024]
constexpr int WC_KB = 1024 / 128;                       // 8 k-blocks: 64 KB resident
constexpr int A_STAGES = 10;                            // assignment ring: 160 KB
constexpr uint32_t V_STAGE_BYTES = 2 * TILE_BYTES;      // VLAD stage: H' [128 points x 128 features] + S'' [128 points x 128 B]
constexpr int V_STAGES = 6;                             // 192 KB
constexpr size_t DATA_BYTES = (size_t)WC_KB * WC_BYTES + (size_t)A_STAGES * TILE_BYTES;       // 224 KB
static_assert((size_t)V_STAGES * V_STAGE_BYTES <= DATA_BYTES, "VLAD ring must fit the shared allocation");
constexpr int MAX_STAGES = A_STAGES > V_STAGES ? A_STAGES : V_STAGES;
constexpr size_t SMEM_BYTES = 1024 + DATA_BYTES + 1024 /*column-sum scratch*/ + 8 * (2 * MAX_STAGES + 5) + 64;

struct Params {
    GemmParams pa;              // assignment: M = rows of the sub-batch, N = 64, K = 1024 (EPI_ASSIGN_FP8 fields)
    GemmParams pv;              // VLAD: M = 1024, N = 64 stored (128 computed), K = points per split slab (EPI_STORE_F32)
    int n_assign;               // CTAs [0, n_assign) run the assignment, the rest VLAD
    int clouds, tiles_per_cloud;
    int reverse;
    int* ready;                 // [clouds] row tiles of the cloud whose S'' is in memory (zeroed before the launch)
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// D[tmem] (+)= A[smem desc] * B[smem desc], e4m3 operands (32 elements of K per instruction); issued by ONE thread
__device__ __forceinline__ void mma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}

__global__ void __launch_bounds__(192, 1)
assign_vlad_fp8_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmWc,
                       const __grid_constant__ CUtensorMap tmS, const Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* scratch = reinterpret_cast<float*>(base + DATA_BYTES);                   // [4][64] (assignment)
    uint64_t* full = reinterpret_cast<uint64_t*>(base + DATA_BYTES + 1024);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* b_full = empty + MAX_STAGES;
    uint64_t* tfull = b_full + 1;          // [2] accumulator ready
    uint64_t* tempty = tfull + 2;          // [2] accumulator drained (4 arrivals: one per epilogue warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool assign_role = (int)blockIdx.x < P.n_assign;
    constexpr uint32_t TMEM_COLS = 256;                  // assignment: 2 x 64 columns; VLAD: 2 x 128

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmH);
        tma_prefetch_desc(assign_role ? &tmWc : &tmS);
        for (int s = 0; s < MAX_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(b_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (assign_role) {
        // ---------------- assignment: Wc'^T resident, H' tiles stream through a 10-deep ring -----------------------------------
        const GemmParams& p = P.pa;
        uint8_t* sB = base;                                              // [8][8 KB]  Wc'^T, resident
        uint8_t* sA = sB + (size_t)WC_KB * WC_BYTES;                     // [A_STAGES][16 KB]
        const int num_m_tiles = (p.M + TC_BM - 1) / TC_BM;
        const int cta = blockIdx.x, stride = P.n_assign;
        auto tile_of = [&](int mt) { return P.reverse ? num_m_tiles - 1 - mt : mt; };
        if (warp == 0) {
            if (lane == 0) {
                mbar_expect_tx(b_full, (uint32_t)WC_KB * WC_BYTES);
                for (int kb = 0; kb < WC_KB; ++kb) tma_load_2d(sB + (size_t)kb * WC_BYTES, &tmWc, b_full, kb * 128, 0);
                int s = 0;
                uint32_t ph = 0;
                for (int mt = cta; mt < num_m_tiles; mt += stride) {
                    for (int kb = 0; kb < WC_KB; ++kb) {
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], TILE_BYTES);
                        tma_load_2d(sA + (size_t)s * TILE_BYTES, &tmH, &full[s], kb * 128, tile_of(mt) * TC_BM);
                        if (++s == A_STAGES) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                constexpr uint32_t idesc = make_idesc(0 /*e4m3*/, TC_BM, 64, 0, 0);
                mbar_wait(b_full, 0);
                int s = 0, tile = 0;
                uint32_t ph = 0;
                for (int mt = cta; mt < num_m_tiles; mt += stride, ++tile) {
                    const int buf = tile & 1;
                    mbar_wait(&tempty[buf], ((tile >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 64);
                    for (int kb = 0; kb < WC_KB; ++kb) {
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(sA + (size_t)s * TILE_BYTES);
                        const uint32_t b_addr = smem_u32(sB + (size_t)kb * WC_BYTES);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)                       // 32 one-byte elements of K = 32 B along the swizzled row
                            mma_f8(tmem_d, smem_desc_sw128(a_addr + kk * 32, 16, 1024), smem_desc_sw128(b_addr + kk * 32, 16, 1024),
                                   idesc, (kb | kk) != 0);
                        mma_commit(&empty[s]);
                        if (++s == A_STAGES) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                    mma_commit(&tfull[buf]);
                }
            }
        } else {
            const int q = warp & 3;
            const int row = q * 32 + lane;
            int tile = 0;
            for (int mt = cta; mt < num_m_tiles; mt += stride, ++tile) {
                const int buf = tile & 1;
                mbar_wait(&tfull[buf], (tile >> 1) & 1);
                tc_fence_after();
                EpiCtx c;
                c.trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 64);
                c.m0 = tile_of(mt) * TC_BM; c.m = c.m0 + row; c.row = row; c.lane = lane; c.n0 = 0; c.mtile = tile_of(mt);
                c.c_off = 0; c.scratch = scratch; c.epi_tid = threadIdx.x - 64; c.bias = nullptr;
                c.col_begin = 0; c.col_end = 64; c.nparts = 1; c.npart = 0; c.warp_slot = warp - 2;
                epilogue_tile<64, EPI_ASSIGN_FP8>(p, c);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]);
                __threadfence();                                           // publish (see head_fused.cu)
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (c.epi_tid == 0) {
                    __threadfence();
                    atomicAdd(P.ready + tile_of(mt) / P.tiles_per_cloud, 1);
                }
            }
        }
    } else {
        // ---------------- VLAD: V[cloud][128 features, 64] (+)= H'^T S'' over one split-K slab per work item; both operands
        // MN-major straight from the row-major buffers, 128 points per stage ----------------------------------------------------
        const GemmParams& p = P.pv;
        const int cta = blockIdx.x - P.n_assign, stride = gridDim.x - P.n_assign;
        const int items_per_cloud = (p.M / TC_BM) * p.splitk;
        const int n_items = P.clouds * items_per_cloud;
        const int nkb = p.K / 128;
        auto decode = [&](int it, int& cloud, int& m_tile, int& split) {
            const int cs = it / items_per_cloud, rem = it - cs * items_per_cloud;
            cloud = P.reverse ? P.clouds - 1 - cs : cs;
            split = rem / (p.M / TC_BM);
            m_tile = rem - split * (p.M / TC_BM);
        };
        if (warp == 0) {
            if (lane == 0) {
                int s = 0;
                uint32_t ph = 0;
                for (int it = cta; it < n_items; it += stride) {
                    int cloud, m_tile, split;
                    decode(it, cloud, m_tile, split);
                    {
                        long long spins = 0;
                        while (ld_acquire(P.ready + cloud) < P.tiles_per_cloud) {
                            __nanosleep(64);
                            if (++spins > (1ll << 26)) __trap();          // several seconds: the assignment role is not running
                        }
                        asm volatile("fence.proxy.async;" ::: "memory");
                    }
                    const int krow0 = cloud * p.k_batch_rows + split * p.K;
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], V_STAGE_BYTES);
                        uint8_t* a = base + (size_t)s * V_STAGE_BYTES;
                        const int k0 = krow0 + kb * 128;
                        tma_load_2d(a, &tmH, &full[s], m_tile * TC_BM, k0);              // box {128 features, 128 points}
                        tma_load_2d(a + TILE_BYTES, &tmS, &full[s], 0, k0);             // box {128 B (64 clusters + padding), 128 points}
                        if (++s == V_STAGES) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                constexpr uint32_t idesc = make_idesc(0 /*e4m3*/, TC_BM, 128, 1, 1);      // both operands MN-major, N = 128
                constexpr uint32_t step = 32 * 128;                                     // 32 points (k rows) per instruction
                int s = 0, tile = 0;
                uint32_t ph = 0;
                for (int it = cta; it < n_items; it += stride, ++tile) {
                    const int buf = tile & 1;
                    mbar_wait(&tempty[buf], ((tile >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 128);
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(base + (size_t)s * V_STAGE_BYTES);
                        const uint32_t b_addr = a_addr + TILE_BYTES;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            mma_f8(tmem_d, smem_desc_sw128(a_addr + kk * step, TILE_BYTES, 1024), smem_desc_sw128(b_addr + kk * step, TILE_BYTES, 1024),
                                   idesc, (kb | kk) != 0);
                        mma_commit(&empty[s]);
                        if (++s == V_STAGES) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                    mma_commit(&tfull[buf]);
                }
            }
        } else {
            const int q = warp & 3;
            const int row = q * 32 + lane;
            int tile = 0;
            for (int it = cta; it < n_items; it += stride, ++tile) {
                int cloud, m_tile, split;
                decode(it, cloud, m_tile, split);
                const int buf = tile & 1;
                mbar_wait(&tfull[buf], (tile >> 1) & 1);
                tc_fence_after();
                EpiCtx c;
                c.trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128);
                c.m0 = m_tile * TC_BM; c.m = c.m0 + row; c.row = row; c.lane = lane; c.n0 = 0; c.mtile = m_tile;
                c.c_off = (long long)cloud * p.c_batch + (long long)split * p.c_slab;
                c.scratch = scratch; c.epi_tid = threadIdx.x - 64; c.bias = nullptr;
                c.col_begin = 0; c.col_end = 64; c.nparts = 1; c.npart = 0; c.warp_slot = warp - 2;     // columns 64..127 = padding
                epilogue_tile<64, EPI_STORE_F32>(p, c);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// per cloud: t = 2^(8 - E) with 1 / min_r |H'_r| <= 2^E  (rows with a vanishing norm are skipped: their H' is ~0 and their
// S'' saturates harmlessly), and 1 / t for the finalise.  rowss [R, parts] partial sums of squares.
__global__ void __launch_bounds__(256) sprime_scale_kernel(const float* __restrict__ rowss, int parts, int N, float* __restrict__ t,
                                                          float* __restrict__ t_inv) {
    __shared__ float s[8];
    const int b = blockIdx.x;
    float mn = INFINITY;
    for (int r = threadIdx.x; r < N; r += blockDim.x) {
        float ss = 0.f;
        for (int i = 0; i < parts; ++i) ss += rowss[((size_t)b * N + r) * parts + i];
        if (ss > 1e-30f) mn = fminf(mn, ss);
    }
    mn = warp_min(mn);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mn = fminf(mn, s[w]);
        float tt = 1.f;
        if (mn < INFINITY) {
            int ex;
            frexpf(1.0f / sqrtf(mn) * 1.0001f, &ex);              // max S' = softmax / |H'| <= 1 / min |H'| <= 2^ex
            ex = max(-100, min(100, ex));
            tt = ldexpf(1.0f, 8 - ex);
        }
        t[b] = tt;
        t_inv[b] = 1.0f / tt;
    }
}

// stand-alone loupe API (caller-provided fp32 rows, used as given): X [R, F] -> fp8 with a per-ROW power-of-two scale s_r
// (|s_r x| <= 256) and rowss[r] := s_r^2, so that the assignment's row factor 1/sqrt(rowss) undoes the scale exactly
__global__ void __launch_bounds__(256) f32_to_fp8_rows_kernel(const float* __restrict__ X, long long R, int F, int rows_per_cloud,
                                                             uint8_t* __restrict__ Y, float* __restrict__ rowss) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float4* x = reinterpret_cast<const float4*>(X + (size_t)r * F);
    float m = 0.f;
    for (int i = lane; i < F / 4; i += 32) {
        const float4 v = __ldg(x + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    m = warp_max(m);
    int ex;
    frexpf(m + 1e-30f, &ex);
    ex = max(-100, min(100, ex));
    const float sc = ldexpf(1.0f, 8 - ex);
    uint32_t* y = reinterpret_cast<uint32_t*>(Y + (size_t)r * F);
    for (int i = lane; i < F / 4; i += 32) {
        const float4 v = __ldg(x + i);
        y[i] = tc::f32x4_to_e4m3_sr(v.x * sc, v.y * sc, v.z * sc, v.w * sc, tc::hash_bits((uint32_t)(r % rows_per_cloud), (uint32_t)(4 * i)));
    }
    if (lane == 0) rowss[r] = sc * sc;
}

}  // namespace h8

// conv5 (models/epc-net.py:136-139) with the fp8 output format described at the head of this file: H8 [R, 1024] e4m3 bytes,
// rowss [R, 8] partial sums of squares of the scaled fp32 values
int tc_conv5_fp8(const __nv_bfloat16* Xc, long long R, int cin, int rows_per_cloud, const __nv_bfloat16* W5t, const float* b5,
                 const float* b5_host, const float* cloud_absmax_dev, float l1max, float bmax, uint8_t* H8, float* rowss, cudaStream_t st) {
    tc::EpiExtra ex;        // by-value kernel argument: output tensor map (one 32-row x 128-byte box per epilogue warp) + the bias
    if (int rc = make_tmap_2d(&ex.tmC, H8, (uint64_t)R, 1024, 1024, 128, 32)) return rc;
    memcpy(ex.bias, b5_host, sizeof(ex.bias));
    tc::GemmParams p = {};
    p.M = (int)R; p.N = 1024; p.K = cin; p.splitk = 1; p.C = H8; p.ldc = 1024; p.bias = b5; p.relu = 1; p.aux = rowss;
    p.cloud_absmax = cloud_absmax_dev; p.l1max = l1max; p.bmax = bmax; p.rows_per_cloud = rows_per_cloud;
    { static const int pf = getenv("EPC_CONV5_PREFETCH") ? atoi(getenv("EPC_CONV5_PREFETCH")) : 0; p.l2_prefetch_tiles = pf; }
    Operand<__nv_bfloat16> a{Xc, R, cin, cin}, b{W5t, 1024, cin, cin};
    static const bool tl = getenv("EPC_BRES_TIMELINE") != nullptr;
    if (tl) {       // debug: per-tile clock64 stamps of CTA 0 (MMA start / issued, epilogue start / end), printed for the first calls
        static long long* dev = nullptr;
        static int calls = 0;
        if (!dev) EPC_CUDA(cudaMalloc(&dev, sizeof(long long) * 256));
        EPC_CUDA(cudaMemsetAsync(dev, 0, sizeof(long long) * 256, st));
        p.timeline = dev;
        const int rc = tc_gemm_bres_launch<__nv_bfloat16, 256, tc::EPI_CONV5_FP8, 8>(a, b, p, st, &ex);
        if (rc == EPC_OK && calls++ < 2) {
            long long h[256];
            EPC_CUDA(cudaStreamSynchronize(st));
            EPC_CUDA(cudaMemcpy(h, dev, sizeof(h), cudaMemcpyDeviceToHost));
            for (int t = 0; t < 64 && h[4 * t + 2]; ++t)
                fprintf(stderr, "conv5 tile %2d: mma start %7lld issued %7lld | epi start %7lld end %7lld | epi %5lld  period %5lld\n", t, h[4 * t] - h[0],
                        h[4 * t + 1] - h[0], h[4 * t + 2] - h[0], h[4 * t + 3] - h[0], h[4 * t + 3] - h[4 * t + 2], t ? h[4 * t + 2] - h[4 * t - 2] : 0ll);
        }
        return rc;
    }
    return tc_gemm_bres_launch<__nv_bfloat16, 256, tc::EPI_CONV5_FP8, 8>(a, b, p, st, &ex);
}

int sprime_scale(const float* rowss, int parts, int clouds, int N, float* t, float* t_inv, cudaStream_t st) {
    if (clouds == 0) return EPC_OK;
    h8::sprime_scale_kernel<<<clouds, 256, 0, st>>>(rowss, parts, N, t, t_inv);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

int f32_to_fp8_rows(const float* X, long long R, int F, int rows_per_cloud, uint8_t* Y, float* rowss, cudaStream_t st) {
    EPC_CHECK_ARG(F % 4 == 0, "f32_to_fp8_rows: F=%d must be a multiple of 4", F);
    if (R == 0) return EPC_OK;
    h8::f32_to_fp8_rows_kernel<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(X, R, F, rows_per_cloud, Y, rowss);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// Soft assignment of the rows of H8 [clouds * N, 1024] (e4m3) and the VLAD sums of every cloud, one launch.  Wct8: Wc'^T
// [64, 1024] e4m3; bn_scale already carries 2^-w; sscale [clouds] = 2^t.  Outputs: S8 [R, 128 B] (64 e4m3 values + padding),
// a_part [R/128, 64], V fp32 slabs [splitk][.., 1024, 64] scaled by 2^t (the finalise multiplies by 2^-t).
int tc_assign_vlad_fp8(const uint8_t* H8, int clouds, int N, const uint8_t* Wct8, const float* rowss, int parts, const float* bn_scale,
                       const float* bn_shift, const float* sscale, uint8_t* S8, float* a_part, float* V, int splitk, long long slab,
                       int* ready, cudaStream_t st) {
    const long long R = (long long)clouds * N;
    EPC_CHECK_ARG(clouds >= 1 && N % 128 == 0 && (N / splitk) % 128 == 0, "tc_assign_vlad_fp8: bad shape clouds=%d N=%d splitk=%d", clouds, N, splitk);
    h8::Params P = {};
    P.pa.M = (int)R; P.pa.N = 64; P.pa.K = 1024; P.pa.splitk = 1; P.pa.C = S8; P.pa.ldc = 128; P.pa.aux = a_part; P.pa.rowss = rowss;
    P.pa.rowss_parts = parts; P.pa.bn_scale = bn_scale; P.pa.bn_shift = bn_shift; P.pa.sscale = sscale; P.pa.rows_per_cloud = N;
    P.pv.M = 1024; P.pv.N = 64; P.pv.K = N / splitk; P.pv.k_batch_rows = N; P.pv.splitk = splitk; P.pv.C = V; P.pv.ldc = 64;
    P.pv.c_batch = 1024ll * 64; P.pv.c_slab = slab;
    P.clouds = clouds; P.tiles_per_cloud = N / 128; P.ready = ready;
    P.reverse = getenv("EPC_ASSIGN_FORWARD") ? 0 : 1;
    const int sms = sm_count();
    static const int env_assign = getenv("EPC_HEAD_ASSIGN_CTAS") ? atoi(getenv("EPC_HEAD_ASSIGN_CTAS")) : 0;
    int n_assign = env_assign > 0 ? env_assign : (sms * 105 + 74) / 148;      // measured best split on B200: 105 / 43
    if (n_assign < 1) n_assign = 1;
    if (n_assign > sms - 1) n_assign = sms - 1;
    P.n_assign = n_assign;
    CUtensorMap tmH, tmWc, tmS;
    if (int rc = make_tmap_2d(&tmH, H8, (uint64_t)R, 1024, 1024, 128, 128)) return rc;
    if (int rc = make_tmap_2d(&tmWc, Wct8, 64, 1024, 1024, 128, 64)) return rc;
    if (int rc = make_tmap_2d(&tmS, S8, (uint64_t)R, 128, 128, 128, 128)) return rc;
    EPC_CUDA(cudaMemsetAsync(ready, 0, sizeof(int) * (size_t)clouds, st));
    static PerDeviceSize attr;
    EPC_CUDA(ensure_dyn_smem(h8::assign_vlad_fp8_kernel, h8::SMEM_BYTES, attr));
    h8::assign_vlad_fp8_kernel<<<sms, 192, h8::SMEM_BYTES, st>>>(tmH, tmWc, tmS, P);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

}  // namespace epc
