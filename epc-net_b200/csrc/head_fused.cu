// Cluster assignment + VLAD accumulate (loupe.py:255-291) in ONE launch with two CTA roles.
//
// The two contractions need different traversals of the bf16 per-point features H (8 MiB per cloud, written by conv5) -- the
// assignment reduces over the 1024 features of a point (S' = softmax(BN(Hn Wc))/|H|), VLAD over the 4096 points of a cloud
// (V = H^T S') and only after the cloud's S' is complete -- and a single CTA cannot hold both accumulators (DESIGN.md
// section 7), so they stay two GEMMs.  Both are bound by how many bytes one SM can keep in flight (shared memory is the
// ring: 96 KB next to the resident Wc^T, 192 KB for VLAD) against a ~2.5 us loaded memory latency, not by the tensor pipe,
// so running them side by side -- the first `n_assign` CTAs the assignment over the row tiles in cloud order, the others
// the VLAD work items (cloud, 128-feature tile, split-K slab) in the same order, each waiting (acquire) for its cloud's tile
// counter -- overlaps two latency-bound streams and removes a launch boundary: 3.52 -> 3.06 us/cloud on B200.  The grid is
// one CTA per SM, all co-resident; CTAs are dispatched in block-id order, so an assignment CTA (low id) is never queued
// behind a waiting VLAD CTA.
// What it does NOT achieve (measured): serving VLAD's re-read of H from the L2.  A tile takes ~7 us to
// stream through an assignment CTA's ring, so a cloud's lines are first touched up to ~20 us before VLAD needs them again;
// at 3 us/cloud that is > 50 MB of fills, and tools/l2_probe.cu shows the L2 keeps a re-read line for ~30 MB of intervening
// fills (73 % hits at 32 MB, 11 % at 64 MB; 59 % at 64 MB with evict_last/evict_first hints, which the loads here carry).
#include <cuda_bf16.h>
#include <stdio.h>
#include <vector>

#include "tc_gemm.cuh"
#include "kernels.h"

namespace epc {
namespace hf {

using namespace tc;

constexpr int BK = 64;                                  // bf16 elements / k rows per pipeline stage
constexpr uint32_t A_BYTES = TC_BM * 128;               // 16 KB: one k-block of a 128-row H tile (assignment, K-major)
constexpr uint32_t WC_BYTES = 64 * 128;                 // 8 KB: one k-block of Wc^T [64 x 1024]
constexpr int WC_KB = 1024 / BK;                        // 16 k-blocks: 128 KB resident
constexpr int A_STAGES = 6;
constexpr uint32_t V_STAGE_BYTES = A_BYTES + 64 * 128;  // 24 KB: H [64 k rows x 128 features] (two boxes) + S' [64 k rows x 64]
constexpr int V_STAGES = 8;
constexpr size_t DATA_BYTES = (size_t)WC_KB * WC_BYTES + (size_t)A_STAGES * A_BYTES;       // 224 KB (VLAD role: 8 x 24 = 192 KB)
static_assert((size_t)V_STAGES * V_STAGE_BYTES <= DATA_BYTES, "VLAD ring must fit the shared allocation");
constexpr int MAX_STAGES = V_STAGES > A_STAGES ? V_STAGES : A_STAGES;
constexpr size_t SMEM_BYTES = 1024 + DATA_BYTES + 1024 /*column-sum scratch*/ + 8 * (2 * MAX_STAGES + 5) + 64;

struct Params {
    GemmParams pa;              // assignment: M = rows of the sub-batch, N = 64, K = 1024 (EPI_ASSIGN fields)
    GemmParams pv;              // VLAD: M = 1024, N = 64, K = points per split slab, k_batch_rows = points per cloud (EPI_STORE_F32)
    int n_assign;               // CTAs [0, n_assign) run the assignment, the rest VLAD
    int clouds, tiles_per_cloud;
    int l2_hints;               // eviction-priority hints on the H loads (EPC_HEAD_L2_HINTS=0 disables)
    int reverse;                // both roles walk the clouds from the last to the first (conv5 wrote the last ones last)
    unsigned long long* dbg;    // timeline (tuning aid, EPC_HEAD_DEBUG=1): [clouds] assignment done, then [clouds][2 splits][2] VLAD start/end of m_tile 0
    int* ready;                 // [clouds] row tiles of the cloud whose S' is in memory (zeroed before the launch)
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// L2 eviction-priority policies for the TMA loads: the assignment's read of H must survive until VLAD re-reads it
// (evict_last), VLAD's read is the last use (evict_first) -- otherwise the L2's insertion policy sacrifices the newly
// streamed lines to stale ones and the re-read misses (measured: 16 MB/cloud of DRAM reads instead of 8.5).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(192, 1)
assign_vlad_kernel(const __grid_constant__ CUtensorMap tmHk, const __grid_constant__ CUtensorMap tmWc,
                   const __grid_constant__ CUtensorMap tmHmn, const __grid_constant__ CUtensorMap tmSmn, const Params P) {
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
    constexpr uint32_t TMEM_COLS = 128;                                             // 2 x 64 accumulator columns, either role

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(assign_role ? &tmHk : &tmHmn);
        tma_prefetch_desc(assign_role ? &tmWc : &tmSmn);
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
        // ---------------- assignment: the persistent B-resident GEMM of tc_gemm.cuh, tiles in cloud order ----------------------
        const GemmParams& p = P.pa;
        uint8_t* sB = base;                                              // [16][8 KB]  Wc^T, resident
        uint8_t* sA = sB + (size_t)WC_KB * WC_BYTES;                     // [A_STAGES][16 KB]
        const int num_m_tiles = (p.M + TC_BM - 1) / TC_BM;
        const int cta = blockIdx.x, stride = P.n_assign;
        auto tile_of = [&](int mt) { return P.reverse ? num_m_tiles - 1 - mt : mt; };
        if (warp == 0) {
            if (lane == 0) {
                mbar_expect_tx(b_full, (uint32_t)WC_KB * WC_BYTES);
                for (int kb = 0; kb < WC_KB; ++kb) tma_load_2d(sB + (size_t)kb * WC_BYTES, &tmWc, b_full, kb * BK, 0);
                int s = 0;
                uint32_t ph = 0;
                const uint64_t keep = (P.l2_hints & 1) ? l2_policy_evict_last() : 0;
                for (int mt = cta; mt < num_m_tiles; mt += stride) {
                    for (int kb = 0; kb < WC_KB; ++kb) {
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], A_BYTES);
                        if (P.l2_hints & 1)
                            tma_load_2d_hint(sA + (size_t)s * A_BYTES, &tmHk, &full[s], kb * BK, tile_of(mt) * TC_BM, keep);
                        else
                            tma_load_2d(sA + (size_t)s * A_BYTES, &tmHk, &full[s], kb * BK, tile_of(mt) * TC_BM);
                        if (++s == A_STAGES) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                constexpr uint32_t idesc = make_idesc(1 /*bf16*/, TC_BM, 64, 0, 0);
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
                        const uint32_t a_addr = smem_u32(sA + (size_t)s * A_BYTES);
                        const uint32_t b_addr = smem_u32(sB + (size_t)kb * WC_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk)
                            mma_ss<true>(tmem_d, smem_desc_sw128(a_addr + kk * 32, 16, 1024), smem_desc_sw128(b_addr + kk * 32, 16, 1024),
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
                epilogue_tile<64, EPI_ASSIGN>(p, c);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]);
                // publish: every epilogue thread's S' stores are fenced, the four warps meet, one release-add per tile
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (c.epi_tid == 0) {
                    __threadfence();
                    const int old = atomicAdd(P.ready + tile_of(mt) / P.tiles_per_cloud, 1);
                    if (P.dbg && old == P.tiles_per_cloud - 1) P.dbg[tile_of(mt) / P.tiles_per_cloud] = gtime();
                }
            }
        }
    } else {
        // ---------------- VLAD: V[cloud][128 features of m_tile, 64] (+)= H^T S' over one split-K slab per work item -------------
        const GemmParams& p = P.pv;
        const int cta = blockIdx.x - P.n_assign, stride = gridDim.x - P.n_assign;
        const int items_per_cloud = (p.M / TC_BM) * p.splitk;
        const int n_items = P.clouds * items_per_cloud;
        const int nkb = p.K / BK;
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
                const uint64_t last_use = (P.l2_hints & 2) ? l2_policy_evict_first() : 0;
                for (int it = cta; it < n_items; it += stride) {
                    int cloud, m_tile, split;
                    decode(it, cloud, m_tile, split);
                    {   // the cloud's soft assignment must be complete (and visible to the TMA reads that follow)
                        long long spins = 0;
                        while (ld_acquire(P.ready + cloud) < P.tiles_per_cloud) {
                            __nanosleep(64);
                            if (++spins > (1ll << 26)) __trap();          // several seconds: the assignment role is not running
                        }
                        asm volatile("fence.proxy.async;" ::: "memory");
                    }
                    const unsigned long long t_start = P.dbg ? gtime() : 0ull;
                    const int krow0 = cloud * p.k_batch_rows + split * p.K;
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], V_STAGE_BYTES);
                        uint8_t* a = base + (size_t)s * V_STAGE_BYTES;
                        uint8_t* b = a + A_BYTES;
                        const int k0 = krow0 + kb * BK;
                        if (P.l2_hints & 2) {                                                       // boxes {64 features, 64 k rows}
                            tma_load_2d_hint(a, &tmHmn, &full[s], m_tile * TC_BM, k0, last_use);
                            tma_load_2d_hint(a + (size_t)BK * 128, &tmHmn, &full[s], m_tile * TC_BM + 64, k0, last_use);
                        } else {
                            tma_load_2d(a, &tmHmn, &full[s], m_tile * TC_BM, k0);
                            tma_load_2d(a + (size_t)BK * 128, &tmHmn, &full[s], m_tile * TC_BM + 64, k0);
                        }
                        tma_load_2d(b, &tmSmn, &full[s], 0, k0);                                    // box {64 clusters, 64 k rows}
                        if (++s == V_STAGES) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                    if (P.dbg && m_tile == 0 && split < 2) {
                        P.dbg[P.clouds + (cloud * 2 + split) * 2] = t_start;
                        P.dbg[P.clouds + (cloud * 2 + split) * 2 + 1] = gtime();
                    }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                constexpr uint32_t idesc = make_idesc(1 /*bf16*/, TC_BM, 64, 1, 1);       // both operands MN-major
                constexpr uint32_t step = 16 * 128, lbo = BK * 128;
                int s = 0, tile = 0;
                uint32_t ph = 0;
                for (int it = cta; it < n_items; it += stride, ++tile) {
                    const int buf = tile & 1;
                    mbar_wait(&tempty[buf], ((tile >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 64);
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(base + (size_t)s * V_STAGE_BYTES);
                        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk)
                            mma_ss<true>(tmem_d, smem_desc_sw128(a_addr + kk * step, lbo, 1024), smem_desc_sw128(b_addr + kk * step, lbo, 1024),
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
                c.trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 64);
                c.m0 = m_tile * TC_BM; c.m = c.m0 + row; c.row = row; c.lane = lane; c.n0 = 0; c.mtile = m_tile;
                c.c_off = (long long)cloud * p.c_batch + (long long)split * p.c_slab;
                c.scratch = scratch; c.epi_tid = threadIdx.x - 64; c.bias = nullptr;
                c.col_begin = 0; c.col_end = 64; c.nparts = 1; c.npart = 0; c.warp_slot = warp - 2;
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

}  // namespace hf

// Soft assignment of the rows of H [clouds * N, 1024] (bf16) and the VLAD sums of every cloud, one launch (see the head of this
// file).  Outputs as tc_assign + tc_vlad (tc_gemm.cu): S [R, 64] bf16, a_part [R/128, 64], V fp32 slabs [splitk][.., 1024, 64].
// `ready`: clouds ints of scratch.
int tc_assign_vlad(const __nv_bfloat16* H, int clouds, int N, const __nv_bfloat16* Wct, const float* rowss, int parts,
                   const float* bn_scale, const float* bn_shift, __nv_bfloat16* S, float* a_part, float* V, int splitk, long long slab,
                   int* ready, cudaStream_t st) {
    const long long R = (long long)clouds * N;
    EPC_CHECK_ARG(clouds >= 1 && N % 128 == 0 && (N / splitk) % hf::BK == 0, "tc_assign_vlad: bad shape clouds=%d N=%d splitk=%d", clouds, N, splitk);
    hf::Params P = {};
    P.pa.M = (int)R; P.pa.N = 64; P.pa.K = 1024; P.pa.splitk = 1; P.pa.C = S; P.pa.ldc = 64; P.pa.aux = a_part; P.pa.rowss = rowss;
    P.pa.rowss_parts = parts; P.pa.bn_scale = bn_scale; P.pa.bn_shift = bn_shift;
    P.pv.M = 1024; P.pv.N = 64; P.pv.K = N / splitk; P.pv.k_batch_rows = N; P.pv.splitk = splitk; P.pv.C = V; P.pv.ldc = 64;
    P.pv.c_batch = 1024ll * 64; P.pv.c_slab = slab;
    P.clouds = clouds; P.tiles_per_cloud = N / 128; P.ready = ready;
    P.reverse = getenv("EPC_ASSIGN_FORWARD") ? 0 : 1;
    P.l2_hints = getenv("EPC_HEAD_L2_HINTS") ? atoi(getenv("EPC_HEAD_L2_HINTS")) : 3;     // bit 0: assignment evict_last, bit 1: VLAD evict_first
    static unsigned long long* dbg_buf = nullptr;
    if (getenv("EPC_HEAD_DEBUG")) {
        if (!dbg_buf) EPC_CUDA(cudaMalloc(&dbg_buf, sizeof(unsigned long long) * 5 * 512));
        EPC_CUDA(cudaMemsetAsync(dbg_buf, 0, sizeof(unsigned long long) * 5 * 512, st));
        P.dbg = dbg_buf;
    }
    const int sms = sm_count();
    static const int env_assign = getenv("EPC_HEAD_ASSIGN_CTAS") ? atoi(getenv("EPC_HEAD_ASSIGN_CTAS")) : 0;
    int n_assign = env_assign > 0 ? env_assign : (sms * 90 + 74) / 148;       // measured best split on B200 (148 SMs): 90 / 58
    if (n_assign < 1) n_assign = 1;
    if (n_assign > sms - 1) n_assign = sms - 1;
    P.n_assign = n_assign;
    CUtensorMap tmHk, tmWc, tmHmn, tmSmn;
    if (int rc = make_tmap_2d(&tmHk, H, (uint64_t)R, 1024, 1024, hf::BK, tc::TC_BM)) return rc;
    if (int rc = make_tmap_2d(&tmWc, Wct, 64, 1024, 1024, hf::BK, 64)) return rc;
    if (int rc = make_tmap_2d(&tmHmn, H, (uint64_t)R, 1024, 1024, 64, hf::BK)) return rc;
    if (int rc = make_tmap_2d(&tmSmn, S, (uint64_t)R, 64, 64, 64, hf::BK)) return rc;
    EPC_CUDA(cudaMemsetAsync(ready, 0, sizeof(int) * (size_t)clouds, st));
    static PerDeviceSize attr;
    EPC_CUDA(ensure_dyn_smem(hf::assign_vlad_kernel, hf::SMEM_BYTES, attr));
    hf::assign_vlad_kernel<<<sms, 192, hf::SMEM_BYTES, st>>>(tmHk, tmWc, tmHmn, tmSmn, P);
    EPC_LAUNCH_CHECK();
    if (P.dbg && clouds <= 512) {
        std::vector<unsigned long long> h((size_t)5 * clouds);
        EPC_CUDA(cudaStreamSynchronize(st));
        EPC_CUDA(cudaMemcpy(h.data(), P.dbg, sizeof(unsigned long long) * 5 * clouds, cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull;
        for (auto v : h) if (v && v < t0) t0 = v;
        for (int c = 0; c < clouds; ++c)
            fprintf(stderr, "HF cloud %3d assign_done %8.2f vlad0 %8.2f..%8.2f vlad1 %8.2f..%8.2f us\n", c, (h[c] - t0) * 1e-3,
                    (h[clouds + 4 * c] - t0) * 1e-3, (h[clouds + 4 * c + 1] - t0) * 1e-3, (h[clouds + 4 * c + 2] - t0) * 1e-3,
                    (h[clouds + 4 * c + 3] - t0) * 1e-3);
    }
    return EPC_OK;
}

}  // namespace epc
