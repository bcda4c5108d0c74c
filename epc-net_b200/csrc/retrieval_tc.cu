// K6 scoring on tcgen05, with the candidate selection fused into the TMEM-drain epilogue: the Q x D score matrix
// (243 MB at the Oxford-scale problem) never exists.  Replaces the scoring half of `KDTree(database_output).query(q, k)`
// (evaluate.py:463,481); retrieval.cu holds the float64 re-rank, the proof of exactness and the exact fallback.
//
//   operands   q = qh + ql, d = dh + dl (bf16 pairs): q.d ~ qh.dh + ql.dh + qh.dl, fp32 accumulation in TMEM -- the same
//              2^-16 relative accuracy as a 3xTF32 split at twice the tensor rate and with dh streamed once for two products.
//   CTA        one 128-query tile x one contiguous range of 128-row database tiles.  The query tile [qh | ql] (128 KB at
//              dim 256) stays resident in shared memory; the database k-blocks stream through a 5-deep TMA ring;
//              accumulators are double-buffered in TMEM, so the epilogue of tile i overlaps the MMAs of tile i+1.
//   epilogue   thread = query (TMEM lane): score_j = (|q|^2 + |d_j|^2) - 2 q.d_j for the tile's 128 columns, then
//                DENSE  store the scores (small databases / the threshold sample);
//                EMIT   append (score, j) to the query's private candidate region iff score <= thr_q, where thr_q is the
//                       query's 32nd smallest score over a 1024-row SAMPLE of the database (a valid upper bound of its
//                       32nd smallest score overall): ~32 D / 1024 entries per query instead of D scores.
#include <cuda_bf16.h>

#include "tc_gemm.cuh"
#include "kernels.h"

namespace epc {
namespace rt {

using namespace tc;

constexpr int BM = 128, BN = 128, BK = 64;            // bf16 elements per 128-byte swizzled row
constexpr uint32_t TILE_BYTES = 128 * 128;            // one k-block of a 128-row operand tile
constexpr int STAGES = 5;
constexpr int MAX_KB = 4;                             // dim <= 256

struct Params {
    int Q, D, kbd;                 // kbd = dim / 64
    int n_tiles, tiles_per_range, n_ranges;
    int tile_stride;               // tile t of this launch is database tile t * tile_stride (the threshold sample strides the database)
    const float* qn;               // [Q]     |q|^2
    const float* dn;               // [>= n_tiles * 128] |d|^2, +inf beyond D
    const float* thr;              // EMIT: [Q]
    float* scores;                 // DENSE: [Q, ld]
    int ld;
    uint2* cand;                   // EMIT: [Q, n_ranges, cap] (score bits, database row)
    int* cand_count;               // EMIT: [Q, n_ranges]; -1 = the region overflowed
    int cap;
};

enum { MODE_DENSE = 0, MODE_EMIT = 1 };

template <int MODE>
__global__ void __launch_bounds__(192, 1)
retr_score_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmD, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int kbd = p.kbd;
    uint8_t* sA = base;                                              // [2 kbd][16 KB]: qh k-blocks, then ql k-blocks
    uint8_t* sB = sA + (size_t)2 * MAX_KB * TILE_BYTES;              // [STAGES][16 KB]
    uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + (size_t)STAGES * TILE_BYTES);
    uint64_t* full = a_full + 1;
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;          // [2]
    uint64_t* tempty = tfull + 2;              // [2] (4 arrivals: one per epilogue warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM;
    const int range = blockIdx.x;
    const int nt0 = range * p.tiles_per_range;
    const int nt1 = min(p.n_tiles, nt0 + p.tiles_per_range);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmD);
        mbar_init(a_full, 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 2 * BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(a_full, (uint32_t)(2 * kbd) * TILE_BYTES);
            for (int kb = 0; kb < 2 * kbd; ++kb) tma_load_2d(sA + (size_t)kb * TILE_BYTES, &tmQ, a_full, kb * BK, m0);
            int it = 0;
            for (int nt = nt0; nt < nt1; ++nt) {
                for (int kb = 0; kb < kbd; ++kb) {
#pragma unroll
                    for (int half = 0; half < 2; ++half, ++it) {          // dh k-block, then dl k-block
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], TILE_BYTES);
                        tma_load_2d(sB + (size_t)s * TILE_BYTES, &tmD, &full[s], (half * kbd + kb) * BK, nt * p.tile_stride * BN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(1 /*bf16*/, BM, BN, 0, 0);
            mbar_wait(a_full, 0);
            int it = 0, tile = 0;
            for (int nt = nt0; nt < nt1; ++nt, ++tile) {
                const int buf = tile & 1;
                mbar_wait(&tempty[buf], ((tile >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < kbd; ++kb) {
                    const uint32_t qh = smem_u32(sA + (size_t)kb * TILE_BYTES), ql = smem_u32(sA + (size_t)(kbd + kb) * TILE_BYTES);
                    {   // dh: qh.dh + ql.dh
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(sB + (size_t)s * TILE_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk)
                            mma_ss<true>(tmem_d, smem_desc_sw128(qh + kk * 32, 16, 1024), smem_desc_sw128(b_addr + kk * 32, 16, 1024),
                                         idesc, (kb | kk) != 0);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk)
                            mma_ss<true>(tmem_d, smem_desc_sw128(ql + kk * 32, 16, 1024), smem_desc_sw128(b_addr + kk * 32, 16, 1024),
                                         idesc, 1);
                        mma_commit(&empty[s]);
                        ++it;
                    }
                    {   // dl: qh.dl
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(sB + (size_t)s * TILE_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk)
                            mma_ss<true>(tmem_d, smem_desc_sw128(qh + kk * 32, 16, 1024), smem_desc_sw128(b_addr + kk * 32, 16, 1024),
                                         idesc, 1);
                        mma_commit(&empty[s]);
                        ++it;
                    }
                }
                mma_commit(&tfull[buf]);
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int m = m0 + row;
        const bool live = m < p.Q;
        const float qq = live ? __ldg(p.qn + m) : 0.f;
        float thr = 0.f;
        uint2* region = nullptr;
        int cnt = 0;
        if (MODE == MODE_EMIT) {
            thr = live ? __ldg(p.thr + m) : -INFINITY;
            region = p.cand + ((size_t)m * p.n_ranges + range) * p.cap;
        }
        int tile = 0;
        for (int nt = nt0; nt < nt1; ++nt, ++tile) {
            const int buf = tile & 1;
            mbar_wait(&tfull[buf], (tile >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                float v[32];
                tmem_ld32(trow + (uint32_t)c0, v);
                const int j0 = nt * p.tile_stride * BN + c0;        // database row of column c0
                const float4* dn4 = reinterpret_cast<const float4*>(p.dn + j0);
                if (MODE == MODE_DENSE) {
                    if (live) {
                        float4* dst = reinterpret_cast<float4*>(p.scores + (size_t)m * p.ld + nt * BN + c0);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 dd = __ldg(dn4 + i);
                            dst[i] = make_float4(fmaf(-2.0f, v[4 * i], qq + dd.x), fmaf(-2.0f, v[4 * i + 1], qq + dd.y),
                                                 fmaf(-2.0f, v[4 * i + 2], qq + dd.z), fmaf(-2.0f, v[4 * i + 3], qq + dd.w));
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 dd = __ldg(dn4 + i);
                        const float sc[4] = {fmaf(-2.0f, v[4 * i], qq + dd.x), fmaf(-2.0f, v[4 * i + 1], qq + dd.y),
                                             fmaf(-2.0f, v[4 * i + 2], qq + dd.z), fmaf(-2.0f, v[4 * i + 3], qq + dd.w)};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (sc[e] <= thr) {
                                if (cnt < p.cap) region[cnt] = make_uint2(__float_as_uint(sc[e]), (uint32_t)(j0 + 4 * i + e));
                                ++cnt;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
        if (MODE == MODE_EMIT && live) p.cand_count[(size_t)m * p.n_ranges + range] = (cnt <= p.cap) ? cnt : -1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

}  // namespace rt

// bf16 pair split of X [R, dim] -> out [R, 2 dim] = (hi | lo), hi = bf16(x), lo = bf16(x - hi); also |x|^2 (fp32, sequential
// FMA over the row, one warp per row) for rows < R, +inf for the padding rows up to Rpad (norms only)
__global__ void split2_bf16_kernel(const float* __restrict__ X, int R, int Rpad, int dim, __nv_bfloat16* __restrict__ out,
                                   float* __restrict__ norms) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= Rpad) return;
    if (r >= R) {
        if (lane == 0) norms[r] = INFINITY;
        return;
    }
    float ss = 0.f;
    for (int i = lane; i < dim; i += 32) {
        const float x = X[(size_t)r * dim + i];
        const __nv_bfloat16 hi = __float2bfloat16_rn(x);
        const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
        out[(size_t)r * 2 * dim + i] = hi;
        out[(size_t)r * 2 * dim + dim + i] = lo;
        ss = fmaf(x, x, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) norms[r] = ss;
}

int retr_split2(const float* X, int R, int Rpad, int dim, __nv_bfloat16* out, float* norms, cudaStream_t st) {
    if (Rpad == 0) return EPC_OK;
    split2_bf16_kernel<<<(Rpad + 7) / 8, 256, 0, st>>>(X, R, Rpad, dim, out, norms);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

bool retr_tc_supported(int dim) { return dim % rt::BK == 0 && dim >= rt::BK && dim <= rt::BK * rt::MAX_KB; }

static size_t retr_smem() {
    return 1024 + (size_t)(2 * rt::MAX_KB + rt::STAGES) * rt::TILE_BYTES + 8 * (2 + 2 * rt::STAGES + 4) + 64;
}

// number of database-tile ranges the tiles of one query tile are split into (one CTA each)
int retr_ranges(int Q, int n_tiles) {
    const int m_tiles = (Q + rt::BM - 1) / rt::BM;
    int sms = 148;
    {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            sms = v;
    }
    int r = sms / (m_tiles > 0 ? m_tiles : 1);
    if (r < 1) r = 1;
    if (r > n_tiles) r = n_tiles;
    return r;
}

// q2 [Q, 2 dim] / db2 [D, 2 dim] bf16 pairs; n_tiles tiles of 128 database rows, tile t = database tile t * tile_stride.
//   dense: scores [Q, ld], column 128 t + c = database row 128 t tile_stride + c
//   emit : cand [Q, n_ranges, cap], cand_count [Q, n_ranges], thr [Q]
int retr_scores(const __nv_bfloat16* q2, int Q, const __nv_bfloat16* db2, int D, int dim, int n_tiles, int tile_stride, int n_ranges,
                const float* qn, const float* dn, float* scores, int ld, const float* thr, uint2* cand, int* cand_count, int cap,
                cudaStream_t st) {
    EPC_CHECK_ARG(retr_tc_supported(dim), "retr_scores: dim=%d unsupported by the tensor-core path", dim);
    if (Q == 0 || n_tiles == 0) return EPC_OK;
    CUtensorMap tmQ, tmD;
    if (int rc = make_tmap_2d(&tmQ, q2, (uint64_t)Q, (uint64_t)2 * dim, (uint64_t)2 * dim, rt::BK, rt::BM)) return rc;
    if (int rc = make_tmap_2d(&tmD, db2, (uint64_t)D, (uint64_t)2 * dim, (uint64_t)2 * dim, rt::BK, rt::BN)) return rc;
    rt::Params p = {};
    p.Q = Q; p.D = D; p.kbd = dim / rt::BK; p.n_tiles = n_tiles; p.n_ranges = n_ranges; p.tile_stride = tile_stride;
    p.tiles_per_range = (n_tiles + n_ranges - 1) / n_ranges;
    p.qn = qn; p.dn = dn; p.thr = thr; p.scores = scores; p.ld = ld; p.cand = cand; p.cand_count = cand_count; p.cap = cap;
    const size_t smem = retr_smem();
    static PerDeviceSize attr_d, attr_e;
    dim3 grid(n_ranges, (Q + rt::BM - 1) / rt::BM);
    if (scores) {
        EPC_CUDA(ensure_dyn_smem(rt::retr_score_kernel<rt::MODE_DENSE>, smem, attr_d));
        rt::retr_score_kernel<rt::MODE_DENSE><<<grid, 192, smem, st>>>(tmQ, tmD, p);
    } else {
        EPC_CUDA(ensure_dyn_smem(rt::retr_score_kernel<rt::MODE_EMIT>, smem, attr_e));
        rt::retr_score_kernel<rt::MODE_EMIT><<<grid, 192, smem, st>>>(tmQ, tmD, p);
    }
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

}  // namespace epc
