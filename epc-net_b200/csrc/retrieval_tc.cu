// K6 scoring on tcgen05, with the candidate selection fused into the TMEM-drain epilogue: the Q x D score matrix
// (243 MB at the Oxford-scale problem) never exists.  Replaces the scoring half of `KDTree(database_output).query(q, k)`
// (evaluate.py:463,481); retrieval.cu holds the float64 re-rank, the proof of exactness and the exact fallback.
//
//   score      s_ij = |d_j|^2 - 2 q_i.d_j  (= |q_i - d_j|^2 - |q_i|^2: the same order per query, one FFMA per accumulator)
//   operands   q = qh + ql, d = dh + dl (bf16 pairs): q.d ~ qh.dh + ql.dh + qh.dl, fp32 accumulation in TMEM -- the same
//              2^-16 relative accuracy as a 3xTF32 split at twice the tensor rate and with dh streamed once for two products.
//   CTA        one 128-query tile x one contiguous range of 128-row database tiles.  The query tile [qh | ql] is the A operand
//              and lives in TENSOR MEMORY (tcgen05.mma with A from TMEM: 128 lanes x dim 32-bit columns of packed bf16 pairs,
//              written once with tcgen05.st), so an MMA reads only its B slice from shared memory (64 B/clk instead of the
//              128 B/clk -- all of the shared-memory bandwidth -- that a 128 x 128 shared/shared MMA needs) and the whole
//              208 KB of shared memory is a 13-deep TMA ring of database k-blocks.  Accumulators are double-buffered in the
//              other half of TMEM (2 x 128 columns): the epilogue of tile i overlaps the MMAs of tile i+1.
//   epilogue   8 warps: thread = (query = TMEM lane, half of the tile's 128 columns)
//                DENSE   store the scores (small databases);
//                DENSE   store the scores (small databases; the threshold sample: every stride-th tile of a large one);
//                EMIT    append (score, j) to the thread's private candidate region iff score <= thr_q, thr_q = an upper
//                        bound of the query's 32nd smallest score taken from the sample (retrieval.cu).
#include <cuda_bf16.h>

#include "tc_gemm.cuh"
#include "kernels.h"

namespace epc {
namespace rt {

using namespace tc;

constexpr int BM = 128, BN = 128, BK = 64;            // BK bf16 elements = one 128-byte swizzled row
constexpr uint32_t B_BYTES = BN * 128;                // one k-block of a database tile
constexpr int MAX_KB = 4;                             // dim <= 256: [qh | ql] fills 256 TMEM columns
constexpr int STAGES = 13;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t TMEM_A = 2 * BN;                   // first TMEM column of the A operand

struct Params {
    int Q, dim, kbd;               // kbd = dim / 64
    int n_tiles, tiles_per_range, n_ranges;
    int tile_stride;               // tile t of this launch is database tile t * tile_stride (the threshold sample strides the database)
    const __nv_bfloat16* q2;       // [Q, 2 dim] (hi | lo)
    const float* dn;               // [>= 128 * database tiles] |d|^2, +inf beyond D
    float* scores;                 // DENSE: [Q, ld]
    int ld;
    const float* thr;              // EMIT: [Q]
    uint2* cand;                   // EMIT: [Q, 2 n_ranges, cap] (score bits, database row)
    int* cand_count;               // EMIT: [Q, 2 n_ranges]; -1 = the region overflowed
    int cap;
};

enum { MODE_DENSE = 0, MODE_EMIT = 1 };

#define EPC_R32(r) r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14], r[15], r[16], \
                   r[17], r[18], r[19], r[20], r[21], r[22], r[23], r[24], r[25], r[26], r[27], r[28], r[29], r[30], r[31]

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// wait for every tcgen05.ld of this thread; the registers are in-out operands so that no use is scheduled above the wait
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
// registers -> TMEM: this thread's lane, 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc], bf16 operands; issued by ONE thread
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}

// per-thread epilogue state
struct Epi {
    float thr;          // EMIT
    uint2* region;      // EMIT
    int cnt;            // EMIT
};

// 32 accumulator columns of one query: database rows j0 .. j0 + 31
template <int MODE>
__device__ __forceinline__ void epi_chunk(const Params& p, Epi& e, const uint32_t (&r)[32], int j0, float* dst, bool live) {
    const float4* dn4 = reinterpret_cast<const float4*>(p.dn + j0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 dd = __ldg(dn4 + i);
        const float sc[4] = {fmaf(-2.0f, __uint_as_float(r[4 * i]), dd.x), fmaf(-2.0f, __uint_as_float(r[4 * i + 1]), dd.y),
                             fmaf(-2.0f, __uint_as_float(r[4 * i + 2]), dd.z), fmaf(-2.0f, __uint_as_float(r[4 * i + 3]), dd.w)};
        if (MODE == MODE_DENSE) {
            if (live) reinterpret_cast<float4*>(dst)[i] = make_float4(sc[0], sc[1], sc[2], sc[3]);
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (sc[u] <= e.thr) {
                    if (e.cnt < p.cap) e.region[e.cnt] = make_uint2(__float_as_uint(sc[u]), (uint32_t)(j0 + 4 * i + u));
                    ++e.cnt;
                }
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
retr_score_kernel(const __grid_constant__ CUtensorMap tmD, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int kbd = p.kbd;
    uint8_t* sB = base;                                              // [STAGES][16 KB]
    uint64_t* a_full = reinterpret_cast<uint64_t*>(base + (size_t)STAGES * B_BYTES);
    uint64_t* full = a_full + 1;
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;          // [2]
    uint64_t* tempty = tfull + 2;              // [2] (EPI_WARPS arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM;
    const int range = blockIdx.x;
    const int nt0 = range * p.tiles_per_range;
    const int nt1 = min(p.n_tiles, nt0 + p.tiles_per_range);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmD);
        mbar_init(a_full, EPI_WARPS);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int nt = nt0; nt < nt1; ++nt) {
                const int row0 = nt * p.tile_stride * BN;
                for (int kb = 0; kb < kbd; ++kb) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {          // dh k-block, then dl k-block
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], B_BYTES);
                        tma_load_2d(sB + (size_t)s * B_BYTES, &tmD, &full[s], (half * kbd + kb) * BK, row0);
                        if (++s == STAGES) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        {       // all 32 lanes, converged (warp-uniform descriptors stay in uniform registers); one elected lane issues -- a loop
                // nested in `if (lane == 0)` spends ~25 instructions per MMA and cannot keep up with a 64-clk 128x128x16 MMA
            constexpr uint32_t idesc = make_idesc(1 /*bf16*/, BM, BN, 0, 0);
            mbar_wait(a_full, 0);
            tc_fence_after();
            const uint32_t a_hi = tmem_base + TMEM_A, a_lo = a_hi + (uint32_t)(p.dim / 2);      // 32 columns per 64-element k-block
            const uint32_t b_lo0 = desc_lo_sw128(smem_u32(sB));
            int s = 0, tile = 0;
            uint32_t ph = 0;
            for (int nt = nt0; nt < nt1; ++nt, ++tile) {
                const int buf = tile & 1;
                mbar_wait(&tempty[buf], ((tile >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < kbd; ++kb) {
                    {   // dh: qh.dh + ql.dh
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t b_lo = b_lo0 + (uint32_t)s * (uint32_t)(B_BYTES >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < BK / 16; ++kk)
                                mma_ts(tmem_d, a_hi + (uint32_t)(kb * 32 + kk * 8), desc_of(b_lo + 2 * kk), idesc, (kb | kk) != 0);
#pragma unroll
                            for (int kk = 0; kk < BK / 16; ++kk)
                                mma_ts(tmem_d, a_lo + (uint32_t)(kb * 32 + kk * 8), desc_of(b_lo + 2 * kk), idesc, 1);
                            mma_commit(&empty[s]);
                        }
                        __syncwarp();
                        if (++s == STAGES) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                    {   // dl: qh.dl
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t b_lo = b_lo0 + (uint32_t)s * (uint32_t)(B_BYTES >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < BK / 16; ++kk)
                                mma_ts(tmem_d, a_hi + (uint32_t)(kb * 32 + kk * 8), desc_of(b_lo + 2 * kk), idesc, 1);
                            mma_commit(&empty[s]);
                        }
                        __syncwarp();
                        if (++s == STAGES) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                }
                if (elect_one()) mma_commit(&tfull[buf]);
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;               // which 64 of the tile's 128 columns (and which half of the A columns)
        const int m = m0 + q * 32 + lane;
        const bool live = m < p.Q;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        {   // the A operand: this query's [hi | lo] row, 2 bf16 per 32-bit TMEM column; this thread writes half of the dim columns
            const int words = p.dim / 2;                // columns written by this thread (a multiple of 32)
            const uint4* src = reinterpret_cast<const uint4*>(p.q2 + (size_t)(live ? m : 0) * 2 * p.dim) + half * (words / 4);
            for (int c0 = 0; c0 < words; c0 += 64) {                // two 32-column chunks per round: 16 loads in flight
                uint32_t r0[32], r1[32];
                const bool two = c0 + 32 < words;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 w = live ? __ldg(src + c0 / 4 + i) : make_uint4(0u, 0u, 0u, 0u);
                    r0[4 * i] = w.x; r0[4 * i + 1] = w.y; r0[4 * i + 2] = w.z; r0[4 * i + 3] = w.w;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 w = (live && two) ? __ldg(src + c0 / 4 + 8 + i) : make_uint4(0u, 0u, 0u, 0u);
                    r1[4 * i] = w.x; r1[4 * i + 1] = w.y; r1[4 * i + 2] = w.z; r1[4 * i + 3] = w.w;
                }
                tmem_st32(lane_addr + TMEM_A + (uint32_t)(half * words + c0), r0);
                if (two) tmem_st32(lane_addr + TMEM_A + (uint32_t)(half * words + c0 + 32), r1);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
        const int slot = range * 2 + half;
        Epi e;
        e.thr = -INFINITY; e.region = nullptr; e.cnt = 0;
        if (MODE == MODE_EMIT && live) {
            e.thr = __ldg(p.thr + m);
            e.region = p.cand + ((size_t)m * (2 * p.n_ranges) + slot) * p.cap;
        }
        int tile = 0;
        for (int nt = nt0; nt < nt1; ++nt, ++tile) {
            const int buf = tile & 1;
            mbar_wait(&tfull[buf], (tile >> 1) & 1);
            tc_fence_after();
            const uint32_t tcol = lane_addr + (uint32_t)(buf * BN + half * 64);
            const int j0 = nt * p.tile_stride * BN + half * 64;                              // database row of this thread's first column
            float* dst = (MODE == MODE_DENSE && live) ? p.scores + (size_t)m * p.ld + nt * BN + half * 64 : nullptr;
            uint32_t ra[32], rb[32];
            tmem_ld32_issue(tcol, ra);
            tmem_ld32_wait(ra);
            tmem_ld32_issue(tcol + 32u, rb);
            epi_chunk<MODE>(p, e, ra, j0, dst, live);
            tmem_ld32_wait(rb);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);               // the accumulator is in registers: the tensor core may refill it
            epi_chunk<MODE>(p, e, rb, j0 + 32, dst + 32, live);
        }
        if (MODE == MODE_EMIT && live) p.cand_count[(size_t)m * (2 * p.n_ranges) + slot] = (e.cnt <= p.cap) ? e.cnt : -1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace rt

// bf16 pair split of X [R, dim] -> out [R, 2 dim] = (hi | lo), hi = bf16(x), lo = bf16(x - hi); also |x|^2 (fp32, one warp per
// row) for rows < R, +inf for the padding rows up to Rpad (norms only)
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

static size_t retr_smem() { return 1024 + (size_t)rt::STAGES * rt::B_BYTES + 8 * (1 + 2 * rt::STAGES + 4) + 64; }

// number of database-tile ranges the (128-row) tiles of one query tile are split into (one CTA each)
int retr_ranges(int Q, int n_tiles) {
    const int m_tiles = (Q + rt::BM - 1) / rt::BM;
    int r = sm_count() / (m_tiles > 0 ? m_tiles : 1);
    if (r > n_tiles) r = n_tiles;
    if (r < 1) return 1;
    const int tpr = (n_tiles + r - 1) / r;
    return (n_tiles + tpr - 1) / tpr;          // no empty range
}

// q2 [Q, 2 dim] / db2 [D, 2 dim] bf16 pairs; n_tiles tiles of 128 database rows, tile t = database tile t * tile_stride,
// split into n_ranges (= retr_ranges(Q, n_tiles)) contiguous ranges.  Exactly one of:
//   scores   != NULL : DENSE,  scores [Q, ld], column 128 t + c = database row 128 t tile_stride + c
//   cand     != NULL : EMIT,   thr [Q]; cand [Q, 2 n_ranges, cap], cand_count [Q, 2 n_ranges]
int retr_scores(const __nv_bfloat16* q2, int Q, const __nv_bfloat16* db2, int D, int dim, int n_tiles, int tile_stride, int n_ranges,
                const float* dn, float* scores, int ld, const float* thr, uint2* cand, int* cand_count, int cap, cudaStream_t st) {
    EPC_CHECK_ARG(retr_tc_supported(dim), "retr_scores: dim=%d unsupported by the tensor-core path", dim);
    if (Q == 0 || n_tiles == 0) return EPC_OK;
    CUtensorMap tmD;
    if (int rc = make_tmap_2d(&tmD, db2, (uint64_t)D, (uint64_t)2 * dim, (uint64_t)2 * dim, rt::BK, rt::BN)) return rc;
    rt::Params p = {};
    p.Q = Q; p.dim = dim; p.kbd = dim / rt::BK; p.n_tiles = n_tiles; p.n_ranges = n_ranges; p.tile_stride = tile_stride;
    p.tiles_per_range = (n_tiles + n_ranges - 1) / n_ranges;
    p.q2 = q2; p.dn = dn; p.scores = scores; p.ld = ld; p.cand = cand; p.cand_count = cand_count; p.cap = cap;
    p.thr = thr;
    const size_t smem = retr_smem();
    static PerDeviceSize attr_d, attr_e;
    dim3 grid(n_ranges, (Q + rt::BM - 1) / rt::BM);
    if (scores) {
        EPC_CUDA(ensure_dyn_smem(rt::retr_score_kernel<rt::MODE_DENSE>, smem, attr_d));
        rt::retr_score_kernel<rt::MODE_DENSE><<<grid, rt::THREADS, smem, st>>>(tmD, p);
    } else {
        EPC_CUDA(ensure_dyn_smem(rt::retr_score_kernel<rt::MODE_EMIT>, smem, attr_e));
        rt::retr_score_kernel<rt::MODE_EMIT><<<grid, rt::THREADS, smem, st>>>(tmD, p);
    }
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

}  // namespace epc
