// Tensor-core GEMMs for sm_100a: tcgen05.mma (kind::tf32 / kind::f16-bf16) with TMEM accumulators, operands
// staged by TMA into 128B-swizzled shared memory, an mbarrier producer/consumer ring, one elected thread issuing
// the MMAs, four epilogue warps draining TMEM with tcgen05.ld.  Hand-written PTX; descriptor bit layouts follow the
// PTX ISA tcgen05 descriptor tables (cross-checked against cute/arch/mma_sm100_desc.hpp).
//
// One kernel template serves every dense contraction of the EPC-Net head:
//   conv5            H  = relu(Xc W5 + b)        A K-major, B K-major   epilogue CONV5_BF16 (bf16 H + row sum-of-squares)
//   conv5 (EPC-Net-L) max_n relu(Xc W5 + b)      A K-major, B K-major   epilogue COLMAX     (H never written)
//   cluster assign   S' = softmax(BN(X Wc))/|H|  A K-major, B K-major   epilogue ASSIGN     (BN = 64)
//   VLAD accumulate  V  = H^T S'                 A MN-major, B MN-major epilogue STORE_F32  (batched, split-K slabs)
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace epc {
namespace tc {

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// 2-D TMA tile load: global (tensor map) -> shared, completion on an mbarrier (bytes)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// 2-D TMA tile load multicast to every CTA of the cluster named in cta_mask: the tile lands at the same shared-memory
// offset in each of them and completes (bytes) on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
// arrive (once all previously issued MMAs of this thread completed) on the mbarrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-D TMA tile store: shared (the layout a load with the same map would have produced) -> global, as a bulk async-group of the
// issuing thread; rows / columns outside the tensor are clipped.  The writes to shared memory must have been fenced
// (fence.proxy.async) by their authors before the issuing thread was synchronised with them.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(tm)),
                 "r"(c0), "r"(c1), "r"(smem_src) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }     // sources may be reused
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }               // stores are complete
// pull a 2-D tile into the L2 only (no shared-memory destination, no barrier): hides the DRAM latency of a tile that a
// shallow shared-memory ring will ask for a few tiles later
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tm, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// TMEM management (one full warp executes alloc/dealloc)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.  KIND_F16: bf16/fp16 operands; else tf32.
template <bool KIND_F16>
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (KIND_F16) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
            "}\n"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
            "}\n"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
            : "memory");
    }
}
// One lane of a CONVERGED warp.  The MMA warp of the persistent kernel runs its loop with all 32 lanes (warp-uniform values: the
// compiler keeps the descriptors in uniform registers) and elects a lane only for the tcgen05 instructions themselves.  A loop
// nested inside `if (lane == 0)` costs ~25 instructions, a runtime modulo and an R2UR round trip per MMA: a single thread then
// issues a 128x256x16 MMA only every ~240 clk (per-tile timeline, EPC_BRES_TIMELINE) -- the tensor pipe wants one every 128.
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(p));
    return p != 0;
}
// shared-memory descriptor split in halves: the high word is the same for every K-major 128B-swizzled tile
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }
constexpr uint32_t DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc_of(uint32_t lo) { return ((uint64_t)DESC_HI_SW128 << 32) | lo; }
// arrive on an mbarrier when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this thread's lane (row), 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
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
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------------------
// Descriptors
// ---------------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for the 128B-swizzle layouts TMA writes ({128 B, rows} boxes: row r at r*128 B,
// 16-byte chunks XOR-ed with (r & 7), 8-row groups 1024 B apart):
//   bits [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 |
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
//  K-major operand : rows = M/N index, the 128 B = one K slab.  SBO = 1024 (next 8 rows); LBO unused.
//  MN-major operand: rows = K index, the 128 B = one chunk of 32 (tf32) / 64 (bf16) M/N elements.
//                    SBO = 1024 (next 8 K rows); LBO = bytes between consecutive M/N chunks (separate TMA boxes).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor: [4,6) D format = 1 (F32) | [7,10) A format | [10,13) B format (kind::f16: 0 = F16, 1 = BF16;
// kind::tf32: 2 = TF32) | [15] A major (0 = K, 1 = MN) | [16] B major | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------
enum { EPI_STORE_F32 = 0, EPI_CONV5_BF16 = 1, EPI_COLMAX = 2, EPI_ASSIGN = 3, EPI_CONV5_FP8 = 4, EPI_ASSIGN_FP8 = 5 };

struct GemmParams {
    int M, N, K;              // per batch; K = contraction length handled by one split
    int k_batch_rows;         // MN-major operands: K-row offset between batches in the 2-D tensor maps (0 if unbatched)
    int splitk;               // grid.z = batch * splitk
    // epilogue
    void* C;                  // STORE_F32: float [M, ldc] (+ bias, relu); CONV5_BF16: bf16 [M, ldc]; ASSIGN: bf16 S' [M, 64]
    long long c_batch, c_slab;  // element offsets per batch / per split-K slab
    int ldc;
    const float* bias;        // [N] or nullptr
    int relu;
    float* aux;               // CONV5_BF16: rowss [M, gridDim.y]; COLMAX: g [clouds, N] (zero-initialised);
                              // ASSIGN: a_part [M/128, 64]
    const float* rowss;       // ASSIGN: [M, rowss_parts] partial sums of squares of the rows of H
    int rowss_parts;
    const float* bn_scale;    // ASSIGN: cluster_bn affine [64]
    const float* bn_shift;
    const float* cloud_absmax;  // CONV5_FP8: [clouds] max |x| over the cloud's conv5 input rows (bounds the output range)
    float l1max, bmax;          // CONV5_FP8: max_f sum_c |W5[c,f]|, max_f |b5[f]|   (|H[r,f]| <= absmax * l1max + bmax)
    const float* sscale;        // ASSIGN_FP8: [clouds] power-of-two scale of the cloud's S' (fp8 range), see head_fp8.cu
    int rows_per_cloud;       // COLMAX / CONV5_FP8 / ASSIGN_FP8: N points per cloud
    const int* cloud_mask;    // persistent kernels: if set, only row tiles of clouds (rows_per_cloud rows each) with a non-zero entry
    long long* timeline;      // debug: CTA 0 records clock64 per tile: [tile][0] MMA start [1] MMA issued [2] epilogue start [3] epilogue end
    int l2_prefetch_tiles;    // persistent kernels: ask the L2 for the A tile this many of the CTA's tiles ahead (0 = off)
    int reverse_m;            // persistent kernels: walk the row tiles from the last to the first (the producer kernel wrote
                              // the last tiles most recently: they are the ones still in L2)
};

// Extra by-value kernel arguments of the persistent kernel (constant bank, no LSU traffic): the output tensor map of epilogues
// that leave through TMA stores, and the bias vector read as instruction operands / LDC instead of shared-memory broadcasts
// (a warp-wide LDS.128 of one address still costs 4 wavefronts of the L1 data pipe: 8 per 32 columns was a third of that pipe).
struct EpiExtra {
    CUtensorMap tmC;
    float bias[1024];
};

template <typename T> struct ElemTraits;
template <> struct ElemTraits<float> {
    static constexpr int PER128 = 32, UMMA_K = 8, FMT = 2;
    static constexpr bool F16 = false;
};
template <> struct ElemTraits<__half> {
    static constexpr int PER128 = 64, UMMA_K = 16, FMT = 0;
    static constexpr bool F16 = true;
};
template <> struct ElemTraits<__nv_bfloat16> {
    static constexpr int PER128 = 64, UMMA_K = 16, FMT = 1;
    static constexpr bool F16 = true;
};

constexpr int TC_BM = 128;

template <typename T, int BN>
struct TileCfg {
    static constexpr int BK = ElemTraits<T>::PER128;                      // k elements (or k rows) per pipeline stage
    static constexpr uint32_t A_BYTES = TC_BM * 128;                      // 16 KB
    static constexpr uint32_t B_BYTES = BN * 128;
    static constexpr int STAGES_MAX = (int)((200 * 1024) / (A_BYTES + B_BYTES));
    static constexpr int STAGES = STAGES_MAX > 8 ? 8 : STAGES_MAX;
    static constexpr size_t smem(int stages) { return (size_t)stages * (A_BYTES + B_BYTES) + 1024 + 256; }
};

// Four floats -> four e4m3 bytes (x0 in the lowest byte) with STOCHASTIC rounding (cvt.rs, sm_100a): unbiased, so the
// quantisation errors of different points average out in the sums over points / features even when the points are identical
// (an all-zero or duplicated cloud quantised round-to-nearest repeats the SAME error 4096 times: measured 1.2e-3 on the
// descriptor, over the tolerance).  The random bits are a hash of (row within the cloud, column): deterministic, and independent
// of where in a batch the cloud sits.
__device__ __forceinline__ uint32_t hash_bits(uint32_t row, uint32_t col) {
    uint32_t h = row * 0x9E3779B1u + col * 0x85EBCA77u + 0x165667B1u;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    h *= 0x297A2D39u;
    h ^= h >> 15;
    return h;
}
__device__ __forceinline__ uint32_t f32x4_to_e4m3_sr(float x0, float x1, float x2, float x3, uint32_t rbits) {
    uint32_t d;
    asm("cvt.rs.satfinite.e4m3x4.f32 %0, {%1, %2, %3, %4}, %5;" : "=r"(d) : "f"(x3), "f"(x2), "f"(x1), "f"(x0), "r"(rbits));
    return d;
}

// packed fp32 pairs (FADD2 / FMUL2 / FFMA2 of sm_100): one issue slot for two lanes of epilogue arithmetic
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// epilogues (shared by the one-tile-per-CTA kernel and the persistent B-resident kernel)
// ---------------------------------------------------------------------------------------------------------
struct EpiCtx {
    uint32_t trow;      // TMEM address of this thread's accumulator row (lane quarter + column base)
    int m, row, lane;   // global row, row within the tile, lane
    int m0, n0;         // tile origin
    int mtile;          // index of the 128-row tile (ASSIGN: a_part row)
    int nparts, npart;  // CONV5: number / index of the N tile (rowss partials)
    long long c_off;    // STORE_F32: element offset of this batch / split-K slab in C
    float* scratch;     // ASSIGN: [4][64] fp32 column-sum partials of the four epilogue warps
    int epi_tid;        // 0..127 within the epilogue warps
    const float* bias;  // bias of this N tile (global or shared), indexed by the column within the tile; may be nullptr
    int col_begin, col_end;   // columns of the tile this warp drains (two warps per lane quarter split the tile)
    int warp_slot;            // CONV5: index of this epilogue warp (private 4 KB staging slot in scratch)
};

// conv5 -> fp8 epilogue of the persistent kernel (models/epc-net.py:136-139 with the output format of head_fp8.cu).
// H' = 2^e relu(acc + b) stored as fp8 e4m3, e per CLOUD from a bound of the cloud's |H| (absmax(x) * l1max + bmax <= 2^E  =>
// e = 8 - E, so |H'| <= 256 < 448 always: no overflow, no fallback); the per-row sum of squares is taken from the SCALED fp32
// values, so every consumer sees H'/|H'| = H/|H| -- the power of two cancels exactly in the row normalisation of
// models/epc-net.py:147-148.  A thread owns 128 columns = one full 128-byte line of its row.
// What bounds it is the L1 data pipe, not arithmetic (ncu: 69 % busy when the bias came from shared memory and every thread
// stored its own 32-byte sectors -- a 256-bit store of 32 lanes to 32 different lines is 64 wavefronts), so:
//   * the bias is read from the constant bank (EpiExtra), packed FADD2 / FMUL2 / FFMA2 do the arithmetic in half the issue slots,
//   * one avalanche hash per 32 columns, the eight rounding words of the chunk are one IMAD each (word k = h * M_k + g),
//   * the warp's 32 x 128-byte block goes to a private, 128B-swizzled 4 KB of shared memory (8 conflict-free st.shared.v4 per
//     thread) and leaves as ONE TMA store per warp and tile.
template <int BN>
__device__ __forceinline__ void epilogue_conv5_fp8(const GemmParams& p, const EpiCtx& c, const EpiExtra& ex, int warp_first_row) {
    const uint32_t trow = c.trow;
    const int m = c.m, lane = c.lane, n0 = c.n0;
    int e;
    frexpf(fmaf(__ldg(p.cloud_absmax + c.m0 / p.rows_per_cloud), p.l1max, p.bmax) + 1e-30f, &e);
    const float scale = ldexpf(1.0f, 8 - e);
    const uint64_t scale2 = pack2(scale, scale);
    const uint32_t row_in_cloud = (uint32_t)(m % p.rows_per_cloud);
    const uint32_t slot = smem_u32(c.scratch) + (uint32_t)c.warp_slot * 4096u;
    const uint32_t my_row = slot + (uint32_t)lane * 128u;
    const float* bias = ex.bias + n0 + c.col_begin;
    if (lane == 0) bulk_wait_read0();                     // the previous tile's store has read the slot
    __syncwarp();
    uint64_t ssA = 0ull, ssB = 0ull;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        float v[32];
        tmem_ld32(trow + (uint32_t)(c.col_begin + 32 * h), v);
        const uint32_t hh = hash_bits(row_in_cloud, (uint32_t)(n0 + c.col_begin + 32 * h));
        const uint32_t g = (hh ^ (hh >> 13)) * 0x9E3779B1u;
        uint32_t pk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            constexpr uint32_t MK[8] = {0x9E3779B1u, 0x85EBCA77u, 0xC2B2AE3Du, 0x27D4EB2Fu, 0x165667B1u, 0x2C1B3C6Du, 0x297A2D39u, 0xD3A2646Du};
            const int j = 32 * h + 4 * k;
            float a0, a1, a2, a3;
            unpack2(add2(pack2(v[4 * k], v[4 * k + 1]), pack2(bias[j], bias[j + 1])), a0, a1);
            unpack2(add2(pack2(v[4 * k + 2], v[4 * k + 3]), pack2(bias[j + 2], bias[j + 3])), a2, a3);
            const uint64_t x01 = mul2(pack2(fmaxf(a0, 0.f), fmaxf(a1, 0.f)), scale2);
            const uint64_t x23 = mul2(pack2(fmaxf(a2, 0.f), fmaxf(a3, 0.f)), scale2);
            ssA = fma2(x01, x01, ssA);
            ssB = fma2(x23, x23, ssB);
            float x0, x1, x2, x3;
            unpack2(x01, x0, x1);
            unpack2(x23, x2, x3);
            pk[k] = f32x4_to_e4m3_sr(x0, x1, x2, x3, hh * MK[k] + g);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int ch = 2 * h + half;                 // 16-byte chunk of the 128-byte row
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + (uint32_t)((ch ^ (lane & 7)) << 4)), "r"(pk[4 * half]),
                         "r"(pk[4 * half + 1]), "r"(pk[4 * half + 2]), "r"(pk[4 * half + 3]) : "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the TMA engine
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(&ex.tmC, slot, n0 + c.col_begin, warp_first_row);       // rows past M are clipped by the tensor map
    }
    float s0, s1, s2, s3;
    unpack2(ssA, s0, s1);
    unpack2(ssB, s2, s3);
    if (m < p.M) p.aux[(size_t)m * c.nparts + c.npart] = (s0 + s1) + (s2 + s3);
}

template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const EpiCtx& c) {
    const uint32_t trow = c.trow;
    const int m = c.m, row = c.row, lane = c.lane, n0 = c.n0;
    (void)row; (void)lane; (void)n0;
        if (EPI == EPI_STORE_F32) {
            float* C = reinterpret_cast<float*>(p.C) + c.c_off;
    #pragma unroll 1
            for (int c0 = c.col_begin; c0 < c.col_end; c0 += 32) {
                float v[32];
                tmem_ld32(trow + (uint32_t)c0, v);
                if (m < p.M) {
                    float* dst = C + (size_t)m * p.ldc + n0 + c0;
    #pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (p.bias) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + i));
                            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                        }
                        if (p.relu) {
                            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                        }
                        *reinterpret_cast<float4*>(dst + i) = o;
                    }
                }
            }
        } else if (EPI == EPI_CONV5_BF16) {
            // H = relu(acc + b) stored as bf16; per-row sum of squares of the fp32 values for the later L2 norm.
            // Thread = row, but a warp store of 32 x 16 B to 32 different rows only half-fills its 32 B sectors; so each warp
            // stages its 32 rows x 64 columns in a private, swizzled 4 KB of shared memory and writes full 128 B rows
            // (8 lanes x 16 B per row, 4 rows per instruction).
            __nv_bfloat16* H = reinterpret_cast<__nv_bfloat16*>(p.C);
            const uint32_t stage = smem_u32(c.scratch) + (uint32_t)c.warp_slot * 4096u;
            const int r_in = lane;                               // this thread's row within the warp's 32
            const int m_warp = m - lane;                         // first row of the warp
            float ss = 0.f;
    #pragma unroll 1
            for (int c0 = c.col_begin; c0 < c.col_end; c0 += 64) {
    #pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float v[32];
                    tmem_ld32(trow + (uint32_t)(c0 + 32 * h), v);
    #pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t pk[4];
    #pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int i = 8 * q + 2 * e;
                            const float2 bb = *reinterpret_cast<const float2*>(c.bias + c0 + 32 * h + i);
                            const float x = fmaxf(v[i] + bb.x, 0.f), y = fmaxf(v[i + 1] + bb.y, 0.f);
                            ss = fmaf(x, x, ss);
                            ss = fmaf(y, y, ss);
                            const __nv_bfloat162 hh = __floats2bfloat162_rn(x, y);
                            pk[e] = *reinterpret_cast<const uint32_t*>(&hh);
                        }
                        const int ch = 4 * h + q;                // 16-byte chunk of the 128-byte row segment
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)r_in * 128u + (uint32_t)((ch ^ (r_in & 7)) << 4)),
                                     "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                    }
                }
                __syncwarp();
    #pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int idx = lane + 32 * i, r = idx >> 3, ch = idx & 7;
                    uint4 val;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                                 : "r"(stage + (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4)) : "memory");
                    if (m_warp + r < p.M) *reinterpret_cast<uint4*>(H + (size_t)(m_warp + r) * p.ldc + n0 + c0 + 8 * ch) = val;
                }
                __syncwarp();
            }
            if (m < p.M) p.aux[(size_t)m * c.nparts + c.npart] = ss;
        } else if (EPI == EPI_COLMAX) {
            // max over the tile's rows of relu(acc + b): values >= 0, so unsigned-int order == float order
            const int cloud = c.m0 / p.rows_per_cloud;
            int* g = reinterpret_cast<int*>(p.aux) + (size_t)cloud * p.N;
    #pragma unroll 1
            for (int c0 = c.col_begin; c0 < c.col_end; c0 += 32) {
                float v[32];
                tmem_ld32(trow + (uint32_t)c0, v);
                if (m >= p.M) {                                  // rows past the end (never with N % 128 == 0): relu(-inf + b) = 0
    #pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = -INFINITY;
                }
                // Column maxima of acc + b as SIGNED integers: whenever the true maximum is >= 0 the integer order is the float
                // order, and when every value is negative the (meaningless) negative result is clamped to relu's 0 below -- so
                // the per-element relu (FMNMX) is not needed.  Lane i then picks column i's maximum with a 5-level select tree
                // on the bits of its lane index (31 SEL) instead of 32 compare + predicated-move pairs.
                const uint32_t bias_s = smem_u32(c.bias) + (uint32_t)c0 * 4u;
                int r[32];
    #pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    uint64_t b01, b23;
                    asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(b01), "=l"(b23) : "r"(bias_s + (uint32_t)i4 * 16u));      // same address in every lane
                    float t0, t1, t2, t3;
                    unpack2(add2(pack2(v[4 * i4], v[4 * i4 + 1]), b01), t0, t1);
                    unpack2(add2(pack2(v[4 * i4 + 2], v[4 * i4 + 3]), b23), t2, t3);
                    r[4 * i4] = __reduce_max_sync(FULL, __float_as_int(t0));
                    r[4 * i4 + 1] = __reduce_max_sync(FULL, __float_as_int(t1));
                    r[4 * i4 + 2] = __reduce_max_sync(FULL, __float_as_int(t2));
                    r[4 * i4 + 3] = __reduce_max_sync(FULL, __float_as_int(t3));
                }
    #pragma unroll
                for (int lvl = 0; lvl < 5; ++lvl) {
                    const bool bit = (lane >> lvl) & 1;
    #pragma unroll
                    for (int j = 0; j < (16 >> lvl); ++j) r[j] = bit ? r[2 * j + 1] : r[2 * j];
                }
                const int mine = max(r[0], 0);
                atomicMax(g + n0 + c0 + lane, (int)mine);
            }
        } else if (EPI == EPI_ASSIGN || EPI == EPI_ASSIGN_FP8) {
            // BN == 64: this thread owns all 64 cluster logits of its point (loupe.py:255-276)
            float v[64];
            {
                float t[32];
                tmem_ld32(trow, t);
    #pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = t[i];
                tmem_ld32(trow + 32u, t);
    #pragma unroll
                for (int i = 0; i < 32; ++i) v[32 + i] = t[i];
            }
            float ssq = 0.f;
            if (m < p.M) {
                if (p.rowss_parts == 8) {            // conv5's partials: one 32-byte row, two vector loads (same summation order)
                    const float4 a = __ldg(reinterpret_cast<const float4*>(p.rowss + (size_t)m * 8));
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.rowss + (size_t)m * 8) + 1);
                    ssq = ((((((a.x + a.y) + a.z) + a.w) + b.x) + b.y) + b.z) + b.w;
                } else {
                    for (int i = 0; i < p.rowss_parts; ++i) ssq += p.rowss[(size_t)m * p.rowss_parts + i];
                }
            }
            const float inv = 1.0f / sqrtf(fmaxf(ssq, L2_EPS));
            float mx = -INFINITY;
    #pragma unroll
            for (int i = 0; i < 64; ++i) {
                v[i] = (v[i] * inv) * __ldg(p.bn_scale + i) + __ldg(p.bn_shift + i);
                mx = fmaxf(mx, v[i]);
            }
            float den = 0.f;
    #pragma unroll
            for (int i = 0; i < 64; ++i) {
                v[i] = expf(v[i] - mx);
                den += v[i];
            }
            const float rden = 1.0f / den;
    #pragma unroll
            for (int i = 0; i < 64; ++i) v[i] *= rden;
            if (EPI == EPI_ASSIGN_FP8) {
                // S'' = t softmax / |H'| as fp8 e4m3 in a 128-byte row (64 values + 64 bytes of padding that the VLAD MMA reads as
                // unused accumulator columns); t = the cloud's power-of-two scale (undone exactly when V is finalised)
                if (m < p.M) {
                    const float ts = inv * __ldg(p.sscale + c.m0 / p.rows_per_cloud);
                    const uint32_t row_in_cloud = (uint32_t)(m % p.rows_per_cloud);
                    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.C) + (size_t)m * 128);
    #pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t pk[4];
                        uint32_t rb = hash_bits(row_in_cloud, 4096u + (uint32_t)(16 * i));
    #pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            rb = (rb ^ (rb >> 13)) * 0x9E3779B1u;
                            pk[j] = f32x4_to_e4m3_sr(v[16 * i + 4 * j] * ts, v[16 * i + 4 * j + 1] * ts, v[16 * i + 4 * j + 2] * ts,
                                                     v[16 * i + 4 * j + 3] * ts, rb);
                        }
                        dst[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                } else {
    #pragma unroll
                    for (int i = 0; i < 64; ++i) v[i] = 0.f;
                }
            } else if (m < p.M) {
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)m * 64);
    #pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint32_t pk[4];
    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * i + 2 * j] * inv, v[8 * i + 2 * j + 1] * inv);
                        pk[j] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    dst[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            } else {
    #pragma unroll
                for (int i = 0; i < 64; ++i) v[i] = 0.f;
            }
            // column sums of the soft assignment over this tile (a_sum partials, loupe.py:276): transpose-reduce across the
            // warp's 32 rows -- each step a lane hands half of its columns to its partner and adds the partner's other half
            // (62 shuffles for 64 columns) -- then the four warps' partials meet in 1 KB of shared memory.  Fixed order.
    #pragma unroll
            for (int h = 32, bit = 16; h >= 2; h >>= 1, bit >>= 1) {
                const bool up = (c.lane & bit) != 0;
    #pragma unroll
                for (int i = 0; i < h; ++i) {
                    const float keep = up ? v[i + h] : v[i];
                    const float send = up ? v[i] : v[i + h];
                    v[i] = keep + __shfl_xor_sync(FULL, send, bit);
                }
            }
            // lane l now holds the warp's sums of columns 2l and 2l+1 (lane bit 16 chose the upper 32 columns, bit 8 the upper 16, ...)
            const int col0 = 2 * c.lane;      // 32*b16 + 16*b8 + 8*b4 + 4*b2 + 2*b1
            float* sS = c.scratch;        // [4 warps][64]
            const int wq = c.row >> 5;
            sS[wq * 64 + col0] = v[0];
            sS[wq * 64 + col0 + 1] = v[1];
            asm volatile("bar.sync 1, 128;" ::: "memory");     // the four epilogue warps only
            const int t = c.epi_tid;
            if (t < 64) p.aux[(size_t)c.mtile * 64 + t] = ((sS[t] + sS[64 + t]) + sS[128 + t]) + sS[192 + t];
            asm volatile("bar.sync 1, 128;" ::: "memory");     // scratch may be rewritten by the next tile (persistent kernel)
        }
}

template <typename T, int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(192)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p,
               const int stages) {
    using Tr = ElemTraits<T>;
    using Cfg = TileCfg<T, BN>;
    constexpr int BK = Cfg::BK;
    constexpr int CH = Tr::PER128;                 // M/N elements per 128-byte chunk of an MN-major operand
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base;
    uint8_t* sB = base + (size_t)stages * Cfg::A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)stages * (Cfg::A_BYTES + Cfg::B_BYTES));
    uint64_t* empty = full + stages;
    uint64_t* acc_full = empty + stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
    const int batch = blockIdx.z / p.splitk, split = blockIdx.z % p.splitk;
    const int nkb = p.K / BK;                       // k-blocks handled by this CTA
    const int krow0 = batch * p.k_batch_rows + split * p.K;   // first K coordinate (MN-major operands / split-K)
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % stages;
                const uint32_t ph = (kb / stages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], Cfg::A_BYTES + Cfg::B_BYTES);
                uint8_t* a = sA + (size_t)s * Cfg::A_BYTES;
                uint8_t* b = sB + (size_t)s * Cfg::B_BYTES;
                const int k0 = krow0 + kb * BK;
                if (!A_MN) {
                    tma_load_2d(a, &tmA, &full[s], k0, m0);                    // box {BK, 128}
                } else {
#pragma unroll
                    for (int c = 0; c < TC_BM / CH; ++c)                        // boxes {CH, BK}: one 128 B chunk of M each
                        tma_load_2d(a + (size_t)c * BK * 128, &tmA, &full[s], m0 + c * CH, k0);
                }
                if (!B_MN) {
                    tma_load_2d(b, &tmB, &full[s], k0, n0);                    // box {BK, BN}
                } else {
#pragma unroll
                    for (int c = 0; c < BN / CH; ++c)
                        tma_load_2d(b + (size_t)c * BK * 128, &tmB, &full[s], n0 + c * CH, k0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(Tr::FMT, TC_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            // per-MMA K advance inside a stage: K-major: UMMA_K elements = 32 B along the swizzled row;
            //                                   MN-major: UMMA_K rows of 128 B
            constexpr uint32_t a_step = A_MN ? Tr::UMMA_K * 128 : 32;
            constexpr uint32_t b_step = B_MN ? Tr::UMMA_K * 128 : 32;
            constexpr uint32_t a_lbo = A_MN ? BK * 128 : 16, b_lbo = B_MN ? BK * 128 : 16;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % stages;
                const uint32_t ph = (kb / stages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sA + (size_t)s * Cfg::A_BYTES);
                const uint32_t b_addr = smem_u32(sB + (size_t)s * Cfg::B_BYTES);
#pragma unroll
                for (int kk = 0; kk < BK / Tr::UMMA_K; ++kk) {
                    const uint64_t da = smem_desc_sw128(a_addr + kk * a_step, a_lbo, 1024);
                    const uint64_t db = smem_desc_sw128(b_addr + kk * b_step, b_lbo, 1024);
                    mma_ss<Tr::F16>(tmem_d, da, db, idesc, (kb | kk) != 0);
                }
                mma_commit(&empty[s]);          // frees the smem stage when these MMAs have read it
            }
            mma_commit(acc_full);               // accumulator complete
        }
    } else {
        // ---- epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31 ------------------------------------------------
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int m = m0 + row;
        const uint32_t trow = tmem_d + ((uint32_t)(q * 32) << 16);
        mbar_wait(acc_full, 0);
        tc_fence_after();
        EpiCtx c;
        c.trow = trow; c.m = m; c.row = row; c.lane = lane; c.m0 = m0; c.n0 = n0; c.mtile = blockIdx.x;
        c.nparts = gridDim.y; c.npart = blockIdx.y;
        c.c_off = (long long)batch * p.c_batch + (long long)split * p.c_slab;
        c.scratch = reinterpret_cast<float*>(sA); c.epi_tid = threadIdx.x - 64;
        c.bias = p.bias ? p.bias + n0 : nullptr; c.col_begin = 0; c.col_end = BN; c.warp_slot = warp - 2;
        epilogue_tile<BN, EPI>(p, c);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------
// Persistent, B-resident variant (both operands K-major).  The one-tile-per-CTA kernel re-fetches its B tile for
// every 128-row tile, which makes conv5 (K = 256) and the assignment GEMM (N = 64) L2-bandwidth bound.  Here a CTA
// owns one N tile, loads its whole [BN x K] slice of B once, and then streams 128-row A tiles through a ring:
//   warp 0  TMA producer        warp 1  MMA issuer (accumulators double-buffered in TMEM: 2 x BN columns)
//   warps 2-5  epilogue of tile i while the tensor core already works on tile i+1
// grid.x = (#CTAs, a multiple of N/BN); CTA c: N tile c % (N/BN), row tiles c / (N/BN), + gridDim.x / (N/BN), ...
// ---------------------------------------------------------------------------------------------------------
template <typename T, int BN, int EPI, int EW /*epilogue warps: 4, or 8 = two per TMEM lane quarter*/,
          int CL = 1 /*cluster size: CL > 1 = the N/BN == CL CTAs of a cluster share every A tile by TMA multicast*/>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
tc_gemm_bres_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p,
                    const int a_stages, const __grid_constant__ EpiExtra ex) {
    using Tr = ElemTraits<T>;
    constexpr int BK = Tr::PER128;
    constexpr uint32_t A_BYTES = TC_BM * 128, B_BYTES = BN * 128;
    constexpr size_t SCRATCH = (EPI == EPI_ASSIGN) ? (size_t)4 * 64 * 4 : (EPI == EPI_CONV5_BF16 || EPI == EPI_CONV5_FP8) ? (size_t)EW * 4096 : 0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int nkb = p.K / BK;
    uint8_t* sB = base;                                              // [nkb][BN x 128 B]   resident
    uint8_t* sA = sB + (size_t)nkb * B_BYTES;                        // [a_stages][128 x 128 B] ring
    float* scratch = reinterpret_cast<float*>(sA + (size_t)a_stages * A_BYTES);
    float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + SCRATCH);     // [BN] bias of this N tile
    uint64_t* full = reinterpret_cast<uint64_t*>(sBias + BN);
    uint64_t* empty = full + a_stages;
    uint64_t* b_full = empty + a_stages;
    uint64_t* tfull = b_full + 1;          // [2] accumulator ready
    uint64_t* tempty = tfull + 2;          // [2] accumulator drained (4 arrivals: one per epilogue warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NT = p.N / BN;
    // CL > 1: cluster = the NT (== CL) N tiles of one M-tile stream; rank in cluster = N tile
    const int n_tile = (CL > 1) ? (int)cluster_ctarank() : (int)(blockIdx.x % NT);
    const int cta_m = blockIdx.x / NT, m_stride = gridDim.x / NT;
    const int n0 = n_tile * BN;
    constexpr uint16_t CL_MASK = (uint16_t)((1u << CL) - 1u);
    const int num_m_tiles = (p.M + TC_BM - 1) / TC_BM;
    auto tile_of = [&](int mt) { return p.reverse_m ? num_m_tiles - 1 - mt : mt; };
    auto tile_on = [&](int mt) { return !p.cloud_mask || p.cloud_mask[(tile_of(mt) * TC_BM) / p.rows_per_cloud] != 0; };
    constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < a_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CL);          // every CTA of the cluster frees the slot (multicast commit)
        }
        mbar_init(b_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], EW);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    if (p.bias && EPI != EPI_CONV5_FP8)
        for (int i = threadIdx.x; i < BN; i += blockDim.x) sBias[i] = p.bias[n0 + i];
    if (EPI == EPI_CONV5_FP8 && warp == 0 && lane == 0) tma_prefetch_desc(&ex.tmC);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync();                // peers' barriers are initialised before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(b_full, (uint32_t)nkb * B_BYTES);
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sB + (size_t)kb * B_BYTES, &tmB, b_full, kb * BK, n0);
            int it = 0;
            for (int mt = cta_m; mt < num_m_tiles; mt += m_stride) {
                if (!tile_on(mt)) continue;
                if (p.l2_prefetch_tiles > 0 && n_tile == (mt / m_stride) % NT) {     // one of the NT CTAs that share the A tile asks the L2 for it early
                    const int ahead = mt + p.l2_prefetch_tiles * m_stride;
                    if (ahead < num_m_tiles)
                        for (int kb = 0; kb < nkb; ++kb) tma_prefetch_l2_2d(&tmA, kb * BK, tile_of(ahead) * TC_BM);
                }
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % a_stages;
                    const uint32_t ph = (it / a_stages) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], A_BYTES);
                    if (CL == 1)
                        tma_load_2d(sA + (size_t)s * A_BYTES, &tmA, &full[s], kb * BK, tile_of(mt) * TC_BM);
                    else if (kb % CL == n_tile)    // this CTA's share of the tile, delivered to the whole cluster
                        tma_load_2d_mc(sA + (size_t)s * A_BYTES, &tmA, &full[s], kb * BK, tile_of(mt) * TC_BM, CL_MASK);
                }
            }
        }
    } else if (warp == 1) {
        {                                      // all 32 lanes, converged; one elected lane issues (see elect_one)
            constexpr uint32_t idesc = make_idesc(Tr::FMT, TC_BM, BN, 0, 0);
            mbar_wait(b_full, 0);
            const uint32_t a_lo0 = desc_lo_sw128(smem_u32(sA)), b_lo0 = desc_lo_sw128(smem_u32(sB));
            int s = 0, tile = 0;
            uint32_t ph = 0;
            for (int mt = cta_m; mt < num_m_tiles; mt += m_stride) {
                if (!tile_on(mt)) continue;
                const int buf = tile & 1;
                mbar_wait(&tempty[buf], ((tile >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator
                tc_fence_after();
                if (p.timeline && blockIdx.x == 0 && tile < 64 && lane == 0) p.timeline[tile * 4 + 0] = clock64();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + (uint32_t)s * (A_BYTES >> 4), b_lo = b_lo0 + (uint32_t)kb * (B_BYTES >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < BK / Tr::UMMA_K; ++kk)
                            mma_ss<Tr::F16>(tmem_d, desc_of(a_lo + 2 * kk), desc_of(b_lo + 2 * kk), idesc, (kb | kk) != 0);
                        if (CL == 1)
                            mma_commit(&empty[s]);
                        else
                            mma_commit_mc(&empty[s], CL_MASK);
                    }
                    __syncwarp();
                    if (++s == a_stages) { s = 0; ph ^= 1; }
                }
                if (elect_one()) mma_commit(&tfull[buf]);
                __syncwarp();
                if (p.timeline && blockIdx.x == 0 && tile < 64 && lane == 0) p.timeline[tile * 4 + 1] = clock64();
                ++tile;
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int tile = 0;
        for (int mt = cta_m; mt < num_m_tiles; mt += m_stride) {
            if (!tile_on(mt)) continue;
            const int buf = tile & 1;
            mbar_wait(&tfull[buf], (tile >> 1) & 1);
            tc_fence_after();
            if (p.timeline && blockIdx.x == 0 && tile < 64 && warp == 2 && lane == 0) p.timeline[tile * 4 + 2] = clock64();
            EpiCtx c;
            c.trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
            c.m0 = tile_of(mt) * TC_BM; c.m = c.m0 + row; c.row = row; c.lane = lane; c.n0 = n0; c.mtile = tile_of(mt);
            c.c_off = 0; c.scratch = scratch; c.epi_tid = threadIdx.x - 64;
            c.bias = p.bias ? sBias : nullptr;
            constexpr int SPLIT = EW / 4;                       // warps per lane quarter
            const int part = (warp - 2) >> 2;                   // which column share this warp drains
            c.col_begin = part * (BN / SPLIT); c.col_end = c.col_begin + BN / SPLIT;
            c.nparts = NT * SPLIT; c.npart = n_tile * SPLIT + part; c.warp_slot = warp - 2;
            if (EPI == EPI_CONV5_FP8)
                epilogue_conv5_fp8<BN>(p, c, ex, c.m0 + q * 32);
            else
                epilogue_tile<BN, EPI>(p, c);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
            if (p.timeline && blockIdx.x == 0 && tile < 64 && warp == 2 && lane == 0) p.timeline[tile * 4 + 3] = clock64();
            ++tile;
        }
        if (EPI == EPI_CONV5_FP8 && lane == 0) bulk_wait0();      // this warp's TMA stores are done
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync();                // no CTA leaves while a peer may still multicast into it
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D tensor [rows, cols] (cols contiguous, row pitch `ld` elements); box = {box_cols, box_rows}, 128B swizzle
template <typename T>
inline int make_tmap_2d(CUtensorMap* tm, const T* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                        uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return EPC_ECUDA;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * sizeof(T)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    CUresult r = enc(tm, dt, 2, const_cast<T*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u ptr=%p", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows, ptr);
        return EPC_ECUDA;
    }
    return EPC_OK;
}

// Operand description: a 2-D row-major array [rows, cols] with pitch ld.
//   K-major operand  : rows = M (or N), cols = K.      MN-major operand: rows = K (all batches stacked), cols = M (or N).
template <typename T>
struct Operand {
    const T* ptr;
    long long rows, cols, ld;
};

template <typename T, int BN, bool A_MN, bool B_MN, int EPI>
inline int tc_gemm_launch(const Operand<T>& A, const Operand<T>& B, const tc::GemmParams& p, int batch, cudaStream_t st,
                          int ctas_per_sm = 1) {
    using Cfg = tc::TileCfg<T, BN>;
    constexpr int PER128 = tc::ElemTraits<T>::PER128;
    EPC_CHECK_ARG(p.K % Cfg::BK == 0 && p.K >= Cfg::BK, "tc_gemm: K=%d must be a multiple of %d", p.K, Cfg::BK);
    EPC_CHECK_ARG(p.N % BN == 0, "tc_gemm: N=%d must be a multiple of %d", p.N, BN);
    EPC_CHECK_ARG((reinterpret_cast<uintptr_t>(A.ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(B.ptr) & 15) == 0 &&
                      (A.ld * sizeof(T)) % 16 == 0 && (B.ld * sizeof(T)) % 16 == 0,
                  "tc_gemm: operands must be 16-byte aligned with 16-byte pitches");
    if (p.M == 0 || batch == 0) return EPC_OK;
    CUtensorMap tmA, tmB;
    if (int rc = A_MN ? make_tmap_2d(&tmA, A.ptr, A.rows, A.cols, A.ld, PER128, Cfg::BK)
                      : make_tmap_2d(&tmA, A.ptr, A.rows, A.cols, A.ld, Cfg::BK, tc::TC_BM))
        return rc;
    if (int rc = B_MN ? make_tmap_2d(&tmB, B.ptr, B.rows, B.cols, B.ld, PER128, Cfg::BK)
                      : make_tmap_2d(&tmB, B.ptr, B.rows, B.cols, B.ld, Cfg::BK, BN))
        return rc;
    int stages = Cfg::STAGES;
    if (ctas_per_sm > 1) {           // leave room for a second resident CTA (its epilogue overlaps our main loop)
        const int fit = (int)((110 * 1024 - 1280) / (Cfg::A_BYTES + Cfg::B_BYTES));
        stages = fit < 2 ? 2 : (fit < stages ? fit : stages);
    }
    const int nkb = p.K / Cfg::BK;
    if (stages > nkb) stages = nkb < 2 ? 2 : nkb;
    const size_t smem = Cfg::smem(stages);
    auto kern = tc::tc_gemm_kernel<T, BN, A_MN, B_MN, EPI>;
    static PerDeviceSize attr_smem;
    EPC_CUDA(ensure_dyn_smem(kern, Cfg::smem(Cfg::STAGES), attr_smem));
    dim3 grid((p.M + tc::TC_BM - 1) / tc::TC_BM, p.N / BN, batch * p.splitk);
    kern<<<grid, 192, smem, st>>>(tmA, tmB, p, stages);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}


inline int sm_count() {
    static PerDeviceSize cache;
    std::atomic<size_t>& c = cache.v[current_device_slot()];
    size_t n = c.load(std::memory_order_acquire);
    if (!n) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        n = v > 0 ? (size_t)v : 148;
        c.store(n, std::memory_order_release);
    }
    return (int)n;
}

// Number of CTAs a persistent kernel launches: one per SM, or fewer when `env` says so (leaves SMs to kernels of other
// streams: Engine EMBED_STREAMS overlaps the ALU-bound kNN of one chunk with the HBM-bound head of another).
inline int persistent_ctas(const char* env) {
    const char* e = getenv(env);
    const int n = e ? atoi(e) : 0;
    return (n > 0 && n < sm_count()) ? n : sm_count();
}

// Persistent B-resident launch: A [M,K], B [N,K] both K-major; one CTA per SM (rounded to a multiple of N/BN).
template <typename T, int BN, int EPI, int EW = 4, int CL = 1>
inline int tc_gemm_bres_launch(const Operand<T>& A, const Operand<T>& B, const tc::GemmParams& p, cudaStream_t st, const tc::EpiExtra* extra = nullptr) {
    constexpr int BK = tc::ElemTraits<T>::PER128;
    constexpr size_t A_BYTES = tc::TC_BM * 128, B_BYTES = BN * 128;
    constexpr size_t SCRATCH = (EPI == tc::EPI_ASSIGN) ? (size_t)4 * 64 * 4 : (EPI == tc::EPI_CONV5_BF16 || EPI == tc::EPI_CONV5_FP8) ? (size_t)EW * 4096 : 0;
    static const tc::EpiExtra no_extra = {};
    const tc::EpiExtra& ex = extra ? *extra : no_extra;
    EPC_CHECK_ARG(EPI != tc::EPI_CONV5_FP8 || extra, "tc_gemm_bres: the fp8 conv5 epilogue needs its output tensor map and bias%s", "");
    EPC_CHECK_ARG(p.K % BK == 0 && p.K >= BK && p.N % BN == 0 && p.splitk == 1, "tc_gemm_bres: bad shape K=%d N=%d", p.K, p.N);
    EPC_CHECK_ARG((reinterpret_cast<uintptr_t>(A.ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(B.ptr) & 15) == 0 &&
                      (A.ld * sizeof(T)) % 16 == 0 && (B.ld * sizeof(T)) % 16 == 0,
                  "tc_gemm_bres: operands must be 16-byte aligned with 16-byte pitches");
    if (p.M == 0) return EPC_OK;
    const int nkb = p.K / BK;
    const size_t fixed = 1024 + (size_t)nkb * B_BYTES + SCRATCH + BN * 4 + 512;
    int a_stages = (int)((227 * 1024 - fixed) / A_BYTES);
    EPC_CHECK_ARG(a_stages >= 2, "tc_gemm_bres: B slice of %zu bytes leaves no room for the A ring", (size_t)nkb * B_BYTES);
    if (a_stages > 8) a_stages = 8;
    const size_t smem = fixed + (size_t)a_stages * A_BYTES;
    CUtensorMap tmA, tmB;
    if (int rc = make_tmap_2d(&tmA, A.ptr, A.rows, A.cols, A.ld, BK, tc::TC_BM)) return rc;
    if (int rc = make_tmap_2d(&tmB, B.ptr, B.rows, B.cols, B.ld, BK, BN)) return rc;
    static_assert(EW == 4 || (EW == 8 && EPI != tc::EPI_ASSIGN && BN >= 64), "8 epilogue warps split the tile's columns");
    auto kern = tc::tc_gemm_bres_kernel<T, BN, EPI, EW, CL>;
    static PerDeviceSize attr_smem;
    EPC_CUDA(ensure_dyn_smem(kern, smem, attr_smem));
    const int NT = p.N / BN;
    const int m_tiles = (p.M + tc::TC_BM - 1) / tc::TC_BM;
    if (CL > 1) {
        // one cluster = the NT N tiles of an M-tile stream; as many clusters as fit on the device at once (persistent)
        EPC_CHECK_ARG(NT == CL, "tc_gemm_bres: cluster size %d needs N/BN == %d, got %d", CL, CL, NT);
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3(CL, 1, 1); cfg.blockDim = dim3(64 + 32 * EW, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cfg.attrs = attr; cfg.numAttrs = 1;
        static PerDeviceSize max_clusters_cache;
        std::atomic<size_t>& mc = max_clusters_cache.v[current_device_slot()];
        int max_clusters = (int)mc.load(std::memory_order_acquire);
        if (!max_clusters) {
            cfg.gridDim = dim3(CL * (sm_count() / CL), 1, 1);
            int n = 0;
            EPC_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
            EPC_CHECK_ARG(n >= 1, "tc_gemm_bres: no cluster of %d CTAs fits on this device", CL);
            max_clusters = n;
            mc.store((size_t)n, std::memory_order_release);
        }
        const int clusters = max_clusters < m_tiles ? max_clusters : m_tiles;
        cfg.gridDim = dim3(clusters * CL, 1, 1);
        EPC_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p, a_stages, ex));
        count_launch();
        return EPC_OK;
    }
    int per_nt = persistent_ctas("EPC_HEAD_CTAS") / NT;
    if (per_nt < 1) per_nt = 1;
    if (per_nt > m_tiles) per_nt = m_tiles;
    kern<<<per_nt * NT, 64 + 32 * EW, smem, st>>>(tmA, tmB, p, a_stages, ex);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

}  // namespace epc
