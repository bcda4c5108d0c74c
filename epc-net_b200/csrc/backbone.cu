// K2 -- ProxyConv backbone (models/epc-net.py:62-132, models/epc-net-l.py:44-80).
//
// The reference realises  m_i = (1/20) sum_{j in N(i)} x_j  as a dense (N x N mask) x (N x 64) batch
// matmul per block (2.15 GFLOP and a 64 MiB read each).  Here the neighbour lists of K1 are gathered
// directly (20 x 128 B bf16 rows per point, L1/L2 resident thanks to the Morton order), and the block body
//     t = m - x ; t = conv_a(t) ; t = conv_b(t) ; out = t + m ; x' = conv_{b+1}(out)
// runs on the tile while it is in shared memory.  Rows whose thresholded set has > 20 members (ties at
// the 20th distance, utils/tf_util.py:663-665) take an exact dense re-scan of the cloud.
//
// Activations between blocks travel in a 16-bit format (fp16, or bf16 for clouds that leave the fp16 range -- see
// FMT_* below), fp32 accumulation everywhere; the three 64x64 layers run on tcgen05 (kind::f16) with TMEM accumulators.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "tc_gemm.cuh"
#include "kernels.h"

namespace epc {

// ---- 16-bit storage formats of the inter-block activations ------------------------------------------------------
//   FMT_F16  fast path: fp16 = the TF32 operand precision (10-bit mantissa); conversions saturate and every producer
//            checks its fp32 values against the fp16 range -- a cloud whose activations leave it is flagged;
//            and re-done afterwards by the fp32/TF32 pass of backbone_f32.cu (e.g. the all-zero "fake" clouds of
//            evaluate.py:425-430, whose thresholded neighbour sets hold all N points so that activations grow by N/20
//            per block);
//   FMT_BF16 the bf16 slice of the concat buffer (operand of the bf16 conv5 GEMM).
enum { FMT_F16 = 0, FMT_BF16 = 1 };
constexpr float F16_MAX = 65504.f;

template <int FMT>
__device__ __forceinline__ uint32_t pack2(float a, float b) {      // (a,b) -> 16-bit pair, a in the low half
    uint32_t r;
    if (FMT == FMT_F16)
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
template <int FMT>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
    if (FMT == FMT_F16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
// lo += (float)u.lo ; hi += (float)u.hi   (one mixed-precision FHADD each on sm_100)
template <int FMT>
__device__ __forceinline__ void add2(float& lo, float& hi, uint32_t u) {
    if (FMT == FMT_F16)
        asm("{.reg .f16 l, h; mov.b32 {l, h}, %2; add.rn.f32.f16 %0, l, %0; add.rn.f32.f16 %1, h, %1;}" : "+f"(lo), "+f"(hi) : "r"(u));
    else
        asm("{.reg .b16 l, h; mov.b32 {l, h}, %2; add.rn.f32.bf16 %0, l, %0; add.rn.f32.bf16 %1, h, %1;}" : "+f"(lo), "+f"(hi) : "r"(u));
}
template <int FMT>
__device__ __forceinline__ void add8(float (&acc)[8], const uint4& v) {
    add2<FMT>(acc[0], acc[1], v.x);
    add2<FMT>(acc[2], acc[3], v.y);
    add2<FMT>(acc[4], acc[5], v.z);
    add2<FMT>(acc[6], acc[7], v.w);
}
template <int FMT>
__device__ __forceinline__ uint4 pack8(const float (&o)[8]) {
    return make_uint4(pack2<FMT>(o[0], o[1]), pack2<FMT>(o[2], o[3]), pack2<FMT>(o[4], o[5]), pack2<FMT>(o[6], o[7]));
}
__device__ __forceinline__ float max8(const float (&o)[8]) {
    return fmaxf(fmaxf(fmaxf(o[0], o[1]), fmaxf(o[2], o[3])), fmaxf(fmaxf(o[4], o[5]), fmaxf(o[6], o[7])));
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void load_bias8(uint32_t addr, float (&b)[8]) {
    const uint4 lo = lds128(addr), hi = lds128(addr + 16);
    b[0] = __uint_as_float(lo.x); b[1] = __uint_as_float(lo.y); b[2] = __uint_as_float(lo.z); b[3] = __uint_as_float(lo.w);
    b[4] = __uint_as_float(hi.x); b[5] = __uint_as_float(hi.y); b[6] = __uint_as_float(hi.z); b[7] = __uint_as_float(hi.w);
}
// x0 = relu(BN(p W1 + b1)), cin = 3  (models/epc-net.py:66-69) -> 16-bit [B,N,64].  grid (ceil(N / 256), B): a CTA
// converts 256 points (8 threads per point per pass, 8 passes), so the weight staging is amortised.
constexpr int CI_POINTS = 256;
template <int FMT>
__global__ void __launch_bounds__(256) conv_in_kernel(const float4* __restrict__ sorted, int N, const float* __restrict__ W,
                                                      const float* __restrict__ bias, uint16_t* __restrict__ x, int* __restrict__ flags) {
    const int b = blockIdx.y;
    __shared__ float sW[4][64];                    // rows 0..2: W[k][c]; row 3: bias
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sW[i >> 6][i & 63] = (i < 192) ? W[i] : bias[i - 192];
    __syncthreads();
    const int c8 = (threadIdx.x & 7) * 8;
    float w0[8], w1[8], w2[8], bb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        w0[i] = sW[0][c8 + i];
        w1[i] = sW[1][c8 + i];
        w2[i] = sW[2][c8 + i];
        bb[i] = sW[3][c8 + i];
    }
    bool over = false;
#pragma unroll 2
    for (int pass = 0; pass < CI_POINTS / 32; ++pass) {
        const int n = blockIdx.x * CI_POINTS + pass * 32 + (threadIdx.x >> 3);
        if (n >= N) break;
        const size_t r = (size_t)b * N + n;
        const float4 p = __ldg(sorted + r);
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaxf(fmaf(p.z, w2[i], fmaf(p.y, w1[i], fmaf(p.x, w0[i], bb[i]))), 0.f);
        if (FMT == FMT_F16) over |= (max8(o) > F16_MAX);
        *reinterpret_cast<uint4*>(x + r * 64 + c8) = pack8<FMT>(o);
    }
    if (FMT == FMT_F16 && over) flags[b] = 1;
}

int conv_in(const float4* sorted, int B, int N, const DenseDev& L, uint16_t* x, int* flags, cudaStream_t st) {
    EPC_CHECK_ARG(L.cin == 3 && L.cout == 64, "conv_in expects a 3->64 layer, got %d->%d", L.cin, L.cout);
    if (B == 0) return EPC_OK;
    dim3 grid((N + CI_POINTS - 1) / CI_POINTS, B);
    conv_in_kernel<FMT_F16><<<grid, 256, 0, st>>>(sorted, N, L.W, L.b, x, flags);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// ------------------------------------------------------------------------------------------------
// ProxyConv block: persistent, warp-specialised, one CTA per SM, tiles of 128 consecutive (Morton-ordered) points.
//
//   warps 8..23  GATHER   lane group (8 lanes) per point, 16 B of each 128 B neighbour row per lane.  The tile's own
//                         128 rows (69 % of all neighbour links in Morton order) are brought into shared memory by one
//                         TMA bulk copy, a tile ahead; K1 lists each point's out-of-tile neighbours first, so those go
//                         out as LDG.128 (up to 10 in flight per lane) while the in-tile rows are summed from shared
//                         memory.  fp32 accumulation (FHADD); writes m (16-bit) and t = m - x (the UMMA A tile, K-major
//                         SW128) of ring stage s, then arrives on full[s]
//   warps 0..3 / 4..7     two CONSUMER groups alternating tiles; thread = point (= TMEM lane).  Per tile
//                         MMA conv_a -> relu -> A tile (in place) -> MMA conv_b -> relu + m -> A tile, block output
//                         (bf16 concat slice via smem staging, coalesced) -> MMA conv_{b+1} -> relu -> 16-bit x' (staged,
//                         coalesced); thread 0 of the group issues the tcgen05.mma's, accumulators live in TMEM.
// Shared memory: 3 weight images (8 KB each, pre-swizzled at model creation) + 4 stages x (A 16 KB + m 16 KB + window 16 KB).
// ------------------------------------------------------------------------------------------------
constexpr int PB_TILE = 128;
#ifndef PB_GROUPS
#define PB_GROUPS 2
#endif
constexpr int PB_CONS_GROUPS = PB_GROUPS;        // consumer groups of 4 warps, taking tiles in turn
constexpr int PB_CONS_WARPS = 4 * PB_CONS_GROUPS;
#ifndef PB_GWARPS
#define PB_GWARPS 16
#endif
constexpr int PB_GATHER_WARPS = PB_GWARPS;         // the 32 quads (4 points) of a tile are dealt round-robin, continuing across tiles
constexpr int PB_THREADS = 32 * (PB_CONS_WARPS + PB_GATHER_WARPS);
constexpr int PB_STAGES = 4;                     // A/M ring (gather -> consumers)
constexpr int PB_WSTAGES = 3;                    // window ring (TMA -> gather), filled two tiles ahead
constexpr uint32_t PB_A_BYTES = 128 * 128;       // 128 rows x 64 x 16 bit
constexpr uint32_t PB_W_BYTES = 64 * 128;        // 64 rows (cout) x 64 (cin) x 16 bit
constexpr uint32_t PB_NBR_BYTES = PB_TILE * KNN_K * 2;     // 5120
constexpr uint32_t PB_CNT_BYTES = PB_TILE * 4;             // 512
constexpr uint32_t PB_STAGE_BYTES = 2 * PB_A_BYTES;        // A | M            (SW128 tiles: 1024 B aligned)
constexpr uint32_t PB_WSTAGE_BYTES = (PB_A_BYTES + PB_NBR_BYTES + PB_CNT_BYTES + 1023) / 1024 * 1024;     // window | nbr | cnt
// offsets from the 1024-aligned base
constexpr uint32_t PB_OFF_W = 0;
constexpr uint32_t PB_OFF_STAGE = 3 * PB_W_BYTES;
constexpr uint32_t PB_OFF_WSTAGE = PB_OFF_STAGE + PB_STAGES * PB_STAGE_BYTES;
constexpr uint32_t PB_OFF_BIAS = PB_OFF_WSTAGE + PB_WSTAGES * PB_WSTAGE_BYTES;
constexpr uint32_t PB_OFF_BAR = PB_OFF_BIAS + 3 * 64 * 4;          // full[S] empty[S] wfull[WS] wempty[WS] mma[2]
constexpr uint32_t PB_OFF_TMEM = PB_OFF_BAR + (2 * PB_STAGES + 2 * PB_WSTAGES + 2) * 8;
constexpr size_t PB_SMEM = 1024 + PB_OFF_TMEM + 16;
static_assert(PB_SMEM <= 227 * 1024, "proxy_block shared memory");
static_assert(PB_GATHER_WARPS <= PB_TILE / 4, "every gather warp must own at least one quad of every tile (barrier counts)");
static_assert(PB_STAGE_BYTES % 1024 == 0 && PB_OFF_STAGE % 1024 == 0, "UMMA SW128 tiles must be 1024-byte aligned");
// per stage
constexpr uint32_t PB_ST_A = 0, PB_ST_M = PB_A_BYTES;
constexpr uint32_t PB_WS_WIN = 0, PB_WS_NBR = PB_A_BYTES, PB_WS_CNT = PB_A_BYTES + PB_NBR_BYTES;

// byte offset of 16-byte chunk c (8 channels) of row r in a [rows x 64] 16-bit K-major SW128 tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

// mbarrier helpers on shared-space addresses (no generic->shared conversion per call)
__device__ __forceinline__ void bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool bar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {          // short waits (MMA completion)
    while (!bar_try(bar, parity)) {}
}
template <int NS = 128>
__device__ __forceinline__ void bar_wait_backoff(uint32_t bar, uint32_t parity) {  // long waits: leave the issue slots to the others
    while (!bar_try(bar, parity)) __nanosleep(NS);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mma_commit_u32(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int FMT>
__device__ __forceinline__ void pb_issue_gemm(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, uint32_t bar) {
    constexpr uint32_t idesc = tc::make_idesc(FMT == FMT_F16 ? 0 : 1, 128, 64, 0, 0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = tc::smem_desc_sw128(a_addr + kk * 32, 16, 1024);
        const uint64_t db = tc::smem_desc_sw128(w_addr + kk * 32, 16, 1024);
        tc::mma_ss<true>(tmem_d, da, db, idesc, kk != 0);
    }
    mma_commit_u32(bar);
}

__device__ __forceinline__ void group_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory"); }

struct PbArgs {
    const uint16_t* x;        // [B,N,64] 16-bit block input (output of the block's first conv)
    const uint16_t* nbr;      // [B,N,20] neighbour positions, out-of-tile ones first
    const float* kthd;        // [B,N]
    const int* cnt;           // [B,N]  |thresholded set| | (#out-of-tile neighbours << 24)
    const uint32_t* tie;      // [B][TIE_WORDS] set members beyond the 20 listed ones (kernels.h)
    const float4* sorted;     // [B,N]
    int* flags;               // [B]  cloud left the fp16 range
    int N, arith, tiles_per_cloud, num_tiles;
    float inv_div;
    const uint4 *Wa_img, *Wb_img, *Wn_img;     // swizzled weight images in the kernel's format
    const float *ba, *bb, *bn;
    float bias[192];          // ba | bb | bn by value: the epilogues read them as constant-bank operands (a warp-wide LDS.128 of one
                              // shared-memory address still costs 4 wavefronts of the L1 data pipe, which this kernel is bound by)
    float* concat32;          // [B,N,ctot] fp32 (TF32-rounded) or nullptr
    __nv_bfloat16* concat16;  // [B,N,ctot] 16-bit or nullptr: bf16, or fp16 when concat_f16 (EPC-Net-L: the fp16 conv5 operand)
    int concat_f16;
    float* cloud_absmax;      // [B] or nullptr: running max of the bf16 concat values of each cloud (>= 0; atomicMax on the bits) --
                              // the fp8 head derives the cloud's conv5 output bound from it (head_fp8.cu)
    int ctot, coff;
    uint16_t* xnext;          // [B,N,64] 16-bit (HAS_NEXT)
    CUtensorMap tmC16, tmXn;  // the 16-bit outputs leave as TMA stores of the swizzled staging tiles: [B*N, ctot] box {64 ch, 128 rows} of
                              // concat16 at column coff; [B*N, 64] box {64, 128} of xnext
};

template <bool HAS_NEXT, int FMT>
__global__ void __launch_bounds__(PB_THREADS, 1) proxy_block_kernel(const __grid_constant__ PbArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;      // shared-space address; every buffer is base + constant
    uint8_t* gbase = smem_raw + (base - tc::smem_u32(smem_raw));          // same location as a generic pointer (prologue only)
    const uint32_t bar_full = base + PB_OFF_BAR, bar_empty = bar_full + 8 * PB_STAGES, bar_wfull = bar_empty + 8 * PB_STAGES,
                   bar_wempty = bar_wfull + 8 * PB_WSTAGES, bar_mma = bar_wempty + 8 * PB_WSTAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N;

    // ---- prologue: weights, barriers, TMEM ------------------------------------------------------------------
    {
        uint4* s = reinterpret_cast<uint4*>(gbase + PB_OFF_W);
        constexpr int per = PB_W_BYTES / 16;
        for (int i = tid; i < 3 * per; i += PB_THREADS) {
            const int w = i / per, o = i - w * per;
            const uint4* src = (w == 0) ? p.Wa_img : (w == 1) ? p.Wb_img : p.Wn_img;
            s[i] = (HAS_NEXT || w < 2) ? __ldg(src + o) : make_uint4(0, 0, 0, 0);
        }
        float* sBias = reinterpret_cast<float*>(gbase + PB_OFF_BIAS);
        if (tid < 64) {
            sBias[tid] = p.ba[tid];
            sBias[64 + tid] = p.bb[tid];
            sBias[128 + tid] = HAS_NEXT ? p.bn[tid] : 0.f;
        }
        if (tid == 0) {
            uint64_t* bars = reinterpret_cast<uint64_t*>(gbase + PB_OFF_BAR);
            for (int s2 = 0; s2 < PB_STAGES; ++s2) {
                tc::mbar_init(&bars[s2], PB_GATHER_WARPS);
                tc::mbar_init(&bars[PB_STAGES + s2], 4);
            }
            for (int s2 = 0; s2 < PB_WSTAGES; ++s2) {
                tc::mbar_init(&bars[2 * PB_STAGES + s2], 1);
                tc::mbar_init(&bars[2 * PB_STAGES + PB_WSTAGES + s2], PB_GATHER_WARPS);
            }
            tc::mbar_init(&bars[2 * PB_STAGES + 2 * PB_WSTAGES], 1);
            tc::mbar_init(&bars[2 * PB_STAGES + 2 * PB_WSTAGES + 1], 1);
            tc::fence_barrier_init();
        }
        if (warp == 0) {
            tc::tmem_alloc(reinterpret_cast<uint32_t*>(gbase + PB_OFF_TMEM), 128);
            tc::tmem_relinquish();
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    }
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + PB_OFF_TMEM);
    const int n_local = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
    auto next_active = [&](int it) { return it + 1; };

    if (warp >= PB_CONS_WARPS) {
        // =========================================== GATHER ===========================================
        const int gw = warp - PB_CONS_WARPS;
        const int grp = lane >> 3, c = lane & 7;
        auto issue_window = [&](int it, int u) {        // one thread: TMA the tile's own rows, neighbour lists and counts
            const int ws = u % PB_WSTAGES;
            const uint32_t use = (uint32_t)(u / PB_WSTAGES);
            const size_t row0 = (size_t)(blockIdx.x + it * gridDim.x) * PB_TILE;
            const uint32_t wst = base + PB_OFF_WSTAGE + (uint32_t)ws * PB_WSTAGE_BYTES;
            bar_wait_backoff(bar_wempty + 8 * ws, (use & 1u) ^ 1u);      // every gather warp is done with this window
            bar_expect_tx(bar_wfull + 8 * ws, PB_A_BYTES + PB_NBR_BYTES + PB_CNT_BYTES);
            bulk_g2s(wst + PB_WS_WIN, p.x + row0 * 64, PB_A_BYTES, bar_wfull + 8 * ws);
            bulk_g2s(wst + PB_WS_NBR, p.nbr + row0 * KNN_K, PB_NBR_BYTES, bar_wfull + 8 * ws);
            bulk_g2s(wst + PB_WS_CNT, p.cnt + row0, PB_CNT_BYTES, bar_wfull + 8 * ws);
        };
        const bool elected = (gw == 0 && lane == 0);
        int it = next_active(-1);
        int it1 = (it < n_local) ? next_active(it) : n_local;       // the window ring runs two tiles ahead
        int u = 0;
        if (elected) {
            if (it < n_local) issue_window(it, 0);
            if (it1 < n_local) issue_window(it1, 1);
        }
        while (it < n_local) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int b = tile / p.tiles_per_cloud;
            const int tile0 = (tile - b * p.tiles_per_cloud) * PB_TILE;
            const int s = u % PB_STAGES, ws = u % PB_WSTAGES;
            const uint32_t use = (uint32_t)(u / PB_STAGES), wuse = (uint32_t)(u / PB_WSTAGES);
            const uint32_t st = base + PB_OFF_STAGE + (uint32_t)s * PB_STAGE_BYTES;
            const uint32_t wst = base + PB_OFF_WSTAGE + (uint32_t)ws * PB_WSTAGE_BYTES;
            uint32_t win = wst + PB_WS_WIN + (uint32_t)c * 16u - (uint32_t)tile0 * 128u;      // + j*128 = row j of the window
            const uint8_t* xb = reinterpret_cast<const uint8_t*>(p.x + (size_t)b * N * 64) + c * 16;      // + j*128
            asm volatile("" : "+l"(xb), "+r"(win));     // keep both bases in registers (ptxas otherwise re-derives them per row)
            const int it2 = (it1 < n_local) ? next_active(it1) : n_local;
            if (elected && it2 < n_local) issue_window(it2, u + 2);
            __syncwarp();
            bar_wait_backoff(bar_wfull + 8 * ws, wuse & 1u);
            bool first = true;
#pragma unroll 1
            for (int quad = (gw + PB_GATHER_WARPS - (u * (PB_TILE / 4)) % PB_GATHER_WARPS) % PB_GATHER_WARPS; quad < PB_TILE / 4;
                 quad += PB_GATHER_WARPS) {
                const int pl = quad * 4 + grp;      // this lane group's point within the tile
                uint32_t idx[10];
                {
                    const uint32_t ia = wst + PB_WS_NBR + (uint32_t)pl * (KNN_K * 2);
#pragma unroll
                    for (int q = 0; q < 5; ++q)
                        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(idx[2 * q]), "=r"(idx[2 * q + 1]) : "r"(ia + 8 * q));
                }
                int cntw;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(cntw) : "r"(wst + PB_WS_CNT + (uint32_t)pl * 4u));
                const int nout = (cntw >> 24) & 0xff;
                const int count = cntw & 0xffffff;
                auto jq = [&](int q) { return (idx[q >> 1] >> ((q & 1) * 16)) & 0xffffu; };
                float acc[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                // rows 0..9: the out-of-tile neighbours (listed first) go out as global loads, the rest come from the window
                uint4 v[10];
#pragma unroll
                for (int q = 0; q < 10; ++q) {
                    const uint32_t j = jq(q);
                    if (q < nout)
                        v[q] = __ldg(reinterpret_cast<const uint4*>(xb + j * 128u));
                    else
                        v[q] = lds128(win + j * 128u);
                }
                // rows 10..19 while those are in flight
#pragma unroll
                for (int q = 10; q < KNN_K; ++q) {
                    const uint32_t j = jq(q);
                    uint4 w;
                    if (q < nout)
                        w = __ldg(reinterpret_cast<const uint4*>(xb + j * 128u));
                    else
                        w = lds128(win + j * 128u);
                    add8<FMT>(acc, w);
                }
                const uint4 xi = lds128(win + (uint32_t)(tile0 + pl) * 128u);
#pragma unroll
                for (int q = 0; q < 10; ++q) add8<FMT>(acc, v[q]);
                // ties at the 20th distance: the set is {j : d_ij <= kthd_i}; re-scan the cloud exactly (rare)
                const unsigned tie_mask = __ballot_sync(FULL, count != KNN_K);
                if (tie_mask) {
                    const size_t row_tile = (size_t)b * N + tile0;
                    const uint32_t* tl = p.tie + (size_t)b * TIE_WORDS;
                    const uint32_t n_tie = __ldg(tl);
                    for (int g = 0; g < 4; ++g) {
                        if (!((tie_mask >> (8 * g)) & 1u)) continue;                    // warp-uniform
                        const size_t trow = row_tile + (pl - grp) + g;
                        const float thr = __ldg(p.kthd + trow);
                        const int want = (__shfl_sync(FULL, count, 8 * g) & 0xffffff) - KNN_K;     // members beyond the listed 20
                        // (a) the kNN kernel logged them in the cloud's tie list: add exactly those
                        int found = 0;
                        float a2[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) a2[i] = 0.f;
                        if (n_tie <= TIE_CAP) {
                            const uint32_t tag = (uint32_t)(tile0 + (pl - grp) + g);
                            for (uint32_t e0 = 0; e0 < n_tie; e0 += 32) {
                                uint2 ent = make_uint2(0xffffffffu, 0u);
                                if (e0 + lane < n_tie) ent = __ldg(reinterpret_cast<const uint2*>(tl + 2) + e0 + lane);
                                unsigned mk = __ballot_sync(FULL, (ent.x >> 16) == tag && ent.y == __float_as_uint(thr));
                                found += __popc(mk);
                                while (mk) {
                                    const int src = __ffs(mk) - 1;
                                    mk &= mk - 1;
                                    const uint32_t j = __shfl_sync(FULL, ent.x, src) & 0xffffu;
                                    if (grp == g) add8<FMT>(a2, __ldg(reinterpret_cast<const uint4*>(xb + (size_t)j * 128)));
                                }
                            }
                        }
                        if (n_tie <= TIE_CAP && found == want) {
                            if (grp == g) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) acc[i] += a2[i];
                            }
                            continue;
                        }
                        // (b) list overflow (degenerate clouds): exact re-scan of the cloud
                        const float4 qp = p.sorted[trow];
                        const float4* sp = p.sorted + (size_t)b * N;
#pragma unroll
                        for (int i = 0; i < 8; ++i) a2[i] = 0.f;
                        // N is a multiple of 128: four 32-point blocks per step, their loads in flight together
                        for (int j0 = 0; j0 < N; j0 += 128) {
                            float4 pj[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) pj[k] = __ldg(sp + j0 + 32 * k + lane);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float d = (p.arith == EPC_KNN_ARITH_MULADD)
                                                    ? canon_dist<0>(qp.x, qp.y, qp.z, qp.w, pj[k].x, pj[k].y, pj[k].z, pj[k].w)
                                                    : canon_dist<1>(qp.x, qp.y, qp.z, qp.w, pj[k].x, pj[k].y, pj[k].z, pj[k].w);
                                unsigned mk = __ballot_sync(FULL, d <= thr);
                                while (mk) {
                                    const int j = j0 + 32 * k + __ffs(mk) - 1;
                                    mk &= mk - 1;
                                    if (grp == g) add8<FMT>(a2, __ldg(reinterpret_cast<const uint4*>(xb + (size_t)j * 128)));
                                }
                            }
                        }
                        if (grp == g) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) acc[i] = a2[i];
                        }
                    }
                }
                float m[8], t[8];
                const float2 x01 = unpack2<FMT>(xi.x), x23 = unpack2<FMT>(xi.y), x45 = unpack2<FMT>(xi.z), x67 = unpack2<FMT>(xi.w);
                const float xs[8] = {x01.x, x01.y, x23.x, x23.y, x45.x, x45.y, x67.x, x67.y};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    m[i] = acc[i] * p.inv_div;           // x1 = matmul(mask, x) / float(k)   (models/epc-net.py:70-71)
                    t[i] = m[i] - xs[i];                 // t1 = x1 - x                       (:72)
                }
                if (FMT == FMT_F16 && max8(m) > F16_MAX) p.flags[b] = 1;     // x >= 0, so |t| <= max(m, x)
                if (first) bar_wait_backoff(bar_empty + 8 * s, (use & 1u) ^ 1u);      // the consumers have released this A/M stage
                first = false;
                const uint32_t off = sw128_off(pl, c);
                sts128(st + PB_ST_M + off, pack8<FMT>(m));
                sts128(st + PB_ST_A + off, pack8<FMT>(t));
            }
            tc::fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
            __syncwarp();
            if (lane == 0) {
                bar_arrive(bar_full + 8 * s);
                bar_arrive(bar_wempty + 8 * ws);
            }
            it = it1;
            it1 = it2;
            ++u;
        }
    } else {
        // ========================================== CONSUMERS ==========================================
        const int group = warp >> 2;                         // 0 / 1: alternate tiles
        const int gtid = tid & 127;                          // thread within the group = point within the tile = TMEM lane
        const bool leader = (gtid == 0);
        const uint32_t tmem_d = tmem_base + (uint32_t)(group * 64);
        const uint32_t trow = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t w0 = base + PB_OFF_W, w1 = w0 + PB_W_BYTES, w2 = w1 + PB_W_BYTES;
        const uint32_t my_mma = bar_mma + 8 * group;
        uint32_t mma_phase = 0;
        int u = 0;
        for (int it = next_active(-1); it < n_local; it = next_active(it), ++u) {
            if ((u % PB_CONS_GROUPS) != group) continue;
            const int tile = blockIdx.x + it * gridDim.x;
            const int b = tile / p.tiles_per_cloud;
            const int s = u % PB_STAGES;
            const uint32_t use = (uint32_t)(u / PB_STAGES);
            const uint32_t a_st = base + PB_OFF_STAGE + (uint32_t)s * PB_STAGE_BYTES + PB_ST_A;
            const uint32_t m_st = a_st + (PB_ST_M - PB_ST_A);
            const size_t grow0 = (size_t)tile * PB_TILE;     // first global row of the tile
            float vmax = 0.f;                                // largest value this thread converts to the 16-bit format
            float cmax = 0.f;                                // largest block output of this thread's row (cloud_absmax)

            bar_wait_backoff<400>(bar_full + 8 * s, use & 1u);
            tc::tc_fence_after();
            // ---- conv_a --------------------------------------------------------------------------------------
            if (leader) pb_issue_gemm<FMT>(tmem_d, a_st, w0, my_mma);
            bar_wait(my_mma, mma_phase);
            mma_phase ^= 1u;
            tc::tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[32];
                tc::tmem_ld32(trow + 32u * h, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float bs[8], o[8];
                    {
#pragma unroll
                        for (int e = 0; e < 8; ++e) bs[e] = p.bias[32 * h + 8 * q + e];
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = fmaxf(v[8 * q + e] + bs[e], 0.f);
                    if (FMT == FMT_F16) vmax = fmaxf(vmax, max8(o));
                    sts128(a_st + sw128_off(gtid, 4 * h + q), pack8<FMT>(o));
                }
            }
            tc::fence_proxy_async();
            tc::tc_fence_before();
            group_sync(group);
            // ---- conv_b, residual, block output --------------------------------------------------------------
            if (leader) {
                tc::tc_fence_after();
                pb_issue_gemm<FMT>(tmem_d, a_st, w1, my_mma);
            }
            bar_wait(my_mma, mma_phase);
            mma_phase ^= 1u;
            tc::tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[32];
                tc::tmem_ld32(trow + 32u * h, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t off = sw128_off(gtid, 4 * h + q);
                    const uint4 mm = lds128(m_st + off);
                    const float2 m01 = unpack2<FMT>(mm.x), m23 = unpack2<FMT>(mm.y), m45 = unpack2<FMT>(mm.z), m67 = unpack2<FMT>(mm.w);
                    const float ms[8] = {m01.x, m01.y, m23.x, m23.y, m45.x, m45.y, m67.x, m67.y};
                    float bs[8], o[8];
                    {
#pragma unroll
                        for (int e = 0; e < 8; ++e) bs[e] = p.bias[64 + 32 * h + 8 * q + e];
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = fmaxf(v[8 * q + e] + bs[e], 0.f) + ms[e];        // x_b = relu(conv_b) + m  (:81)
                    if (HAS_NEXT) {
                        if (FMT == FMT_F16) vmax = fmaxf(vmax, max8(o));
                        sts128(a_st + off, pack8<FMT>(o));
                    } else if (p.concat_f16) {
                        vmax = fmaxf(vmax, max8(o));           // the fp16 concat slice must stay in range too
                    }
                    if (p.concat16)                            // 16-bit concat slice, staged over the consumed m chunk
                        sts128(m_st + off, p.concat_f16 ? pack8<FMT_F16>(o) : pack8<FMT_BF16>(o));
                    if (p.cloud_absmax) cmax = fmaxf(cmax, max8(o));
                    if (p.concat32) {                      // operand of the TF32 conv5 (EPC-Net-L, KD export): store it rounded
                        float4* dst = reinterpret_cast<float4*>(p.concat32 + (grow0 + gtid) * p.ctot + p.coff + 32 * h + 8 * q);
                        dst[0] = make_float4(round_tf32(o[0]), round_tf32(o[1]), round_tf32(o[2]), round_tf32(o[3]));
                        dst[1] = make_float4(round_tf32(o[4]), round_tf32(o[5]), round_tf32(o[6]), round_tf32(o[7]));
                    }
                }
            }
            tc::fence_proxy_async();
            tc::tc_fence_before();
            group_sync(group);
            if (HAS_NEXT && leader) {
                tc::tc_fence_after();
                pb_issue_gemm<FMT>(tmem_d, a_st, w2, my_mma);     // first conv of the next block
            }
            if (p.concat16) {
                // the staged, 128B-swizzled [128 rows x 64 channels] tile is exactly a TMA box: one store per tile instead of 8
                // LDS.128 + 8 STG.128 per thread (the kernel is bound by L1 data-pipe wavefronts)
                if (leader) tc::tma_store_2d(&p.tmC16, m_st, p.coff, (int)grow0);
                if (p.cloud_absmax) {                      // block outputs are >= 0 (relu + a mean of non-negative rows); the maximum of
                                                           // the bf16-rounded values = the rounded maximum (rounding is monotone)
                    const float r = __bfloat162float(__float2bfloat16_rn(cmax));
                    const unsigned mx = __reduce_max_sync(FULL, __float_as_uint(fmaxf(r, 0.f)));
                    if ((gtid & 31) == 0 && mx) atomicMax(reinterpret_cast<unsigned*>(p.cloud_absmax) + b, mx);
                }
            }
            if (HAS_NEXT) {
                bar_wait(my_mma, mma_phase);
                mma_phase ^= 1u;
                tc::tc_fence_after();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float v[32];
                    tc::tmem_ld32(trow + 32u * h, v);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float bs[8], o[8];
                        {
#pragma unroll
                        for (int e = 0; e < 8; ++e) bs[e] = p.bias[128 + 32 * h + 8 * q + e];
                    }
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = fmaxf(v[8 * q + e] + bs[e], 0.f);
                        if (FMT == FMT_F16) vmax = fmaxf(vmax, max8(o));
                        sts128(a_st + sw128_off(gtid, 4 * h + q), pack8<FMT>(o));   // staging
                    }
                }
                tc::fence_proxy_async();
                tc::tc_fence_before();
                group_sync(group);
                if (leader) tc::tma_store_2d(&p.tmXn, a_st, 0, (int)grow0);       // the staged tile, one store
            }
            if (FMT == FMT_F16 && vmax > F16_MAX) p.flags[b] = 1;
            if (leader) tc::bulk_wait_read0();             // the tile's TMA stores have read the stage before the gather warps get it back
            __syncwarp();
            if (lane == 0) bar_arrive(bar_empty + 8 * s);
        }
        if (leader) tc::bulk_wait0();                      // this group's TMA stores are complete before the CTA retires
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 128);
}

int proxy_block(const uint16_t* x, const KnnState& g, int B, int N, int arith, float divisor, const DenseDev& conv_a,
                const DenseDev& conv_b, const DenseDev* conv_next, float* concat, __nv_bfloat16* concat16, int ctot,
                int coff, uint16_t* xnext, int* flags, float* cloud_absmax, int concat_f16, cudaStream_t st) {
    EPC_CHECK_ARG(conv_a.cin == 64 && conv_a.cout == 64 && conv_b.cin == 64 && conv_b.cout == 64,
                  "ProxyConv block layers must be 64->64");
    EPC_CHECK_ARG(N % PB_TILE == 0, "proxy_block: N=%d must be a multiple of %d", N, PB_TILE);
    EPC_CHECK_ARG(conv_a.Wimg && conv_b.Wimg && (!conv_next || conv_next->Wimg), "proxy_block: missing swizzled weight images");
    EPC_CHECK_ARG(ctot % 8 == 0 && coff % 8 == 0, "proxy_block: concat slice must be 16-byte aligned");
    if (B == 0) return EPC_OK;
    static PerDeviceSize attr_a, attr_b;
    EPC_CUDA(ensure_dyn_smem(proxy_block_kernel<true, FMT_F16>, PB_SMEM, attr_a));
    EPC_CUDA(ensure_dyn_smem(proxy_block_kernel<false, FMT_F16>, PB_SMEM, attr_b));
    auto img = [&](const DenseDev& L) { return reinterpret_cast<const uint4*>(L.Wimg); };
    PbArgs a = {};
    a.x = x; a.nbr = g.nbr; a.kthd = g.kthd; a.cnt = g.cnt; a.tie = g.tie; a.sorted = g.sorted; a.flags = flags;
    a.N = N; a.arith = arith; a.tiles_per_cloud = N / PB_TILE; a.num_tiles = B * (N / PB_TILE);
    a.inv_div = 1.0f / divisor;                           // one rounding away from a true division by float(k)
    a.Wa_img = img(conv_a); a.ba = conv_a.b;
    a.Wb_img = img(conv_b); a.bb = conv_b.b;
    a.Wn_img = conv_next ? img(*conv_next) : nullptr;
    a.bn = conv_next ? conv_next->b : nullptr;
    EPC_CHECK_ARG(conv_a.b_host && conv_b.b_host && (!conv_next || conv_next->b_host), "proxy_block: missing host copies of the biases%s", "");
    memcpy(a.bias, conv_a.b_host, 64 * sizeof(float));
    memcpy(a.bias + 64, conv_b.b_host, 64 * sizeof(float));
    if (conv_next) memcpy(a.bias + 128, conv_next->b_host, 64 * sizeof(float));
    if (concat16)
        if (int rc = make_tmap_2d(&a.tmC16, reinterpret_cast<const uint16_t*>(concat16), (uint64_t)B * N, (uint64_t)ctot, (uint64_t)ctot, 64, PB_TILE)) return rc;
    if (conv_next)
        if (int rc = make_tmap_2d(&a.tmXn, reinterpret_cast<const uint16_t*>(xnext), (uint64_t)B * N, 64, 64, 64, PB_TILE)) return rc;
    a.concat32 = concat; a.concat16 = concat16; a.ctot = ctot; a.coff = coff; a.xnext = xnext; a.cloud_absmax = cloud_absmax; a.concat_f16 = concat_f16;
    const int ctas = persistent_ctas("EPC_BLOCK_CTAS");
    const int grid = a.num_tiles < ctas ? a.num_tiles : ctas;
    if (conv_next)
        proxy_block_kernel<true, FMT_F16><<<grid, PB_THREADS, PB_SMEM, st>>>(a);
    else
        proxy_block_kernel<false, FMT_F16><<<grid, PB_THREADS, PB_SMEM, st>>>(a);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// Host: [cin=64][cout=64] folded weights -> shared-memory images of the K-major, 128B-swizzled 16-bit B operand
// (element (n,k) = W[k][n]): 64 rows (n) x 128 B; 16-byte chunk (k>>3) XOR-ed with (n & 7).
// img: 4096 fp16 = 8 KB.
void make_w64_image(const float* W, uint16_t* img) {
    for (int k = 0; k < 64; ++k)
        for (int n = 0; n < 64; ++n) {
            const size_t off = ((size_t)n * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2) / 2;
            const __half h = __float2half_rn(W[k * 64 + n]);
            img[off] = *reinterpret_cast<const uint16_t*>(&h);
        }
}

}  // namespace epc
