// Tensor-core GEMMs for sm_100a: tcgen05.mma (kind::tf32) with TMEM accumulators, operands staged by TMA into
// 128B-swizzled shared memory, mbarrier producer/consumer pipeline, one elected thread issuing the MMAs.
// Hand-written PTX; descriptor bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables
// (cross-checked against cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace epc {
namespace tc {

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

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

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// A operand from TMEM (M lanes x K columns of 32-bit), B from shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
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
// Shared-memory matrix descriptor, K-major operand tile in the 128B-swizzle layout TMA writes for a
// {32 x fp32 = 128 B, rows} box: row r at r*128 B, 16-byte chunks XOR-ed with (r & 7); 8-row groups 1024 B apart.
//   bits [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major) |
//   [32,46) stride byte offset >> 4 (= 1024 B between 8-row groups) | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major:
//   [4,6) D format = 1 (F32) | [7,10) A format = 2 (TF32) | [10,13) B format = 2 | [15] A major = 0 (K) | [16] B major = 0 (K)
//   [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------
// C[M,N] = act(A[M,K] . B[N,K]^T + bias)     A, B fp32 K-major (row-major), TF32 tensor cores, fp32 out
// grid (ceil(M/128), N/BN); 192 threads: warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2-5 epilogue
// ---------------------------------------------------------------------------------------------------------
constexpr int TC_BM = 128;
constexpr int TC_BK = 32;   // 32 fp32 = 128 B = one swizzle row

template <int BN>
struct NtCfg {
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr uint32_t A_BYTES = TC_BM * TC_BK * 4;
    static constexpr uint32_t B_BYTES = BN * TC_BK * 4;
    static constexpr size_t SMEM = (size_t)STAGES * (A_BYTES + B_BYTES) + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
tc_gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
                  const float* __restrict__ bias, int M, int N, int K, int ldc, int relu) {
    using Cfg = NtCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base;
    uint8_t* sB = base + (size_t)STAGES * Cfg::A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
    const int nkb = K / TC_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], Cfg::A_BYTES + Cfg::B_BYTES);
                tma_load_2d(sA + (size_t)s * Cfg::A_BYTES, &tmA, &full[s], kb * TC_BK, m0);
                tma_load_2d(sB + (size_t)s * Cfg::B_BYTES, &tmB, &full[s], kb * TC_BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_tf32(TC_BM, BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint64_t da = smem_desc_k_sw128(smem_u32(sA + (size_t)s * Cfg::A_BYTES));
                const uint64_t db = smem_desc_k_sw128(smem_u32(sB + (size_t)s * Cfg::B_BYTES));
#pragma unroll
                for (int kk = 0; kk < TC_BK / 8; ++kk) {
                    // advance 8 tf32 = 32 bytes along K inside the 128B swizzle row: +2 in the (addr >> 4) field
                    mma_tf32_ss(tmem_d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (kb | kk) != 0);
                }
                mma_commit(&empty[s]);          // frees the smem stage when these MMAs have read it
            }
            mma_commit(acc_full);               // accumulator complete
        }
    } else {
        // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31
        const int q = warp & 3;
        const int row = q * 32 + lane;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int m = m0 + row;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (m < M) {
                float* dst = C + (size_t)m * ldc + n0 + c0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 o;
                    o.x = v[i] + (bias ? __ldg(bias + n0 + c0 + i) : 0.f);
                    o.y = v[i + 1] + (bias ? __ldg(bias + n0 + c0 + i + 1) : 0.f);
                    o.z = v[i + 2] + (bias ? __ldg(bias + n0 + c0 + i + 2) : 0.f);
                    o.w = v[i + 3] + (bias ? __ldg(bias + n0 + c0 + i + 3) : 0.f);
                    if (relu) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    *reinterpret_cast<float4*>(dst + i) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, BN < 32 ? 32 : BN);
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

// 2-D fp32 tensor [rows, cols] (cols contiguous, row pitch `ld` elements); box = {box_cols, box_rows}, 128B swizzle
inline int make_tmap_2d(CUtensorMap* tm, const float* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                        uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return EPC_ECUDA;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u ptr=%p", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows, ptr);
        return EPC_ECUDA;
    }
    return EPC_OK;
}

struct TcGemmNT {
    const float* A;      // [M, K] row-major (lda = K)
    const float* B;      // [N, K] row-major
    float* C;            // [M, ldc]
    const float* bias;   // [N] or nullptr
    int M, N, K, ldc, relu;
    int BN;              // 64 | 128 | 256
};

template <int BN>
inline int tc_gemm_nt_launch(const TcGemmNT& g, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    if (int rc = make_tmap_2d(&tmA, g.A, g.M, g.K, g.K, tc::TC_BK, tc::TC_BM)) return rc;
    if (int rc = make_tmap_2d(&tmB, g.B, g.N, g.K, g.K, tc::TC_BK, BN)) return rc;
    static bool attr = false;
    if (!attr) {
        EPC_CUDA(cudaFuncSetAttribute(tc::tc_gemm_nt_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)tc::NtCfg<BN>::SMEM));
        attr = true;
    }
    dim3 grid((g.M + tc::TC_BM - 1) / tc::TC_BM, g.N / BN);
    tc::tc_gemm_nt_kernel<BN><<<grid, 192, tc::NtCfg<BN>::SMEM, st>>>(tmA, tmB, g.C, g.bias, g.M, g.N, g.K, g.ldc, g.relu);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

inline int tc_gemm_nt(const TcGemmNT& g, cudaStream_t st) {
    EPC_CHECK_ARG(g.K % tc::TC_BK == 0 && g.K >= tc::TC_BK, "tc_gemm_nt: K=%d must be a multiple of %d", g.K, tc::TC_BK);
    EPC_CHECK_ARG(g.N % g.BN == 0, "tc_gemm_nt: N=%d must be a multiple of BN=%d", g.N, g.BN);
    EPC_CHECK_ARG((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 && g.ldc % 4 == 0,
                  "tc_gemm_nt: operands must be 16-byte aligned");
    if (g.M == 0) return EPC_OK;
    switch (g.BN) {
        case 64: return tc_gemm_nt_launch<64>(g, st);
        case 128: return tc_gemm_nt_launch<128>(g, st);
        case 256: return tc_gemm_nt_launch<256>(g, st);
    }
    set_error("tc_gemm_nt: BN=%d unsupported", g.BN);
    return EPC_EINVAL;
}

}  // namespace epc
