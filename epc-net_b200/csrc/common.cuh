// Shared helpers for the EPC-Net B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/epc_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "epc_b200 kernels are written for sm_100a (B200) only"
#endif

namespace epc {

constexpr unsigned FULL = 0xffffffffu;
constexpr int KNN_K = 20;            // literal 20 of utils/tf_util.py:660
constexpr float BN_EPS = 1e-3f;      // utils/tf_util.py:490 and slim defaults
constexpr float L2_EPS = 1e-12f;     // tf.nn.l2_normalize

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Optional per-stage device timing (epc_profile_* in the C ABI): CUDA events recorded on the launching stream
// around a stage's kernels.  Costs nothing when profiling is off.
struct ScopedStage {
    int id;
    cudaStream_t st;
    bool on;
    ScopedStage(int stage_id, cudaStream_t stream);
    ~ScopedStage();
};

#define EPC_CHECK_ARG(cond, ...)                         \
    do {                                                 \
        if (!(cond)) {                                   \
            ::epc::set_error(__VA_ARGS__);               \
            return EPC_EINVAL;                           \
        }                                                \
    } while (0)

#define EPC_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            ::epc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return EPC_ECUDA;                                                                       \
        }                                                                                           \
    } while (0)

#define EPC_LAUNCH_CHECK()                                                                          \
    do {                                                                                            \
        ::epc::count_launch();                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                       \
        if (e__ != cudaSuccess) {                                                                   \
            ::epc::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return EPC_ECUDA;                                                                       \
        }                                                                                           \
    } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Function attributes (opt-in dynamic shared memory), SM counts and cluster occupancies are PER DEVICE, and one process may
// drive several devices (Engine(device=...), ensure_device()): every such cache is an array indexed by the current device,
// with atomics because the C ABI is called from threads that run with the GIL released.
constexpr int EPC_MAX_DEVICES = 64;
inline int current_device_slot() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0) d = 0;
    return d % EPC_MAX_DEVICES;
}
struct PerDeviceSize {
    std::atomic<size_t> v[EPC_MAX_DEVICES];
    PerDeviceSize() { for (auto& x : v) x.store(0, std::memory_order_relaxed); }
};
// raise cudaFuncAttributeMaxDynamicSharedMemorySize of `kern` on the current device to at least `bytes` (idempotent, cheap)
template <typename K>
inline cudaError_t ensure_dyn_smem(K kern, size_t bytes, PerDeviceSize& cache) {
    std::atomic<size_t>& c = cache.v[current_device_slot()];
    if (bytes <= c.load(std::memory_order_acquire)) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) c.store(bytes, std::memory_order_release);
    return e;
}

// Bump allocator over the caller's workspace.
struct Arena {
    char* base;
    size_t cap;
    size_t off = 0;
    Arena(void* p, size_t c) : base(static_cast<char*>(p)), cap(c) {}
    template <typename T>
    T* take(size_t n) {
        size_t bytes = align_up(n * sizeof(T));
        T* r = reinterpret_cast<T*>(base + off);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= cap; }
};

// ------------------------------------------------------------------------------------------------
// Canonical fp32 evaluation of d_ij = (s_i + (-2 p_i.p_j)) + s_j   (a_ij = -d_ij)
// utils/tf_util.py:651-656; see include/epc_b200.h EPC_KNN_ARITH_*.
// fma(-2, inner, s_i) == fadd(s_i, fmul(-2, inner)) bit for bit because -2*inner is exact.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float canon_sq(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

template <int ARITH>
__device__ __forceinline__ float canon_dist(float qx, float qy, float qz, float qs, float px, float py, float pz,
                                            float ps) {
    float inner;
    if (ARITH == EPC_KNN_ARITH_MULADD) {
        inner = __fadd_rn(__fadd_rn(__fmul_rn(qx, px), __fmul_rn(qy, py)), __fmul_rn(qz, pz));
    } else {
        inner = __fmaf_rn(qz, pz, __fmaf_rn(qy, py, __fmul_rn(qx, px)));
    }
    return __fadd_rn(__fmaf_rn(-2.0f, inner, qs), ps);
}

// two query rows at once on the packed fp32x2 pipe (FMUL2/FADD2/FFMA2 are IEEE-RN per element)
template <int ARITH>
__device__ __forceinline__ float2 canon_dist2(float2 qx, float2 qy, float2 qz, float2 qs, float2 px, float2 py,
                                              float2 pz, float2 ps) {
    float2 inner;
    if (ARITH == EPC_KNN_ARITH_MULADD) {
        // ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with explicit .rn and
        // -fmad=false (verified in SASS), which would silently turn this mode into the FMA one.  It does
        // NOT fuse a packed multiply with a *scalar* add, so the two sums stay scalar.
        // tests/test_build.py::test_sass_muladd_not_contracted guards this.
        const float2 a = __fmul2_rn(qx, px), b = __fmul2_rn(qy, py), c = __fmul2_rn(qz, pz);
        inner.x = __fadd_rn(__fadd_rn(a.x, b.x), c.x);
        inner.y = __fadd_rn(__fadd_rn(a.y, b.y), c.y);
    } else {
        inner = __ffma2_rn(qz, pz, __ffma2_rn(qy, py, __fmul2_rn(qx, px)));
    }
    return __fadd2_rn(__ffma2_rn(make_float2(-2.0f, -2.0f), inner, qs), ps);
}

// round to the nearest TF32 (10-bit mantissa): kind::tf32 MMAs ignore the low 13 mantissa bits, i.e. truncate;
// rounding the operand when it is produced removes that bias
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

}  // namespace epc
