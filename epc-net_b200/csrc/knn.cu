// K1 -- per-cloud kNN graph with the reference's threshold/tie semantics.
// Replaces tf_util.pairwise_distance_mask (utils/tf_util.py:647-666): instead of materialising the
// B x N x N distance matrix and 0/1 mask (2 x 64 MiB per cloud), each cloud is
//   (1) Morton-sorted in shared memory (sort_kernel) so that index-near == space-near, and
//   (2) scanned by warps that keep 8 query rows in registers, 32 candidates per step across lanes,
//       distances on the packed fp32x2 pipe in the *canonical* arithmetic (common.cuh), a
//       warp-distributed sorted top-20 list per row, and exact AABB pruning of 32-point blocks.
// Output (sorted space): nbr [B,N,20] u16, kthd [B,N] (20th smallest d), cnt [B,N] = |{j: d_ij <= kthd_i}|.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace epc {

// ------------------------------------------------------------------------------------------------
// sort_kernel: one CTA per cloud.  key = (30-bit Morton code << 32) | original index  (unique keys
// => deterministic order).  Emits sorted float4 (x,y,z,|p|^2) and perm (sorted pos -> original idx).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// 10-bit cell coordinates -> 30-bit Hilbert index (Skilling's transpose algorithm): consecutive indices are always adjacent
// cells, so a run of 32 sorted points is a connected, compact set (a Morton run can straddle a jump across the whole cloud,
// which blows up its bounding box and with it the kNN block pruning).
__device__ __forceinline__ uint32_t hilbert3(uint32_t x0, uint32_t x1, uint32_t x2) {
    uint32_t X[3] = {x0, x1, x2};
    constexpr uint32_t M = 1u << 9;
#pragma unroll
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) {
                X[0] ^= P;
            } else {
                const uint32_t t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
#pragma unroll
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t;
    X[1] ^= t;
    X[2] ^= t;
    return (spread3(X[0]) << 2) | (spread3(X[1]) << 1) | spread3(X[2]);       // X[0] carries the most significant bit of each triple
}

// Bitonic sort of E*1024 unique 64-bit keys by 1024 threads.  Element e = slot*1024 + tid lives in register k[slot]:
// compare-exchange distances j >= 1024 stay inside the thread, j < 32 are warp shuffles, and only 32 <= j <= 512 go
// through shared memory (one barrier per pass, plus one to spill and one to reload) -- 32 barriers instead of 78 at N=4096.
template <int E>
__device__ __forceinline__ void bitonic_regs(unsigned long long* __restrict__ keys, int tid) {
    unsigned long long k[E];
#pragma unroll
    for (int s = 0; s < E; ++s) k[s] = keys[s * 1024 + tid];
    __syncthreads();
    for (int K = 2; K <= E * 1024; K <<= 1) {
        int j = K >> 1;
        // distances >= 1024: both elements in this thread
#pragma unroll
        for (int js = E >> 1; js > 0; js >>= 1) {
            if (j == js * 1024) {
#pragma unroll
                for (int s = 0; s < E; ++s) {
                    const int t = s ^ js;
                    if (s < t) {
                        const bool up = (((s * 1024 + tid) & K) == 0);
                        const unsigned long long a = k[s], c = k[t];
                        if ((a > c) == up) {
                            k[s] = c;
                            k[t] = a;
                        }
                    }
                }
                j >>= 1;
            }
        }
        // distances 512..32: through shared memory
        if (j >= 32) {
#pragma unroll
            for (int s = 0; s < E; ++s) keys[s * 1024 + tid] = k[s];
            __syncthreads();
            for (; j >= 32; j >>= 1) {
#pragma unroll
                for (int s = 0; s < E; ++s) {
                    const int i = s * 1024 + tid, ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = keys[i], c = keys[ixj];
                        const bool up = ((i & K) == 0);
                        if ((a > c) == up) {
                            keys[i] = c;
                            keys[ixj] = a;
                        }
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int s = 0; s < E; ++s) k[s] = keys[s * 1024 + tid];
        }
        // distances 16..1: warp shuffles
        for (; j > 0; j >>= 1) {
            const bool lower = ((tid & j) == 0);
#pragma unroll
            for (int s = 0; s < E; ++s) {
                const bool up = (((s * 1024 + tid) & K) == 0);
                const unsigned long long o = __shfl_xor_sync(FULL, k[s], j);
                const bool take_min = (lower == up);
                k[s] = take_min ? (o < k[s] ? o : k[s]) : (o > k[s] ? o : k[s]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < E; ++s) keys[s * 1024 + tid] = k[s];
    __syncthreads();
}

__global__ void __launch_bounds__(1024) sort_kernel(const float* __restrict__ xyz, int N, int NP, int curve,
                                                     float4* __restrict__ sorted, int* __restrict__ perm,
                                                     uint16_t* __restrict__ perm16, float4* __restrict__ aabb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    __shared__ float red[6][32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* p = xyz + (size_t)b * N * 3;

    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < N; i += blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v = p[3 * i + a];
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = warp_min(lo[a]);
        hi[a] = warp_max(hi[a]);
        if (lane == 0) {
            red[a][wid] = lo[a];
            red[3 + a][wid] = hi[a];
        }
    }
    __syncthreads();
    const int nwarp = blockDim.x >> 5;
    float scale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float l = INFINITY, h = -INFINITY;
        for (int w = 0; w < nwarp; ++w) {
            l = fminf(l, red[a][w]);
            h = fmaxf(h, red[3 + a][w]);
        }
        lo[a] = l;
        float ext = h - l;
        scale[a] = (ext > 0.f) ? 1023.0f / ext : 0.f;
    }
    for (int i = tid; i < NP; i += blockDim.x) {
        unsigned long long k = ~0ull;
        if (i < N) {
            uint32_t q[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float f = (p[3 * i + a] - lo[a]) * scale[a];
                f = fminf(fmaxf(f, 0.f), 1023.f);   // NaN -> 0
                q[a] = (uint32_t)f;
            }
            const uint32_t code = curve ? hilbert3(q[0], q[1], q[2]) : (spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2));
            k = ((unsigned long long)code << 32) | (unsigned)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    // bitonic sort, ascending
    if (NP >= 1024) {
        if (NP == 1024) bitonic_regs<1>(keys, tid);
        else if (NP == 2048) bitonic_regs<2>(keys, tid);
        else if (NP == 4096) bitonic_regs<4>(keys, tid);
        else bitonic_regs<8>(keys, tid);
    } else {
        for (int k = 2; k <= NP; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < NP; i += blockDim.x) {
                    int ixj = i ^ j;
                    if (ixj > i) {
                        unsigned long long a = keys[i], c = keys[ixj];
                        bool up = ((i & k) == 0);
                        if ((a > c) == up) {
                            keys[i] = c;
                            keys[ixj] = a;
                        }
                    }
                }
                __syncthreads();
            }
        }
    }
    for (int i = tid; i < N; i += blockDim.x) {
        int src = (int)(keys[i] & 0xffffffffu);
        float x = p[3 * src], y = p[3 * src + 1], z = p[3 * src + 2];
        const float sq = canon_sq(x, y, z);
        sorted[(size_t)b * N + i] = make_float4(x, y, z, sq);
        perm[(size_t)b * N + i] = src;
        perm16[(size_t)b * N + i] = (uint16_t)src;      // N <= 8192: the copy the kNN kernel stages in shared memory
        // axis-aligned box (and max |p|^2) of every 32-point block of the sorted order: the kNN kernel's pruning test
        const int nblk = N >> 5;                   // N % 32 == 0 and i ascends by blockDim (a multiple of 32):
        const float lx = warp_min(x), ly = warp_min(y), lz = warp_min(z);      // each warp holds exactly one block
        const float hx = warp_max(x), hy = warp_max(y), hz = warp_max(z);
        const float sm = warp_max(sq);
        if (lane == 0) {
            aabb[(size_t)b * nblk * 2 + (i >> 5)] = make_float4(lx, ly, lz, sm);
            aabb[(size_t)b * nblk * 2 + nblk + (i >> 5)] = make_float4(hx, hy, hz, 0.f);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// kNN graph (round 2).  A CTA stages its whole cloud in shared memory once and serves 128 query rows; a warp owns 8
// consecutive rows of the Morton order, and FOUR lanes share a row, each taking every fourth candidate pair of a 32-point
// block (lane = row + 8 * quarter).  Pruning works on 8-row groups (a third of the pairs a 32-row group would evaluate),
// every lane still runs the same straight-line code, and there is no shuffle or ballot per candidate:
//
//   knn_bound_kernel    pass A: a valid, tight UPPER BOUND U_i of the row's 20th smallest distance.  Cheap arithmetic
//                       (3 FFMA2 + FADD2 per candidate pair); candidates below the row's running threshold are parked in a
//                       per-lane shared-memory buffer and merged, two at a time, into a short sorted register list by a
//                       branch-free min/max network (M[i] = min3(L[i], max(L[i-1],c1), max(L[i-2],c2)): FMNMX/FMNMX3 only).
//                       Running threshold = max over the row's four lanes of their 5th smallest value (>= 20 candidates lie
//                       below it); final U = the exact 20th smallest of the union of the four lists (a bitonic merge across the
//                       four lanes).  Any 20 candidates give a valid bound, so neither the pruning nor the arithmetic of this
//                       pass can affect the result -- only how tight U is.
//   knn_collect_kernel  pass B: every block whose rigorous lower bound is <= U_i is scanned with the same cheap arithmetic
//                       and the filter d' <= U_i + slack; the canonical arithmetic (common.cuh) is spent on the 20-odd listed
//                       candidates only: thread-per-row, the exact 20th distance is selected among them, then the thresholded
//                       set {j : d_ij <= kth_i} is written out: 20 listed neighbours (out-of-tile first), the row's count, and
//                       the members beyond 20 in the cloud's tie list.
//   knn_slow_kernel     rows whose candidate list overflowed (mass ties: quantised / duplicated / all-zero clouds):
//                       warp-per-row exact radix select over the whole cloud.  Launched on the overflow list only.
//   knn_public_kernel   API outputs in original point order: idx in tf.nn.top_k order (d ascending, ties -> lower original
//                       index), kth = -d20, count.
// Output (sorted space): nbr [B,N,20] u16, kthd [B,N] (20th smallest d), cnt [B,N] = |{j: d_ij <= kthd_i}| | n_out << 24.
// ------------------------------------------------------------------------------------------------
constexpr int KNN_THREADS = 512;      // 16 warps x 8 rows
constexpr int KNN_G = 8;              // rows per warp; 32 / KNN_G = 4 lanes per row
constexpr int KNN_ROWS = (KNN_THREADS / 32) * KNN_G;      // 128 rows per CTA
constexpr int KNN_BLK = 8;            // candidates a lane takes from one 32-point block (4 pairs)
constexpr int KNN_CAPL = 24;          // pass B: listed candidates per lane (u16 each)
constexpr int KNN_CAPB = 40;          // pass B: listed candidates per row
constexpr int KNN_MAXMASK = 8;        // 32-block masks: N <= 8192
constexpr int KNN_WROW = 2 * KNN_G + KNN_MAXMASK / 4 + 4;  // float4 per warp: 8 rows x 2, the block masks, the live-tile list (64 B)

struct KnnLayout {
    int off_lo, off_hi, off_tlo, off_thi, off_misc, off_row, off_aux, off_buf;
    int bytes;
};
// aux_bytes: pass B's per-row lists and per-lane counts
__host__ __device__ inline KnnLayout knn_layout(int N, int buf_bytes_per_lane, int aux_bytes) {
    const int nblk = N >> 5, ntile = (nblk + 3) >> 2;
    KnnLayout L;
    int o = N * 16;                   // points, pair-SoA: (x0,x1,y0,y1) (z0,z1,s0,s1) per pair of points
    L.off_lo = o;  o += nblk * 16;    // block boxes: (lo.xyz, max |p|^2)
    L.off_hi = o;  o += nblk * 16;    //              (hi.xyz, -)
    L.off_tlo = o; o += ntile * 16;   // 128-point tile boxes
    L.off_thi = o; o += ntile * 16;
    L.off_misc = o; o += 16;          // max |p|^2 over the cloud
    L.off_row = o; o += (KNN_THREADS / 32) * KNN_WROW * 16;   // per warp: 8 rows x {(x, y, z, threshold), (|p|^2,-,-,-)} for the
                                                              // lane-parallel box tests, then the warp's block masks
    L.off_aux = o; o += aux_bytes;
    L.off_buf = o; o += KNN_THREADS * buf_bytes_per_lane;
    L.bytes = o;
    return L;
}

// the whole cloud -> shared memory (pair-SoA so that a candidate pair is two LDS.128 whose halves are FFMA2 operands)
__device__ __forceinline__ void knn_stage(const float4* __restrict__ pts, const float4* __restrict__ box, int N,
                                          unsigned char* smem, const KnnLayout& lay) {
    const int nblk = N >> 5, ntile = (nblk + 3) >> 2;
    float4* sp = reinterpret_cast<float4*>(smem);
    for (int pr = threadIdx.x; pr < (N >> 1); pr += blockDim.x) {
        const float4 a = __ldg(pts + 2 * pr), c = __ldg(pts + 2 * pr + 1);
        sp[2 * pr] = make_float4(a.x, c.x, a.y, c.y);
        sp[2 * pr + 1] = make_float4(a.z, c.z, a.w, c.w);
    }
    float4* slo = reinterpret_cast<float4*>(smem + lay.off_lo);
    float4* shi = reinterpret_cast<float4*>(smem + lay.off_hi);
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
        slo[i] = __ldg(box + i);
        shi[i] = __ldg(box + nblk + i);
    }
    __syncthreads();
    float4* stlo = reinterpret_cast<float4*>(smem + lay.off_tlo);
    float4* sthi = reinterpret_cast<float4*>(smem + lay.off_thi);
    for (int t = threadIdx.x; t < ntile; t += blockDim.x) {
        float4 lo = slo[4 * t], hi = shi[4 * t];
        for (int k = 1; k < 4 && 4 * t + k < nblk; ++k) {
            const float4 l2 = slo[4 * t + k], h2 = shi[4 * t + k];
            lo = make_float4(fminf(lo.x, l2.x), fminf(lo.y, l2.y), fminf(lo.z, l2.z), fmaxf(lo.w, l2.w));
            hi = make_float4(fmaxf(hi.x, h2.x), fmaxf(hi.y, h2.y), fmaxf(hi.z, h2.z), 0.f);
        }
        stlo[t] = lo;
        sthi[t] = hi;
    }
    if (threadIdx.x < 32) {           // max |p|^2 over the cloud (the slack of the cheap arithmetic)
        float m = 0.f;
        for (int i = threadIdx.x; i < nblk; i += 32) m = fmaxf(m, slo[i].w);
        m = warp_max(m);
        if (threadIdx.x == 0) *reinterpret_cast<float*>(smem + lay.off_misc) = m;
    }
    __syncthreads();
}

// squared distance from (x,y,z) to the box [lo,hi]
__device__ __forceinline__ float box_lb(float x, float y, float z, const float4 lo, const float4 hi) {
    const float dx = fmaxf(fmaxf(lo.x - x, x - hi.x), 0.f);
    const float dy = fmaxf(fmaxf(lo.y - y, y - hi.y), 0.f);
    const float dz = fmaxf(fmaxf(lo.z - z, z - hi.z), 0.f);
    return dx * dx + dy * dy + dz * dz;
}
// rigorous lower bound of the *computed* canonical d over a box (DESIGN.md "pruning"): true |p-q|^2 >= lb_true >= lb(1-8u);
// computed d >= true - 16u (s_i + s_j)
__device__ __forceinline__ float box_bound_rigorous(float x, float y, float z, float s, const float4 lo, const float4 hi) {
    return box_lb(x, y, z, lo, hi) * (1.0f - 1e-6f) - 1e-6f * (s + lo.w);
}

// the per-lane buffers are addressed through 32-bit shared-space addresses (one register; a generic pointer costs ptxas three)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}
// d' = s_i + s_j - 2 p_i.p_j for a pair of candidates as one FFMA2 chain (q2* = -2 * query coordinate): within
// 1e-6 (s_i + s_j) of the canonical d of either arithmetic
__device__ __forceinline__ float2 fast_dist2(float2 q2x, float2 q2y, float2 q2z, float2 qs, const float4 A, const float4 Bv) {
    float2 t = __ffma2_rn(q2x, make_float2(A.x, A.y), qs);
    t = __ffma2_rn(q2y, make_float2(A.z, A.w), t);
    t = __ffma2_rn(q2z, make_float2(Bv.x, Bv.y), t);
    return __fadd2_rn(t, make_float2(Bv.z, Bv.w));
}

// sorted ascending list L[0..n) <- the n smallest of L u {c1, c2}: M[i] = min(L[i], max(L[i-1], lo), max(L[i-2], hi))
template <int LEN>
__device__ __forceinline__ void merge2(float (&L)[LEN], float c1, float c2) {
    const float lo = fminf(c1, c2), hi = fmaxf(c1, c2);
#pragma unroll
    for (int i = LEN - 1; i >= 2; --i) L[i] = fminf(fminf(L[i], fmaxf(L[i - 1], lo)), fmaxf(L[i - 2], hi));
    L[1] = fminf(fminf(L[1], fmaxf(L[0], lo)), hi);
    L[0] = fminf(L[0], lo);
}

// Lane-parallel box tests of a warp's 8 rows: which 128-point tiles, then which 32-point blocks can hold a candidate below
// the rows' thresholds.  Tiles are tested against the GROUP (the 8 rows' bounding box and their largest threshold: one test
// per lane instead of eight); the blocks of the surviving tiles are enumerated densely over the lanes and tested per row.
// RIGOROUS: the bound of pass B (canonical arithmetic); otherwise the plain box distance (pass A, where pruning only affects
// how tight the bound gets).  srow: the warp's 8 x {(x,y,z,threshold), (|p|^2,...)}; bmask / tlist: the warp's shared scratch.
template <bool RIGOROUS>
__device__ __forceinline__ void knn_block_masks(const float4* __restrict__ srow, const float4* __restrict__ slo,
                                                const float4* __restrict__ shi, const float4* __restrict__ stlo,
                                                const float4* __restrict__ sthi, int nblk, int ntile, int b0, bool prune,
                                                int lane, uint32_t* __restrict__ bmask, unsigned char* __restrict__ tlist) {
    auto rows_need = [&](const float4 lo, const float4 hi) {
        bool need = false;
#pragma unroll
        for (int rr = 0; rr < KNN_G; ++rr) {
            const float4 q = srow[2 * rr];
            if (RIGOROUS)
                need |= !(box_bound_rigorous(q.x, q.y, q.z, srow[2 * rr + 1].x, lo, hi) > q.w);
            else
                need |= !(box_lb(q.x, q.y, q.z, lo, hi) > q.w);
        }
        return need;
    };
    // the group: bounding box of the 8 rows, largest threshold, largest |p|^2
    float gx0, gy0, gz0, gx1, gy1, gz1, gthr, gs;
    {
        const float4 q = srow[2 * (lane & (KNN_G - 1))];
        gx0 = gx1 = q.x; gy0 = gy1 = q.y; gz0 = gz1 = q.z; gthr = q.w;
        gs = srow[2 * (lane & (KNN_G - 1)) + 1].x;
#pragma unroll
        for (int o = 1; o < KNN_G; o <<= 1) {
            gx0 = fminf(gx0, __shfl_xor_sync(FULL, gx0, o)); gx1 = fmaxf(gx1, __shfl_xor_sync(FULL, gx1, o));
            gy0 = fminf(gy0, __shfl_xor_sync(FULL, gy0, o)); gy1 = fmaxf(gy1, __shfl_xor_sync(FULL, gy1, o));
            gz0 = fminf(gz0, __shfl_xor_sync(FULL, gz0, o)); gz1 = fmaxf(gz1, __shfl_xor_sync(FULL, gz1, o));
            gthr = fmaxf(gthr, __shfl_xor_sync(FULL, gthr, o));
            gs = fmaxf(gs, __shfl_xor_sync(FULL, gs, o));
        }
    }
    if (lane < KNN_MAXMASK) bmask[lane] = 0;
    int nalive = 0;
#pragma unroll 1
    for (int t0 = 0; t0 < ntile; t0 += 32) {
        const int t = t0 + lane;
        bool alive = false;
        if (t < ntile) {
            const float4 lo = stlo[t], hi = sthi[t];
            // box-to-box distance: every row's point lies in the group box and no row's threshold exceeds gthr
            const float dx = fmaxf(fmaxf(lo.x - gx1, gx0 - hi.x), 0.f);
            const float dy = fmaxf(fmaxf(lo.y - gy1, gy0 - hi.y), 0.f);
            const float dz = fmaxf(fmaxf(lo.z - gz1, gz0 - hi.z), 0.f);
            const float lb = dx * dx + dy * dy + dz * dz;
            const float bound = RIGOROUS ? lb * (1.0f - 1e-6f) - 1e-6f * (gs + lo.w) : lb;
            alive = !prune || !(bound > gthr);
        }
        const unsigned m = __ballot_sync(FULL, alive);
        if (alive) tlist[nalive + __popc(m & ((1u << lane) - 1u))] = (unsigned char)t;
        nalive += __popc(m);
    }
    __syncwarp();
#pragma unroll 1
    for (int e0 = 0; e0 < 4 * nalive; e0 += 32) {
        const int e = e0 + lane;
        if (e < 4 * nalive) {
            const int blk = 4 * (int)tlist[e >> 2] + (e & 3);
            if (blk < nblk && blk != b0 && blk != b0 + 1 && blk != b0 - 1) {
                if (!prune || rows_need(slo[blk], shi[blk])) atomicOr(&bmask[blk >> 5], 1u << (blk & 31));
            }
        }
    }
    __syncwarp();
}

// ---- pass A ---------------------------------------------------------------------------------------------------
// LEN: entries of a lane's sorted list (>= 5).  The running threshold only needs the lanes' 5th values; the longer the
// lists, the more often the final 20th-of-the-union is the exact 20th smallest (8: always when no quarter holds more than
// 8 of the 20 nearest).
template <int LEN>
__global__ void __launch_bounds__(KNN_THREADS, 2)
knn_bound_kernel(const float4* __restrict__ sorted, const float4* __restrict__ aabb, int N, int cap, int prune,
                 float* __restrict__ U) {
    static_assert(LEN >= 5 && LEN <= 8, "list length");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nblk = N >> 5, ntile = (nblk + 3) >> 2;
    const KnnLayout lay = knn_layout(N, cap * 4, 0);
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    knn_stage(sorted + (size_t)b * N, aabb + (size_t)b * nblk * 2, N, smem_raw, lay);
    const int r0 = blockIdx.x * KNN_ROWS + warp * KNN_G;
    if (r0 >= N) return;                                   // whole warps: N % 32 == 0
    const float4* sp = reinterpret_cast<const float4*>(smem_raw);
    const float4* slo = reinterpret_cast<const float4*>(smem_raw + lay.off_lo);
    const float4* shi = reinterpret_cast<const float4*>(smem_raw + lay.off_hi);
    const float4* stlo = reinterpret_cast<const float4*>(smem_raw + lay.off_tlo);
    const float4* sthi = reinterpret_cast<const float4*>(smem_raw + lay.off_thi);
    float4* srow = reinterpret_cast<float4*>(smem_raw + lay.off_row) + warp * KNN_WROW;
    uint32_t* bmask = reinterpret_cast<uint32_t*>(srow + 2 * KNN_G);
    unsigned char* tlist = reinterpret_cast<unsigned char*>(srow + 2 * KNN_G + KNN_MAXMASK / 4);
    const uint32_t bufp = smem_addr(smem_raw + lay.off_buf) + 4u * tid;      // slot i at bufp + i * KNN_THREADS * 4
    constexpr uint32_t SLOT = KNN_THREADS * 4;

    const int rr = lane & (KNN_G - 1), qd = lane >> 3;      // row within the warp, quarter of the candidates
    const int r = r0 + rr;
    float x, y, z, s;
    {
        const float* qb = reinterpret_cast<const float*>(smem_raw) + (r >> 1) * 8 + (r & 1);
        x = qb[0]; y = qb[2]; z = qb[4]; s = qb[6];
    }
    if (qd == 0) {
        srow[2 * rr] = make_float4(x, y, z, INFINITY);
        srow[2 * rr + 1] = make_float4(s, 0.f, 0.f, 0.f);
    }
    const float2 q2x = make_float2(-2.f * x, -2.f * x), q2y = make_float2(-2.f * y, -2.f * y),
                 q2z = make_float2(-2.f * z, -2.f * z), qs2 = make_float2(s, s);

    float L[LEN];
#pragma unroll
    for (int i = 0; i < LEN; ++i) L[i] = INFINITY;
    float thr = INFINITY;
    uint32_t wp = bufp;
    const uint32_t wlim = bufp + (uint32_t)(cap - KNN_BLK) * SLOT;
    const int b0 = r0 >> 5;                                  // the rows' own block (warp-uniform)

    // merge the parked values into the lane's list; row threshold = max of its four lanes' 5th values
    auto compact = [&]() {
        const int n = (int)((wp - bufp) / SLOT);
        const int nmax = __reduce_max_sync(FULL, n);
#pragma unroll 1
        for (int i = 0; i < nmax; i += 2) {
            const float c1 = (i < n) ? lds_f32(bufp + i * SLOT) : INFINITY;
            const float c2 = (i + 1 < n) ? lds_f32(bufp + (i + 1) * SLOT) : INFINITY;
            merge2<LEN>(L, c1, c2);
        }
        wp = bufp;
        float t = L[4];
        t = fmaxf(t, __shfl_xor_sync(FULL, t, 8));
        t = fmaxf(t, __shfl_xor_sync(FULL, t, 16));
        thr = t;
        if (qd == 0) srow[2 * rr].w = t;
    };
    // one 32-point block: this lane's four candidate pairs (pairs qd, qd+4, qd+8, qd+12 of the block)
    auto scan_block = [&](int blk, bool force) {
        const float4* pp = sp + blk * 32 + 2 * qd;
#pragma unroll
        for (int u = 0; u < KNN_BLK / 2; ++u) {
            const float2 t = fast_dist2(q2x, q2y, q2z, qs2, pp[8 * u], pp[8 * u + 1]);
            if (t.x < thr) {
                sts_f32(wp, t.x);
                wp += SLOT;
            }
            if (t.y < thr) {
                sts_f32(wp, t.y);
                wp += SLOT;
            }
        }
        if (__any_sync(FULL, (wp > wlim) || (force && wp != bufp))) compact();
    };

    // the own block (first threshold), its index neighbours (a tight threshold before the box tests), then every block the
    // tests leave
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
        const int blk = (t == 0) ? b0 : (t == 1) ? b0 + 1 : b0 - 1;
        if (blk >= 0 && blk < nblk) scan_block(blk, t != 1);
    }
    __syncwarp();
    knn_block_masks<false>(srow, slo, shi, stlo, sthi, nblk, ntile, b0, prune != 0, lane, bmask, tlist);
#pragma unroll 1
    for (int br = 0; br * 32 < nblk; ++br) {
        uint32_t m = bmask[br];
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            scan_block(br * 32 + bit, false);
        }
    }
    if (__any_sync(FULL, wp != bufp)) compact();

    // U = the 20th smallest of the union of the row's four sorted lists (padded to 8 with +inf): a bitonic merge across the
    // four lanes.  Round 1: lanes (q, q^1) -> sorted 16 (low half on the even quarter); round 2: the 16 largest of the 32 as a
    // bitonic sequence H; its 4th smallest (= rank 19 of the 32) by two half-cleaners and a max.
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (i < LEN) ? L[i] : INFINITY;
    {
        float c[8];
        const bool odd = (qd & 1) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float o = __shfl_xor_sync(FULL, a[7 - i], 8);
            c[i] = odd ? fmaxf(a[i], o) : fminf(a[i], o);
        }
#pragma unroll
        for (int d = 4; d >= 1; d >>= 1)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if ((i & d) == 0) {
                    const float lo = fminf(c[i], c[i + d]), hi = fmaxf(c[i], c[i + d]);
                    c[i] = lo;
                    c[i + d] = hi;
                }
        // pair (0,1) holds S[0..15], pair (2,3) holds T[0..15]; H[i] = max(S[i], T[15-i]): quarter q pairs with quarter 3-q
        float h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = fmaxf(c[i], __shfl_xor_sync(FULL, c[7 - i], 24));
        // quarters 0 and 3 hold H[0..7] (3 reversed), quarters 1 and 2 hold H[8..15]: the 8 smallest of H
        float m8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m8[i] = fminf(h[i], __shfl_xor_sync(FULL, h[i], 8));
        // careful: quarter 0 pairs with quarter 1 (xor 8) -> H[i] with H[8+i]; quarter 3 (H reversed) pairs with quarter 2
        // (H[8..15] reversed): the same 8 values in reverse order.  4 smallest of the 8, then their maximum.
        float u4 = fmaxf(fmaxf(fminf(m8[0], m8[4]), fminf(m8[1], m8[5])), fmaxf(fminf(m8[2], m8[6]), fminf(m8[3], m8[7])));
        thr = u4;
    }
    // d' is within 1e-6 (s_i + s_j) of the canonical d (either arithmetic): the slack keeps U a valid upper bound of the
    // canonical 20th distance
    if (qd == 0) U[(size_t)b * N + r] = thr + 2e-6f * (s + *reinterpret_cast<const float*>(smem_raw + lay.off_misc));
}

// ---- pass B ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tie_append(uint32_t* __restrict__ tie, uint32_t row_pos, uint32_t j, float kth) {
    if (*reinterpret_cast<volatile uint32_t*>(tie) > TIE_CAP) return;       // overflowed already: stop counting
    const uint32_t slot = atomicAdd(tie, 1u);
    if (slot < TIE_CAP) reinterpret_cast<uint2*>(tie + 2)[slot] = make_uint2((row_pos << 16) | j, __float_as_uint(kth));
}

constexpr int KNN_AUX_B = KNN_ROWS * KNN_CAPB * 2;      // pass B: the CTA's row lists, assembled for coalesced stores

template <int ARITH>
__global__ void __launch_bounds__(KNN_THREADS, 2)
knn_collect_kernel(const float4* __restrict__ sorted, const float4* __restrict__ aabb, const float* __restrict__ U, int N,
                   int prune, uint16_t* __restrict__ glist, int* __restrict__ gcount) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nblk = N >> 5, ntile = (nblk + 3) >> 2;
    const KnnLayout lay = knn_layout(N, KNN_CAPL * 2, KNN_AUX_B);
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    knn_stage(sorted + (size_t)b * N, aabb + (size_t)b * nblk * 2, N, smem_raw, lay);
    const int r0 = blockIdx.x * KNN_ROWS + warp * KNN_G;
    if (r0 >= N) return;
    const float4* sp = reinterpret_cast<const float4*>(smem_raw);
    const float* spf = reinterpret_cast<const float*>(smem_raw);
    const float4* slo = reinterpret_cast<const float4*>(smem_raw + lay.off_lo);
    const float4* shi = reinterpret_cast<const float4*>(smem_raw + lay.off_hi);
    const float4* stlo = reinterpret_cast<const float4*>(smem_raw + lay.off_tlo);
    const float4* sthi = reinterpret_cast<const float4*>(smem_raw + lay.off_thi);
    float4* srow = reinterpret_cast<float4*>(smem_raw + lay.off_row) + warp * KNN_WROW;
    uint32_t* bmask = reinterpret_cast<uint32_t*>(srow + 2 * KNN_G);
    unsigned char* tlist = reinterpret_cast<unsigned char*>(srow + 2 * KNN_G + KNN_MAXMASK / 4);
    const uint32_t bufp = smem_addr(smem_raw + lay.off_buf) + 2u * tid;                // slot i at bufp + i * KNN_THREADS * 2
    constexpr uint32_t SLOT = KNN_THREADS * 2;

    // ---- scan: 8 rows per warp, four lanes per row ----------------------------------------------------------------
    const int rr = lane & (KNN_G - 1), qd = lane >> 3;
    const int r = r0 + rr;
    const float x = spf[(r >> 1) * 8 + (r & 1)], y = spf[(r >> 1) * 8 + (r & 1) + 2], z = spf[(r >> 1) * 8 + (r & 1) + 4],
                s = spf[(r >> 1) * 8 + (r & 1) + 6];
    const float2 q2x = make_float2(-2.f * x, -2.f * x), q2y = make_float2(-2.f * y, -2.f * y),
                 q2z = make_float2(-2.f * z, -2.f * z), qs2 = make_float2(s, s);
    const size_t row = (size_t)b * N + r;
    const float Ui = U[row];
    if (qd == 0) {
        srow[2 * rr] = make_float4(x, y, z, Ui);
        srow[2 * rr + 1] = make_float4(s, 0.f, 0.f, 0.f);
    }
    const float smax = *reinterpret_cast<const float*>(smem_raw + lay.off_misc);
    // scan filter: a candidate whose canonical d is <= U_i has d' <= U_i + 1e-6 (s_i + s_j); the canonical arithmetic
    // itself is spent (knn_finalize_kernel) only on the handful that pass
    float Uf = Ui + 2e-6f * (s + smax);
    uint32_t wp = bufp;
    const uint32_t wlim = bufp + (uint32_t)(KNN_CAPL - KNN_BLK) * SLOT;
    bool over = false;
    const int b0 = r0 >> 5;
    __syncwarp();
    knn_block_masks<true>(srow, slo, shi, stlo, sthi, nblk, ntile, b0, prune != 0, lane, bmask, tlist);
    // ascending position order within a lane: b0-1, b0, b0+1 are part of the sweep
#pragma unroll 1
    for (int br = 0; br * 32 < nblk; ++br) {
        uint32_t m = bmask[br];
#pragma unroll
        for (int t = -1; t <= 1; ++t) {
            const int nb = b0 + t;
            if (nb >= 0 && nb < nblk && (nb >> 5) == br) m |= 1u << (nb & 31);
        }
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int blk = br * 32 + bit;
            const float4* pp = sp + blk * 32 + 2 * qd;
            const uint32_t j = (uint32_t)(blk * 32 + 2 * qd);
#pragma unroll
            for (int u = 0; u < KNN_BLK / 2; ++u) {
                const float2 d = fast_dist2(q2x, q2y, q2z, qs2, pp[8 * u], pp[8 * u + 1]);
                if (d.x <= Uf) {
                    sts_u16(wp, j + 8 * u);
                    wp += SLOT;
                }
                if (d.y <= Uf) {
                    sts_u16(wp, j + 8 * u + 1);
                    wp += SLOT;
                }
            }
            if (wp > wlim) {                // mass ties: this row goes to the warp-per-row exact path; stop listing
                over = true;
                Uf = -INFINITY;
                wp = bufp;
            }
        }
    }
    // ---- the row's list = its four lanes' lists, quarter-major (a fixed order) ---------------------------------------
    const int n = (int)((wp - bufp) / SLOT);
    int off = 0, tot = n;
    bool any_over = over;
    {
        const int n1 = __shfl_xor_sync(FULL, n, 8);          // quarter qd ^ 1
        const int s01 = n + n1;                              // pair sum
        const int n23 = __shfl_xor_sync(FULL, s01, 16);      // the other pair's sum
        off = ((qd & 1) ? n1 : 0) + ((qd & 2) ? n23 : 0);
        tot = s01 + n23;
        any_over |= __shfl_xor_sync(FULL, (int)over, 8) != 0;
        any_over |= __shfl_xor_sync(FULL, (int)any_over, 16) != 0;
    }
    const bool fits = !any_over && tot <= KNN_CAPB;
    const int nmax = __reduce_max_sync(FULL, fits ? n : 0);
    // the warp's 8 rows x 80 B are contiguous in glist: assemble them in shared memory, then 16-byte coalesced stores
    // (lane-scattered 2-byte global stores cost a 32-byte sector write each)
    const uint32_t stg = smem_addr(smem_raw + lay.off_aux) + (uint32_t)warp * (KNN_G * KNN_CAPB * 2);
    const uint32_t mine = stg + (uint32_t)(rr * KNN_CAPB + off) * 2u;
#pragma unroll 1
    for (int i = 0; i < nmax; ++i)
        if (fits && i < n) sts_u16(mine + 2u * i, lds_u16(bufp + i * SLOT));
    __syncwarp();
    {
        uint4* dst = reinterpret_cast<uint4*>(glist + ((size_t)b * N + r0) * KNN_CAPB);
        constexpr int CH = KNN_G * KNN_CAPB * 2 / 16;           // 40 chunks of 16 B
        for (int c = lane; c < CH; c += 32) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(stg + 16u * c) : "memory");
            dst[c] = v;
        }
    }
    if (qd == 0) gcount[row] = fits ? tot : -1;
}

// ---- pass B, second half: thread per row (a separate launch: the dependent gather / min-max chains need the latency hiding
// of a full SM of warps, which the scan kernel's two big-shared-memory CTAs cannot give) -- the exact 20th distance among the
// listed candidates in the CANONICAL arithmetic, then the thresholded set
constexpr int KNN_FIN_THREADS = 128;
template <int ARITH>
__global__ void __launch_bounds__(KNN_FIN_THREADS)
knn_finalize_kernel(const float4* __restrict__ sorted, const uint16_t* __restrict__ glist, const int* __restrict__ gcount,
                    int N, long long rows, uint32_t* __restrict__ tie_all, uint16_t* __restrict__ nbr, float* __restrict__ kthd,
                    int* __restrict__ cnt, int* __restrict__ slow) {
    __shared__ uint16_t sj[KNN_CAPB][KNN_FIN_THREADS];         // the row's candidates, [slot][thread]: conflict-free columns
    __shared__ float sd[KNN_CAPB][KNN_FIN_THREADS];            // ... and their canonical distances
    __shared__ __align__(16) uint16_t snbr[KNN_FIN_THREADS * KNN_K];   // the CTA's neighbour rows, written out coalesced
    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * KNN_FIN_THREADS;
    const long long row = row0 + tid;
    const bool live = row < rows;                              // whole warps: rows % 32 == 0
    if (live) {
    const int b = (int)(row / N), r = (int)(row - (long long)b * N);
    const float4* pts = sorted + (size_t)b * N;
    const float4 q = __ldg(pts + r);
    const int nc = gcount[row];
    const bool ok = (nc >= KNN_K) && (nc <= KNN_CAPB);
    const int n = ok ? nc : 0;
    const int nmax = __reduce_max_sync(FULL, n);
    // the list (80 B per row, 16-byte aligned) -> shared memory; then every distance, independent gathers in flight together
    {
        const uint4* g4 = reinterpret_cast<const uint4*>(glist + row * KNN_CAPB);
#pragma unroll
        for (int k = 0; k < KNN_CAPB / 8; ++k) {
            if (8 * k < nmax) {
                const uint4 v = __ldg(g4 + k);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    sj[8 * k + 2 * e][tid] = (uint16_t)(w[e] & 0xffffu);
                    sj[8 * k + 2 * e + 1][tid] = (uint16_t)(w[e] >> 16);
                }
            }
        }
    }
#pragma unroll 4
    for (int i = 0; i < nmax; ++i) {
        if (i < n) {
            const float4 p = __ldg(pts + sj[i][tid]);
            sd[i][tid] = canon_dist<ARITH>(q.x, q.y, q.z, q.w, p.x, p.y, p.z, p.w);
        }
    }
    // the 20th smallest of n values is the (n-19)-th largest: with nmax - 19 <= 8 (no mass ties in this warp) an 8-entry
    // descending list does it at 40 % of the min/max work of the 20-entry ascending one
    float kth;
    if (nmax - (KNN_K - 1) <= 8) {
        float D[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) D[i] = -INFINITY;
#pragma unroll 1
        for (int i = 0; i < nmax; i += 2) {
            const float c1 = (i < n) ? sd[i][tid] : -INFINITY;
            const float c2 = (i + 1 < n) ? sd[i + 1][tid] : -INFINITY;
            const float hi = fmaxf(c1, c2), lo = fminf(c1, c2);
#pragma unroll
            for (int k = 7; k >= 2; --k) D[k] = fmaxf(fmaxf(D[k], fminf(D[k - 1], hi)), fminf(D[k - 2], lo));
            D[1] = fmaxf(fmaxf(D[1], fminf(D[0], hi)), lo);
            D[0] = fmaxf(D[0], hi);
        }
        kth = D[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) kth = (n - KNN_K == k) ? D[k] : kth;
    } else {
        float L[KNN_K];
#pragma unroll
        for (int i = 0; i < KNN_K; ++i) L[i] = INFINITY;
#pragma unroll 1
        for (int i = 0; i < nmax; i += 2) {
            const float c1 = (i < n) ? sd[i][tid] : INFINITY;
            const float c2 = (i + 1 < n) ? sd[i + 1][tid] : INFINITY;
            merge2<KNN_K>(L, c1, c2);
        }
        kth = L[KNN_K - 1];
    }
    int total = 0, n_out = 0, n_in = 0;
    uint32_t* tie = tie_all + (size_t)b * TIE_WORDS;
#pragma unroll 1
    for (int i = 0; i < nmax; ++i) {
        if (i < n && sd[i][tid] <= kth) {
            const int j = sj[i][tid];
            if (total < KNN_K) {
                const bool outside = (j >> 7) != (r >> 7);
                const int pos = outside ? n_out++ : (KNN_K - 1) - n_in++;
                snbr[tid * KNN_K + pos] = (uint16_t)j;
            } else {
                tie_append(tie, (uint32_t)r, (uint32_t)j, kth);
            }
            ++total;
        }
    }
    if (ok) {
        kthd[row] = kth;
        cnt[row] = total | (n_out << 24);
    } else {
        slow[1 + atomicAdd(slow, 1)] = (int)row;           // mass ties (or NaN input): the warp-per-row exact path; it
#pragma unroll                                             // rewrites the row, which here only needs in-range indices
        for (int i = 0; i < KNN_K; ++i) snbr[tid * KNN_K + i] = (uint16_t)r;
    }
    }
    __syncthreads();
    {
        const long long nrows = (rows - row0 < KNN_FIN_THREADS) ? (rows - row0) : KNN_FIN_THREADS;
        const int chunks = (int)(nrows * KNN_K * 2 / 16);       // rows % 32 == 0 -> a whole number of 16-byte chunks
        uint4* dst = reinterpret_cast<uint4*>(nbr + row0 * KNN_K);
        const uint4* src = reinterpret_cast<const uint4*>(snbr);
        for (int c = tid; c < chunks; c += KNN_FIN_THREADS) dst[c] = src[c];
    }
}

// ---- overflow rows: warp per row, exact, any number of ties ---------------------------------------------------------
__device__ __forceinline__ uint32_t ordered_key(float d) {
    const uint32_t u = __float_as_uint(d + 0.0f);           // -0.0 -> +0.0: equal as floats, equal as keys
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int KNN_SLOW_WARPS = 4;
template <int ARITH>
__global__ void __launch_bounds__(KNN_SLOW_WARPS * 32)
knn_slow_kernel(const float4* __restrict__ sorted, const float* __restrict__ U, int N, const int* __restrict__ slow,
                uint32_t* __restrict__ tie_all, uint16_t* __restrict__ nbr, float* __restrict__ kthd, int* __restrict__ cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)wid * N;                        // [N] per warp
    uint16_t* pos = reinterpret_cast<uint16_t*>(smem_raw + (size_t)KNN_SLOW_WARPS * N * 4) + (size_t)wid * N;
    const int count = slow[0];
    for (int e = blockIdx.x * KNN_SLOW_WARPS + wid; e < count; e += gridDim.x * KNN_SLOW_WARPS) {
        const int row = slow[1 + e];
        const int b = row / N, r = row - b * N;
        const float4* pts = sorted + (size_t)b * N;
        const float4 q = pts[r];
        const float Ui = U[row];
        // (1) the candidates with d <= U_i (at least 20 unless the input holds NaN), ascending position
        int n = 0;
        __syncwarp();
        for (int j0 = 0; j0 < N; j0 += 32) {
            const float4 p = __ldg(pts + j0 + lane);
            const float d = canon_dist<ARITH>(q.x, q.y, q.z, q.w, p.x, p.y, p.z, p.w);
            const unsigned m = __ballot_sync(FULL, d <= Ui);
            if (d <= Ui) {
                const int o = n + __popc(m & ((1u << lane) - 1u));
                keys[o] = ordered_key(d);
                pos[o] = (uint16_t)(j0 + lane);
            }
            n += __popc(m);
        }
        __syncwarp();
        uint32_t* tie = tie_all + (size_t)b * TIE_WORDS;
        if (n < KNN_K) {                                    // NaN coordinates: keep every index in range, nothing else is defined
            if (lane < KNN_K) nbr[(size_t)row * KNN_K + lane] = (uint16_t)r;
            if (lane == 0) {
                kthd[row] = INFINITY;
                cnt[row] = KNN_K;
            }
            continue;
        }
        // (2) radix select: the 20th smallest key
        uint32_t prefix = 0;
        int want = KNN_K;
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t hi_mask = (bit == 31) ? 0u : ~((2u << bit) - 1u);
            int c = 0;
            for (int i = lane; i < n; i += 32) {
                const uint32_t kk = keys[i];
                c += ((kk & hi_mask) == prefix && !((kk >> bit) & 1u)) ? 1 : 0;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
            if (c < want) {
                want -= c;
                prefix |= (1u << bit);
            }
        }
        const uint32_t kkey = prefix;
        const float kth = __uint_as_float((kkey & 0x80000000u) ? (kkey & 0x7fffffffu) : ~kkey);      // inverse of ordered_key
        // (3) the thresholded set in ascending position: the first 20 are listed (out-of-tile ones from the front, in-tile
        // ones from the back), the rest go to the cloud's tie list
        int total = 0, n_out = 0, n_in = 0;
        const unsigned lt = (1u << lane) - 1u;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const bool mem = (i < n) && (keys[i] <= kkey);
            const int j = (i < n) ? pos[i] : 0;
            const bool outside = (j >> 7) != (r >> 7);
            const unsigned mm = __ballot_sync(FULL, mem), mo = __ballot_sync(FULL, mem && outside);
            const int lim = max(0, KNN_K - total);                        // members of this step that are still listed
            unsigned sel = mm;
            if (__popc(mm) > lim) sel = lim ? (mm & ((1u << __fns(mm, 0, lim + 1)) - 1u)) : 0u;
            if (mem) {
                if ((sel >> lane) & 1u) {
                    const int p = outside ? n_out + __popc(sel & mo & lt) : (KNN_K - 1) - (n_in + __popc(sel & ~mo & lt));
                    nbr[(size_t)row * KNN_K + p] = (uint16_t)j;
                } else {
                    tie_append(tie, (uint32_t)r, (uint32_t)j, kth);
                }
            }
            n_out += __popc(sel & mo);
            n_in += __popc(sel & ~mo);
            total += __popc(mm);
        }
        if (lane == 0) {
            kthd[row] = kth;
            cnt[row] = total | (n_out << 24);
        }
    }
}

// ---- API outputs in original point order ---------------------------------------------------------------------------
__device__ __forceinline__ bool key_lt(float av, uint32_t ak, float bv, uint32_t bk) {
    return (av < bv) || (av == bv && ak < bk);
}

// full-key ascending bitonic sort, one (v,k) per lane
__device__ __forceinline__ void warp_sort32_keys(float& v, uint32_t& k, int lane) {
#pragma unroll
    for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            const float ov = __shfl_xor_sync(FULL, v, j);
            const uint32_t ok = __shfl_xor_sync(FULL, k, j);
            const bool up = ((lane & kk) == 0);
            const bool lower = ((lane & j) == 0);
            const bool other_less = key_lt(ov, ok, v, k);
            const bool take = (lower == up) ? other_less : (!other_less && !(ov == v && ok == k));
            if (take) {
                v = ov;
                k = ok;
            }
        }
    }
}

constexpr int KNN_PUB_WARPS = 4;
template <int ARITH>
__global__ void __launch_bounds__(KNN_PUB_WARPS * 32)
knn_public_kernel(const float4* __restrict__ sorted, const int* __restrict__ perm, const uint16_t* __restrict__ nbr,
                  const float* __restrict__ kthd, const int* __restrict__ cnt, int N, long long rows,
                  int32_t* __restrict__ idx_out, float* __restrict__ kth_out, int32_t* __restrict__ count_out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * KNN_PUB_WARPS + wid;
    if (row >= rows) return;
    const long long b = row / N;
    const int r = (int)(row - b * N);
    const float4* pts = sorted + b * N;
    const int* pm = perm + b * N;
    const long long orow = b * N + pm[r];
    const float kth = kthd[row];
    const int count = cnt[row] & 0xffffff;
    if (lane == 0) {
        if (kth_out) kth_out[orow] = -kth;
        if (count_out) count_out[orow] = count;
    }
    if (!idx_out) return;
    const float4 q = pts[r];
    if (count == KNN_K) {
        float v = INFINITY;
        uint32_t k = 0xffffffffu;
        if (lane < KNN_K) {
            const int j = nbr[row * KNN_K + lane];
            const float4 p = pts[j];
            v = canon_dist<ARITH>(q.x, q.y, q.z, q.w, p.x, p.y, p.z, p.w);
            k = (uint32_t)pm[j];
        }
        warp_sort32_keys(v, k, lane);
        if (lane < KNN_K) idx_out[orow * KNN_K + lane] = (int32_t)k;
        return;
    }
    // ties beyond the 20th place: the 20 smallest (d, original index) of the whole thresholded set, by repeated extraction
    float last_v = -INFINITY;
    uint32_t last_k = 0;
    bool first = true;
    for (int t = 0; t < KNN_K; ++t) {
        float bv = INFINITY;
        uint32_t bk = 0xffffffffu;
        for (int j0 = 0; j0 < N; j0 += 32) {
            const float4 p = __ldg(pts + j0 + lane);
            const float d = canon_dist<ARITH>(q.x, q.y, q.z, q.w, p.x, p.y, p.z, p.w);
            const uint32_t k = (uint32_t)__ldg(pm + j0 + lane);
            if (d <= kth && (first || key_lt(last_v, last_k, d, k)) && key_lt(d, k, bv, bk)) {
                bv = d;
                bk = k;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL, bv, o);
            const uint32_t ok = __shfl_xor_sync(FULL, bk, o);
            if (key_lt(ov, ok, bv, bk)) {
                bv = ov;
                bk = ok;
            }
        }
        if (lane == 0) idx_out[orow * KNN_K + t] = (int32_t)bk;
        last_v = bv;
        last_k = bk;
        first = false;
    }
}

// Dense exports (API parity): mask_ij = (a_ij >= kth_i), dist_ij = -a_ij, original point order.
template <int ARITH>
__global__ void knn_dense_kernel(const float* __restrict__ xyz, const float* __restrict__ kth, int N,
                                 float* __restrict__ mask, float* __restrict__ dist) {
    const int b = blockIdx.z, i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const float* p = xyz + (size_t)b * N * 3;
    const float qx = p[3 * i], qy = p[3 * i + 1], qz = p[3 * i + 2];
    const float px = p[3 * j], py = p[3 * j + 1], pz = p[3 * j + 2];
    const float d = canon_dist<ARITH>(qx, qy, qz, canon_sq(qx, qy, qz), px, py, pz, canon_sq(px, py, pz));
    const size_t o = ((size_t)b * N + i) * N + j;
    if (dist) dist[o] = d;
    if (mask) mask[o] = ((-d) >= kth[(size_t)b * N + i]) ? 1.0f : 0.0f;
}

// tf_util.knn(adj, k): one warp per row, k smallest, ascending value, ties -> lower column first.
__global__ void rows_topk_smallest_kernel(const float* __restrict__ adj, long long R, int M, int k,
                                          int32_t* __restrict__ idx) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    const float* a = adj + row * (long long)M;
    float val = INFINITY;   // lane l < filled holds the l-th smallest so far
    int vi = 0;
    int filled = 0;         // warp-uniform
    float thr = INFINITY;   // value of entry k-1 once the list is full
    for (int j0 = 0; j0 < M; j0 += 32) {
        const int j = j0 + lane;
        const float d = (j < M) ? a[j] : INFINITY;
        // columns ascend, so once the list is full an equal value loses the tie: strict '<'
        unsigned m = __ballot_sync(FULL, (j < M) && (filled < k || d < thr));
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float c = __shfl_sync(FULL, d, src);
            const bool before = (lane < filled) && (val <= c);
            const int pos = __popc(__ballot_sync(FULL, before));
            if (pos < k) {
                const float upv = __shfl_up_sync(FULL, val, 1);
                const int upi = __shfl_up_sync(FULL, vi, 1);
                if (lane == pos) {
                    val = c;
                    vi = j0 + src;
                } else if (lane > pos && lane < k) {
                    val = upv;
                    vi = upi;
                }
                filled = min(filled + 1, k);
                thr = (filled == k) ? __shfl_sync(FULL, val, k - 1) : INFINITY;
            }
        }
    }
    if (lane < k) idx[row * k + lane] = vi;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int sm_count_knn() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
}

static int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

int knn_check_n(int N) {
    if (N < 32 || N > 8192 || (N % 32) != 0) {
        set_error("N=%d unsupported: need a multiple of 32 in [32, 8192]", N);
        return EPC_EINVAL;
    }
    return EPC_OK;
}

// Tuning knobs (environment; results never depend on them): EPC_KNN_CAP = parked values per row in pass A (even, >= 16,
// default: what lets two CTAs share an SM), EPC_KNN_LEN = 5 | 6 | 8 entries of a lane's sorted list in pass A,
// EPC_SORT_CURVE = 0 (Morton) | 1 (Hilbert, default) order of the points.
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int knn_build(const float* xyz, int B, int N, int arith, bool prune, const KnnState& s, int32_t* idx_out, float* kth_out,
              int32_t* count_out, cudaStream_t st) {
    if (int rc = knn_check_n(N)) return rc;
    EPC_CHECK_ARG(arith == EPC_KNN_ARITH_MULADD || arith == EPC_KNN_ARITH_FMA, "bad knn arith %d", arith);
    if (B == 0) return EPC_OK;
    const int NP = next_pow2(N);
    const size_t sort_smem = (size_t)NP * sizeof(unsigned long long);
    // pass A buffer: as many parked values per row as two resident CTAs per SM allow (one CTA when the cloud is large)
    const int fixed = knn_layout(N, 0, 0).bytes;
    const int budget2 = (227 * 1024) / 2 - 1024, budget1 = 227 * 1024 - 1024;
    int cap = (budget2 - fixed) / (KNN_THREADS * 4) / 2 * 2;
    if (cap < 16) cap = (budget1 - fixed) / (KNN_THREADS * 4) / 2 * 2;
    if (cap > 64) cap = 64;
    {
        const int e = env_int("EPC_KNN_CAP", 0);
        if (e >= 16 && e % 2 == 0 && fixed + e * KNN_THREADS * 4 <= budget1) cap = e;
    }
    EPC_CHECK_ARG(cap >= 16, "kNN: N=%d leaves no shared memory for the candidate buffers", N);
    const int len = env_int("EPC_KNN_LEN", 8);
    const size_t smemA = (size_t)knn_layout(N, cap * 4, 0).bytes, smemB = (size_t)knn_layout(N, KNN_CAPL * 2, KNN_AUX_B).bytes;
    EPC_CHECK_ARG(smemB <= (size_t)budget1, "kNN: N=%d leaves no shared memory for the candidate lists", N);
    const size_t smemC = (size_t)KNN_SLOW_WARPS * N * 6;
    static PerDeviceSize a_sort, a_A5, a_A6, a_A8, a_B0, a_B1, a_C0, a_C1;
    EPC_CUDA(ensure_dyn_smem(sort_kernel, 64 * 1024, a_sort));
    EPC_CUDA(ensure_dyn_smem(knn_bound_kernel<5>, smemA, a_A5));
    EPC_CUDA(ensure_dyn_smem(knn_bound_kernel<6>, smemA, a_A6));
    EPC_CUDA(ensure_dyn_smem(knn_bound_kernel<8>, smemA, a_A8));
    EPC_CUDA(ensure_dyn_smem(knn_collect_kernel<0>, smemB, a_B0));
    EPC_CUDA(ensure_dyn_smem(knn_collect_kernel<1>, smemB, a_B1));
    EPC_CUDA(ensure_dyn_smem(knn_slow_kernel<0>, smemC, a_C0));
    EPC_CUDA(ensure_dyn_smem(knn_slow_kernel<1>, smemC, a_C1));
    {
        ScopedStage ss(EPC_STAGE_SORT, st);
        sort_kernel<<<B, 1024, sort_smem, st>>>(xyz, N, NP, env_int("EPC_SORT_CURVE", 1), s.sorted, s.perm, s.perm16, s.aabb);
        EPC_LAUNCH_CHECK();
    }
    EPC_CUDA(cudaMemsetAsync(s.tie, 0, (size_t)B * TIE_WORDS * sizeof(uint32_t), st));      // counters (and stale entries)
    EPC_CUDA(cudaMemsetAsync(s.slow, 0, sizeof(int), st));                                   // overflow-row counter
    ScopedStage ss(EPC_STAGE_KNN, st);
    dim3 grid((N + KNN_ROWS - 1) / KNN_ROWS, B);
    if (len == 5)
        knn_bound_kernel<5><<<grid, KNN_THREADS, smemA, st>>>(s.sorted, s.aabb, N, cap, prune ? 1 : 0, s.U);
    else if (len == 6)
        knn_bound_kernel<6><<<grid, KNN_THREADS, smemA, st>>>(s.sorted, s.aabb, N, cap, prune ? 1 : 0, s.U);
    else
        knn_bound_kernel<8><<<grid, KNN_THREADS, smemA, st>>>(s.sorted, s.aabb, N, cap, prune ? 1 : 0, s.U);
    EPC_LAUNCH_CHECK();
    const int slow_ctas = 2 * sm_count_knn();
    const long long rows_all = (long long)B * N;
    const unsigned fin_grid = (unsigned)((rows_all + KNN_FIN_THREADS - 1) / KNN_FIN_THREADS);
    if (arith == EPC_KNN_ARITH_MULADD) {
        knn_collect_kernel<0><<<grid, KNN_THREADS, smemB, st>>>(s.sorted, s.aabb, s.U, N, prune ? 1 : 0, s.glist, s.gcount);
        EPC_LAUNCH_CHECK();
        knn_finalize_kernel<0><<<fin_grid, KNN_FIN_THREADS, 0, st>>>(s.sorted, s.glist, s.gcount, N, rows_all, s.tie, s.nbr, s.kthd, s.cnt, s.slow);
        EPC_LAUNCH_CHECK();
        knn_slow_kernel<0><<<slow_ctas, KNN_SLOW_WARPS * 32, smemC, st>>>(s.sorted, s.U, N, s.slow, s.tie, s.nbr, s.kthd, s.cnt);
    } else {
        knn_collect_kernel<1><<<grid, KNN_THREADS, smemB, st>>>(s.sorted, s.aabb, s.U, N, prune ? 1 : 0, s.glist, s.gcount);
        EPC_LAUNCH_CHECK();
        knn_finalize_kernel<1><<<fin_grid, KNN_FIN_THREADS, 0, st>>>(s.sorted, s.glist, s.gcount, N, rows_all, s.tie, s.nbr, s.kthd, s.cnt, s.slow);
        EPC_LAUNCH_CHECK();
        knn_slow_kernel<1><<<slow_ctas, KNN_SLOW_WARPS * 32, smemC, st>>>(s.sorted, s.U, N, s.slow, s.tie, s.nbr, s.kthd, s.cnt);
    }
    EPC_LAUNCH_CHECK();
    if (idx_out || kth_out || count_out) {
        const long long rows = (long long)B * N;
        const unsigned g = (unsigned)((rows + KNN_PUB_WARPS - 1) / KNN_PUB_WARPS);
        if (arith == EPC_KNN_ARITH_MULADD)
            knn_public_kernel<0><<<g, KNN_PUB_WARPS * 32, 0, st>>>(s.sorted, s.perm, s.nbr, s.kthd, s.cnt, N, rows, idx_out, kth_out, count_out);
        else
            knn_public_kernel<1><<<g, KNN_PUB_WARPS * 32, 0, st>>>(s.sorted, s.perm, s.nbr, s.kthd, s.cnt, N, rows, idx_out, kth_out, count_out);
        EPC_LAUNCH_CHECK();
    }
    return EPC_OK;
}

size_t knn_state_bytes(int B, int N) {
    const size_t R = (size_t)B * N;
    return align_up((size_t)B * TIE_WORDS * sizeof(uint32_t)) + align_up(R * sizeof(float4)) + align_up(R * sizeof(int)) + align_up(R * sizeof(uint16_t)) + align_up(R * KNN_K * sizeof(uint16_t)) +
           align_up(R * sizeof(float)) + align_up(R * sizeof(int)) + align_up(R / 16 * sizeof(float4)) + align_up(R * sizeof(float)) + align_up((R + 1) * sizeof(int)) +
           align_up(R * KNN_CAPB * sizeof(uint16_t)) + align_up(R * sizeof(int));
}

KnnState knn_state_carve(Arena& ar, int B, int N) {
    const size_t R = (size_t)B * N;
    KnnState s;
    s.tie = ar.take<uint32_t>((size_t)B * TIE_WORDS);
    s.sorted = ar.take<float4>(R);
    s.perm = ar.take<int>(R);
    s.perm16 = ar.take<uint16_t>(R);
    s.nbr = ar.take<uint16_t>(R * KNN_K);
    s.kthd = ar.take<float>(R);
    s.cnt = ar.take<int>(R);
    s.aabb = ar.take<float4>(R / 16);      // [B][2][N/32]
    s.U = ar.take<float>(R);
    s.slow = ar.take<int>(R + 1);
    s.glist = ar.take<uint16_t>(R * KNN_CAPB);
    s.gcount = ar.take<int>(R);
    return s;
}

int knn_dense(const float* xyz, int B, int N, int arith, const float* kth, float* mask, float* dist, cudaStream_t st) {
    dim3 grid((N + 255) / 256, N, B);
    if (arith == EPC_KNN_ARITH_MULADD)
        knn_dense_kernel<0><<<grid, 256, 0, st>>>(xyz, kth, N, mask, dist);
    else
        knn_dense_kernel<1><<<grid, 256, 0, st>>>(xyz, kth, N, mask, dist);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

int rows_topk_smallest(const float* adj, long long R, int M, int k, int32_t* idx, cudaStream_t st) {
    EPC_CHECK_ARG(k >= 1 && k <= 32 && k <= M, "rows_topk_smallest: k=%d unsupported (1..min(32,M))", k);
    if (R == 0) return EPC_OK;
    const int warps = 8;
    rows_topk_smallest_kernel<<<(unsigned)((R + warps - 1) / warps), warps * 32, 0, st>>>(adj, R, M, k, idx);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

}  // namespace epc
