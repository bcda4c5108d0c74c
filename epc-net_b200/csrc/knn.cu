// K1 -- per-cloud kNN graph with the reference's threshold/tie semantics.
// Replaces tf_util.pairwise_distance_mask (utils/tf_util.py:647-666): instead of materialising the
// B x N x N distance matrix and 0/1 mask (2 x 64 MiB per cloud), each cloud is
//   (1) Morton-sorted in shared memory (sort_kernel) so that index-near == space-near, and
//   (2) scanned by warps that keep 8 query rows in registers, 32 candidates per step across lanes,
//       distances on the packed fp32x2 pipe in the *canonical* arithmetic (common.cuh), a
//       warp-distributed sorted top-20 list per row, and exact AABB pruning of 32-point blocks.
// Output (sorted space): nbr [B,N,20] u16, kthd [B,N] (20th smallest d), cnt [B,N] = |{j: d_ij <= kthd_i}|.
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

__global__ void __launch_bounds__(1024) sort_kernel(const float* __restrict__ xyz, int N, int NP,
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
            uint32_t code = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
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
// Per-row candidate buffer in shared memory: up to 64 (value, key) pairs; after a compaction the first 20 are the
// current best set (unsorted) and `thr` is the 20th smallest value.  A block's hits are *appended* with one ballot,
// one popc and a predicated store, however many there are.  A *compaction* selects the 20th smallest VALUE with a
// value-only bitonic network (one shuffle + one min/max per step), then keeps the entries below it (plus the ties
// needed, by lowest original index) with a ballot-rank scatter.  Only the final 20 are sorted by the full key
// (d, original index) = tf.nn.top_k order.  `extra` counts dropped candidates equal to the 20th value (the size of
// the thresholded set minus 20).
// ------------------------------------------------------------------------------------------------
struct RowState {
    float thr;           // 20th smallest value at the last compaction (warp-uniform)
    int n;               // entries in the buffer (warp-uniform)
    int extra;           // # dropped candidates with value == thr (warp-uniform)
};
constexpr uint32_t KEY_EMPTY = 0xffffffffu;
constexpr int KNN_CAP = 64;      // buffer entries per row = two per lane during a compaction
constexpr int KNN_ROWS_PER_WARP = 8;
constexpr int KNN_WARPS = 8;
constexpr int KNN_ROWS_PER_CTA = KNN_ROWS_PER_WARP * KNN_WARPS;
constexpr size_t KNN_BUF_BYTES = (size_t)KNN_ROWS_PER_CTA * KNN_CAP * 8;     // candidate values + keys: the first bytes of the dynamic smem

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

// value-only bitonic sort of one float per lane (ASC or descending)
template <bool ASC>
__device__ __forceinline__ float warp_sort32_vals(float v, int lane) {
#pragma unroll
    for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            const float o = __shfl_xor_sync(FULL, v, j);
            const bool keep_min = (((lane & j) == 0) == ((((lane & kk) == 0)) == ASC));
            v = keep_min ? fminf(v, o) : fmaxf(v, o);
        }
    }
    return v;
}

// buffer [0,n) -> its 20 smallest keys in [0,20) (unsorted), thr, extra.   NOT inlined, state by value: one copy of
// the networks (inlining them per row and call site blew the instruction cache: 40 no-instruction stalls per issue).
__device__ __noinline__ RowState compact(RowState R, float* __restrict__ bv, uint32_t* __restrict__ bk) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    const float o0 = (lane < R.n) ? bv[lane] : INFINITY;
    const float o1 = (lane + 32 < R.n) ? bv[lane + 32] : INFINITY;
    const uint32_t k0 = (lane < R.n) ? bk[lane] : KEY_EMPTY;
    const uint32_t k1 = (lane + 32 < R.n) ? bk[lane + 32] : KEY_EMPTY;
    // two bitonic networks in lock step (o0 ascending, o1 descending): two independent shuffle chains in flight
    float s = o0, s1 = o1;
#pragma unroll
    for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            const bool keep_min = (((lane & j) == 0) == ((lane & kk) == 0));
            const float oa = __shfl_xor_sync(FULL, s, j), od = __shfl_xor_sync(FULL, s1, j);
            s = keep_min ? fminf(s, oa) : fmaxf(s, oa);
            s1 = keep_min ? fmaxf(s1, od) : fminf(s1, od);
        }
    }
    s = fminf(s, s1);                               // bitonic sequence holding the 32 smallest values
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const float o = __shfl_xor_sync(FULL, s, j);
        s = ((lane & j) == 0) ? fminf(s, o) : fmaxf(s, o);
    }
    const float thr_new = __shfl_sync(FULL, s, KNN_K - 1);
    bool keep0 = o0 < thr_new, keep1 = o1 < thr_new;
    const bool eq0 = (o0 == thr_new), eq1 = (o1 == thr_new);
    const int c_less = __popc(__ballot_sync(FULL, keep0)) + __popc(__ballot_sync(FULL, keep1));
    const unsigned me0 = __ballot_sync(FULL, eq0), me1 = __ballot_sync(FULL, eq1);
    const int c_eq = __popc(me0) + __popc(me1);
    const int need_eq = KNN_K - c_less;            // >= 1
    if (c_eq == need_eq) {
        keep0 |= eq0;
        keep1 |= eq1;
    } else {
        // ties straddle the 20th place: keep the need_eq tied entries with the lowest original index (rare)
        bool t0 = eq0, t1 = eq1;
        for (int i = 0; i < need_eq; ++i) {
            uint32_t best = min(t0 ? k0 : KEY_EMPTY, t1 ? k1 : KEY_EMPTY);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(FULL, best, o));
            if (t0 && k0 == best) { t0 = false; keep0 = true; }
            if (t1 && k1 == best) { t1 = false; keep1 = true; }
        }
        // The dropped tied candidates belong to the row's thresholded set if thr_new turns out to be its final threshold:
        // log them (row, position, value) in the cloud's tie list so that the ProxyConv gather can add them without
        // re-scanning the cloud.  An overflowing list (degenerate clouds) makes the gather fall back to the re-scan.
        // (row and list are recovered from the buffer address: the candidate buffers open the dynamic shared memory and
        // the kernel leaves the list pointer right behind them -- extra arguments would cost this call's ABI registers)
        extern __shared__ __align__(16) unsigned char smem_raw[];
        const uint32_t row_pos = blockIdx.x * KNN_ROWS_PER_CTA + (uint32_t)((bv - reinterpret_cast<float*>(smem_raw)) / KNN_CAP);
        uint32_t* tie = *reinterpret_cast<uint32_t**>(smem_raw + KNN_BUF_BYTES);
        const bool room = *reinterpret_cast<volatile uint32_t*>(tie) <= TIE_CAP;      // stop counting once the list has overflowed
        if (t0 && room) {
            const uint32_t slot = atomicAdd(tie, 1u);
            if (slot < TIE_CAP) reinterpret_cast<uint2*>(tie + 2)[slot] = make_uint2((row_pos << 16) | (k0 & 0xffffu), __float_as_uint(o0));
        }
        if (t1 && room) {
            const uint32_t slot = atomicAdd(tie, 1u);
            if (slot < TIE_CAP) reinterpret_cast<uint2*>(tie + 2)[slot] = make_uint2((row_pos << 16) | (k1 & 0xffffu), __float_as_uint(o1));
        }
    }
    R.extra = ((thr_new == R.thr) ? R.extra : 0) + (c_eq - need_eq);
    R.thr = thr_new;
    R.n = KNN_K;
    const unsigned mk0 = __ballot_sync(FULL, keep0), mk1 = __ballot_sync(FULL, keep1);
    const unsigned lt = (1u << lane) - 1u;
    __syncwarp();                                   // every lane has read its entries: safe to overwrite [0,20)
    if (keep0) {
        const int e = __popc(mk0 & lt);
        bv[e] = o0;
        bk[e] = k0;
    }
    if (keep1) {
        const int e = __popc(mk0) + __popc(mk1 & lt);
        bv[e] = o1;
        bk[e] = k1;
    }
    __syncwarp();
    return R;
}

// the row's 20 survivors -> lane l gets the key of the l-th smallest (d, original index).  Not inlined (code size).
__device__ __noinline__ uint32_t sorted_key(const float* __restrict__ rv, const uint32_t* __restrict__ rk) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    float v = (lane < KNN_K) ? rv[lane] : INFINITY;
    uint32_t k = (lane < KNN_K) ? rk[lane] : KEY_EMPTY;
    warp_sort32_keys(v, k, lane);
    return k;
}

// queue the candidates flagged in m (lane s holds value d for sorted position jbase + s)
__device__ __forceinline__ void append_hits(RowState& R, unsigned m, float d, int jbase, const unsigned short* __restrict__ sperm,
                                            float* __restrict__ bv, uint32_t* __restrict__ bk, int lane) {
    int h = __popc(m);
    if (R.n + h > KNN_CAP) {
        R = compact(R, bv, bk);
        m &= __ballot_sync(FULL, d <= R.thr);      // the bound just tightened
        h = __popc(m);
        if (h == 0) return;
    }
    if ((m >> lane) & 1u) {
        const int e = R.n + __popc(m & ((1u << lane) - 1u));
        bv[e] = d;
        bk[e] = ((uint32_t)sperm[jbase + lane] << 16) | (uint32_t)(jbase + lane);
    }
    R.n += h;
}

// a row is re-compacted after the index-neighbour blocks only if its buffer holds more than this many candidates
// (measured: 20 -> 9.11, 32 -> 8.79, 44 -> 8.85, 64 = never -> 8.97 us/cloud)
constexpr int KNN_STAGE1_MIN = 32;

template <int ARITH, bool PRUNE>
__global__ void __launch_bounds__(KNN_WARPS * 32, 2)
knn_kernel(const float4* __restrict__ sorted, const uint16_t* __restrict__ perm16, const float4* __restrict__ aabb,
           uint32_t* __restrict__ tie_all, int N,
           uint16_t* __restrict__ nbr, float* __restrict__ kthd, int* __restrict__ cnt, int32_t* __restrict__ idx_out,
           float* __restrict__ kth_out, int32_t* __restrict__ count_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nblk = N >> 5;
    float* sbv = reinterpret_cast<float*>(smem_raw);                          // [warps][rows][CAP] candidate values (compact() relies on offset 0)
    uint32_t* sbk = reinterpret_cast<uint32_t*>(sbv + KNN_WARPS * KNN_ROWS_PER_WARP * KNN_CAP);   // ... and keys
    uint32_t** stie = reinterpret_cast<uint32_t**>(smem_raw + KNN_BUF_BYTES);                     // this cloud's tie list (for compact())
    uint64_t* ldbar = reinterpret_cast<uint64_t*>(stie + 1);
    float4* spts = reinterpret_cast<float4*>(smem_raw + KNN_BUF_BYTES + 16);   // [N]   (x,y,z,s)
    float4* sblo = spts + N;                              // [nblk] (lo.xyz, max s)
    float4* sbhi = sblo + nblk;                           // [nblk] (hi.xyz, -)
    unsigned short* sperm = reinterpret_cast<unsigned short*>(sbhi + nblk);   // [N] sorted position -> original index
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // the whole cloud (points, block boxes, permutation) arrives by three TMA bulk copies issued by one thread
    {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(ldbar);
        if (tid == 0) {
            *stie = tie_all + (size_t)blockIdx.y * TIE_WORDS;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t b_pts = (uint32_t)N * 16u, b_box = (uint32_t)nblk * 32u, b_perm = (uint32_t)N * 2u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b_pts + b_box + b_perm) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(spts)), "l"(sorted + (size_t)b * N), "r"(b_pts), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(sblo)), "l"(aabb + (size_t)b * nblk * 2), "r"(b_box), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(sperm)), "l"(perm16 + (size_t)b * N), "r"(b_perm), "r"(bar) : "memory");
        }
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
        }
    }

    const int r0 = blockIdx.x * KNN_ROWS_PER_CTA + wid * KNN_ROWS_PER_WARP;
    if (r0 >= N) return;
    const int b0 = r0 >> 5;

    float2 qx[4], qy[4], qz[4], qs[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const float4 a = spts[r0 + 2 * rr], c = spts[r0 + 2 * rr + 1];
        qx[rr] = make_float2(a.x, c.x);
        qy[rr] = make_float2(a.y, c.y);
        qz[rr] = make_float2(a.z, c.z);
        qs[rr] = make_float2(a.w, c.w);
    }
    RowState L[KNN_ROWS_PER_WARP];
    float* bv = sbv + (size_t)wid * KNN_ROWS_PER_WARP * KNN_CAP;
    uint32_t* bk = sbk + (size_t)wid * KNN_ROWS_PER_WARP * KNN_CAP;

    auto distances = [&](int blk, float (&d)[KNN_ROWS_PER_WARP]) {
        const float4 p = spts[blk * 32 + lane];
        const float2 px = make_float2(p.x, p.x), py = make_float2(p.y, p.y), pz = make_float2(p.z, p.z),
                     ps = make_float2(p.w, p.w);
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const float2 dd = canon_dist2<ARITH>(qx[rr], qy[rr], qz[rr], qs[rr], px, py, pz, ps);
            d[2 * rr] = dd.x;
            d[2 * rr + 1] = dd.y;
        }
    };
    auto scan_block = [&](int blk) {
        float d[KNN_ROWS_PER_WARP];
        distances(blk, d);
        bool any = false;
#pragma unroll
        for (int r = 0; r < KNN_ROWS_PER_WARP; ++r) any |= (d[r] <= L[r].thr);
        if (__any_sync(FULL, any)) {
#pragma unroll
            for (int r = 0; r < KNN_ROWS_PER_WARP; ++r) {
                const unsigned m = __ballot_sync(FULL, d[r] <= L[r].thr);
                if (m) append_hits(L[r], m, d[r], blk * 32, sperm, bv + r * KNN_CAP, bk + r * KNN_CAP, lane);
            }
        }
    };

    // ---- phase 1: the rows' own block and its index-neighbours (thr = +inf until the first compaction) --------
    int init_blk[5];
    int n_init = 0;
    {
        const int offs[5] = {0, 1, -1, 2, -2};
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            int blk = (b0 + offs[t] + nblk) % nblk;
            bool dup = false;
            for (int u = 0; u < n_init; ++u) dup |= (init_blk[u] == blk);
            if (!dup) init_blk[n_init++] = blk;
        }
    }
#pragma unroll
    for (int r = 0; r < KNN_ROWS_PER_WARP; ++r) {
        L[r].thr = INFINITY;
        L[r].n = 0;
        L[r].extra = 0;
    }
    // Three stages, each closed by a compaction of all eight rows (a single call site):
    //   0: the rows' own block and the next one (64 candidates, thr = +inf) -> first bound
    //   1: the other index-neighbours                                       -> a tight bound before pruning
    //   2: every remaining block whose AABB can still hold a candidate <= thr -> the final 20
    const int nw = (nblk + 31) >> 5;
#pragma unroll 1
    for (int stage = 0; stage < 3; ++stage) {
        if (stage < 2) {
            const int t0 = stage == 0 ? 0 : 2, t1 = stage == 0 ? (n_init < 2 ? n_init : 2) : n_init;
            for (int t = t0; t < t1; ++t) scan_block(init_blk[t]);
        } else {
            for (int w = 0; w < nw; ++w) {
                const int blk = w * 32 + lane;
                bool need = false;
                if (blk < nblk) {
                    bool done = false;
                    for (int u = 0; u < n_init; ++u) done |= (init_blk[u] == blk);
                    if (!done) {
                        if (!PRUNE) {
                            need = true;
                        } else {
                            const float4 lo = sblo[blk], hi = sbhi[blk];
#pragma unroll
                            for (int r = 0; r < KNN_ROWS_PER_WARP; ++r) {
                                const float x = (r & 1) ? qx[r >> 1].y : qx[r >> 1].x;
                                const float y = (r & 1) ? qy[r >> 1].y : qy[r >> 1].x;
                                const float z = (r & 1) ? qz[r >> 1].y : qz[r >> 1].x;
                                const float s = (r & 1) ? qs[r >> 1].y : qs[r >> 1].x;
                                const float dx = fmaxf(fmaxf(lo.x - x, x - hi.x), 0.f);
                                const float dy = fmaxf(fmaxf(lo.y - y, y - hi.y), 0.f);
                                const float dz = fmaxf(fmaxf(lo.z - z, z - hi.z), 0.f);
                                const float lb = dx * dx + dy * dy + dz * dz;
                                // rigorous lower bound of the *computed* d over the block (DESIGN.md "pruning"):
                                // true |p-q|^2 >= lb_true >= lb(1-8u); computed d >= true - 16u (s_i + s_j)
                                const float bound = lb * (1.0f - 1e-6f) - 1e-6f * (s + lo.w);
                                need |= !(bound > L[r].thr);
                            }
                        }
                    }
                }
                unsigned mask = __ballot_sync(FULL, need);
                while (mask) {
                    const int bit = __ffs(mask) - 1;
                    mask &= mask - 1;
                    scan_block(w * 32 + bit);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < KNN_ROWS_PER_WARP; ++r)
            if (L[r].n > (stage == 1 ? KNN_STAGE1_MIN : KNN_K)) L[r] = compact(L[r], bv + r * KNN_CAP, bk + r * KNN_CAP);
    }

    // ---- outputs: final compaction; the public idx output is sorted by the full key (tf.nn.top_k order), the internal
    // list is partitioned instead: neighbours outside the row's 128-point tile first (the ProxyConv gather, backbone.cu,
    // fetches those from global memory and the rest from its shared-memory window), their number in cnt's top byte.
#pragma unroll
    for (int r = 0; r < KNN_ROWS_PER_WARP; ++r) {
        float* rv = bv + r * KNN_CAP;
        uint32_t* rk = bk + r * KNN_CAP;
        const size_t row = (size_t)b * N + r0 + r;
        const bool want_pub = (idx_out || kth_out || count_out);
        uint32_t k;
        if (idx_out) {
            k = sorted_key(rv, rk);
        } else {
            __syncwarp();
            k = (lane < KNN_K) ? rk[lane] : KEY_EMPTY;
        }
        {
            const int j = (int)(k & 0xffffu);
            const bool valid = lane < KNN_K;
            const bool outside = valid && ((j >> 7) != ((r0 + r) >> 7));
            const unsigned mo = __ballot_sync(FULL, outside), mi = __ballot_sync(FULL, valid && !outside);
            const unsigned lt = (1u << lane) - 1u;
            const int n_out = __popc(mo);
            const int pos = outside ? __popc(mo & lt) : n_out + __popc(mi & lt);
            if (valid) nbr[row * KNN_K + pos] = (uint16_t)j;
            if (lane == 0) {
                kthd[row] = L[r].thr;
                cnt[row] = (KNN_K + L[r].extra) | (n_out << 24);
            }
        }
        if (want_pub) {
            const size_t orow = (size_t)b * N + sperm[r0 + r];
            if (idx_out && lane < KNN_K) idx_out[orow * KNN_K + lane] = (int32_t)(k >> 16);
            if (kth_out && lane == 0) kth_out[orow] = -L[r].thr;
            if (count_out && lane == 0) count_out[orow] = KNN_K + L[r].extra;
        }
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

int knn_build(const float* xyz, int B, int N, int arith, bool prune, float4* sorted, int* perm, uint16_t* perm16, float4* aabb, uint32_t* tie, uint16_t* nbr,
              float* kthd, int* cnt, int32_t* idx_out, float* kth_out, int32_t* count_out, cudaStream_t st) {
    if (int rc = knn_check_n(N)) return rc;
    EPC_CHECK_ARG(arith == EPC_KNN_ARITH_MULADD || arith == EPC_KNN_ARITH_FMA, "bad knn arith %d", arith);
    if (B == 0) return EPC_OK;
    const int NP = next_pow2(N);
    const size_t sort_smem = (size_t)NP * sizeof(unsigned long long);
    static bool attr_done = false;
    const size_t knn_smem = (size_t)N * 16 + (size_t)(N / 32) * 32 + (size_t)N * 2 + (size_t)KNN_WARPS * KNN_ROWS_PER_WARP * KNN_CAP * 8 + 16;
    if (!attr_done) {
        EPC_CUDA(cudaFuncSetAttribute(sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        EPC_CUDA(cudaFuncSetAttribute(knn_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        EPC_CUDA(cudaFuncSetAttribute(knn_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        EPC_CUDA(cudaFuncSetAttribute(knn_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        EPC_CUDA(cudaFuncSetAttribute(knn_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    {
        ScopedStage ss(EPC_STAGE_SORT, st);
        sort_kernel<<<B, 1024, sort_smem, st>>>(xyz, N, NP, sorted, perm, perm16, aabb);
        EPC_LAUNCH_CHECK();
    }
    EPC_CUDA(cudaMemsetAsync(tie, 0, (size_t)B * TIE_WORDS * sizeof(uint32_t), st));      // counters (and stale entries)
    ScopedStage ss(EPC_STAGE_KNN, st);
    dim3 grid((N + KNN_ROWS_PER_CTA - 1) / KNN_ROWS_PER_CTA, B);
    const int th = KNN_WARPS * 32;
    if (arith == EPC_KNN_ARITH_MULADD) {
        if (prune)
            knn_kernel<0, true><<<grid, th, knn_smem, st>>>(sorted, perm16, aabb, tie, N, nbr, kthd, cnt, idx_out, kth_out, count_out);
        else
            knn_kernel<0, false><<<grid, th, knn_smem, st>>>(sorted, perm16, aabb, tie, N, nbr, kthd, cnt, idx_out, kth_out, count_out);
    } else {
        if (prune)
            knn_kernel<1, true><<<grid, th, knn_smem, st>>>(sorted, perm16, aabb, tie, N, nbr, kthd, cnt, idx_out, kth_out, count_out);
        else
            knn_kernel<1, false><<<grid, th, knn_smem, st>>>(sorted, perm16, aabb, tie, N, nbr, kthd, cnt, idx_out, kth_out, count_out);
    }
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

size_t knn_state_bytes(int B, int N) {
    const size_t R = (size_t)B * N;
    return align_up((size_t)B * TIE_WORDS * sizeof(uint32_t)) + align_up(R * sizeof(float4)) + align_up(R * sizeof(int)) + align_up(R * sizeof(uint16_t)) + align_up(R * KNN_K * sizeof(uint16_t)) +
           align_up(R * sizeof(float)) + align_up(R * sizeof(int)) + align_up(R / 16 * sizeof(float4));
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
