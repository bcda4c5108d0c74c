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

__global__ void __launch_bounds__(1024) sort_kernel(const float* __restrict__ xyz, int N, int NP,
                                                     float4* __restrict__ sorted, int* __restrict__ perm) {
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
    for (int i = tid; i < N; i += blockDim.x) {
        int src = (int)(keys[i] & 0xffffffffu);
        float x = p[3 * src], y = p[3 * src + 1], z = p[3 * src + 2];
        sorted[(size_t)b * N + i] = make_float4(x, y, z, canon_sq(x, y, z));
        perm[(size_t)b * N + i] = src;
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-distributed sorted list: lane l (< 20) holds the l-th smallest (d, original index) seen so far.
// ------------------------------------------------------------------------------------------------
struct RowList {
    float val;    // this lane's entry value (+inf on lanes >= 20)
    int vi;       // sorted-space position of the entry
    float thr;    // value of entry 19 (warp-uniform)
    int extra;    // # seen candidates with d == thr that are not in the list (warp-uniform)
};

// Insert every candidate flagged in `m` (lane s holds candidate value d at sorted position jbase+s).
__device__ __forceinline__ void insert_hits(RowList& L, unsigned m, float d, int jbase, const int* __restrict__ gperm,
                                            int lane) {
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float c = __shfl_sync(FULL, d, src);
        if (c > L.thr) continue;   // the mask was taken against an older (larger) threshold: no longer a member
        const int cj = jbase + src;
        bool before = (L.val < c);
        if (__any_sync(FULL, L.val == c)) {      // rare: equal distances are ordered by ORIGINAL index (tf.nn.top_k)
            const int co = __ldg(gperm + cj);
            const int vo = __ldg(gperm + L.vi);
            before = before || (L.val == c && vo < co);
        }
        const int pos = __popc(__ballot_sync(FULL, before) & 0xFFFFFu);
        if (pos < KNN_K) {
            const float ev = L.thr;
            const float upv = __shfl_up_sync(FULL, L.val, 1);
            const int upi = __shfl_up_sync(FULL, L.vi, 1);
            if (lane == pos) {
                L.val = c;
                L.vi = cj;
            } else if (lane > pos && lane < KNN_K) {
                L.val = upv;
                L.vi = upi;
            }
            const float nthr = __shfl_sync(FULL, L.val, KNN_K - 1);
            L.extra = (nthr == ev) ? L.extra + 1 : 0;
            L.thr = nthr;
        } else {
            L.extra += 1;   // c == thr and it loses the index tie-break: member of the thresholded set only
        }
    }
}

// Build the list from one full block of 32 candidates (bitonic sort across lanes on (d, original idx)).
__device__ __forceinline__ void init_list(RowList& L, float d, int jbase, const int* __restrict__ gperm, int lane) {
    float val = d;
    int vi = jbase + lane;
    int vo = __ldg(gperm + vi);
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const float ov = __shfl_xor_sync(FULL, val, j);
            const int oi = __shfl_xor_sync(FULL, vi, j);
            const int oo = __shfl_xor_sync(FULL, vo, j);
            const bool up = ((lane & k) == 0);
            const bool lower = ((lane & j) == 0);
            const bool other_less = (ov < val) || (ov == val && oo < vo);
            const bool take = (lower == up) ? other_less : !other_less;
            if (take) {
                val = ov;
                vi = oi;
                vo = oo;
            }
        }
    }
    L.thr = __shfl_sync(FULL, val, KNN_K - 1);
    L.extra = __popc(__ballot_sync(FULL, lane >= KNN_K && val == L.thr));
    L.val = (lane < KNN_K) ? val : INFINITY;
    L.vi = (lane < KNN_K) ? vi : jbase;
}

constexpr int KNN_ROWS_PER_WARP = 8;
constexpr int KNN_WARPS = 8;
constexpr int KNN_ROWS_PER_CTA = KNN_ROWS_PER_WARP * KNN_WARPS;

template <int ARITH, bool PRUNE>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_kernel(const float4* __restrict__ sorted, const int* __restrict__ perm, int N, uint16_t* __restrict__ nbr,
           float* __restrict__ kthd, int* __restrict__ cnt, int32_t* __restrict__ idx_out,
           float* __restrict__ kth_out, int32_t* __restrict__ count_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nblk = N >> 5;
    float4* spts = reinterpret_cast<float4*>(smem_raw);   // [N]   (x,y,z,s)
    float4* sblo = spts + N;                              // [nblk] (lo.xyz, max s)
    float4* sbhi = sblo + nblk;                           // [nblk] (hi.xyz, -)
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float4* gp = sorted + (size_t)b * N;
    const int* gperm = perm + (size_t)b * N;

    for (int i = tid; i < N; i += blockDim.x) spts[i] = gp[i];
    __syncthreads();
    for (int blk = wid; blk < nblk; blk += KNN_WARPS) {
        const float4 p = spts[blk * 32 + lane];
        const float lx = warp_min(p.x), ly = warp_min(p.y), lz = warp_min(p.z);
        const float hx = warp_max(p.x), hy = warp_max(p.y), hz = warp_max(p.z);
        const float sm = warp_max(p.w);
        if (lane == 0) {
            sblo[blk] = make_float4(lx, ly, lz, sm);
            sbhi[blk] = make_float4(hx, hy, hz, 0.f);
        }
    }
    __syncthreads();

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
    RowList L[KNN_ROWS_PER_WARP];

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
                if (m) insert_hits(L[r], m, d[r], blk * 32, gperm, lane);
            }
        }
    };

    // ---- phase 1: the query rows' own block builds the lists; its index-neighbours tighten them ----
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
    {
        float d[KNN_ROWS_PER_WARP];
        distances(b0, d);
#pragma unroll
        for (int r = 0; r < KNN_ROWS_PER_WARP; ++r) init_list(L[r], d[r], b0 * 32, gperm, lane);
    }
    for (int t = 1; t < n_init; ++t) scan_block(init_blk[t]);

    // ---- phase 2: every remaining block whose AABB can still hold a candidate <= thr --------------
    const int nw = (nblk + 31) >> 5;
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

    // ---- outputs ------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < KNN_ROWS_PER_WARP; ++r) {
        const size_t row = (size_t)b * N + r0 + r;
        if (lane < KNN_K) nbr[row * KNN_K + lane] = (uint16_t)L[r].vi;
        if (lane == 0) {
            kthd[row] = L[r].thr;
            cnt[row] = KNN_K + L[r].extra;
        }
        if (idx_out || kth_out || count_out) {
            const size_t orow = (size_t)b * N + __ldg(gperm + r0 + r);
            if (idx_out && lane < KNN_K) idx_out[orow * KNN_K + lane] = __ldg(gperm + L[r].vi);
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

int knn_build(const float* xyz, int B, int N, int arith, bool prune, float4* sorted, int* perm, uint16_t* nbr,
              float* kthd, int* cnt, int32_t* idx_out, float* kth_out, int32_t* count_out, cudaStream_t st) {
    if (int rc = knn_check_n(N)) return rc;
    EPC_CHECK_ARG(arith == EPC_KNN_ARITH_MULADD || arith == EPC_KNN_ARITH_FMA, "bad knn arith %d", arith);
    if (B == 0) return EPC_OK;
    const int NP = next_pow2(N);
    const size_t sort_smem = (size_t)NP * sizeof(unsigned long long);
    static bool attr_done = false;
    const size_t knn_smem = (size_t)N * 16 + (size_t)(N / 32) * 32;
    if (!attr_done) {
        EPC_CUDA(cudaFuncSetAttribute(sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        EPC_CUDA(cudaFuncSetAttribute(knn_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
        EPC_CUDA(cudaFuncSetAttribute(knn_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
        EPC_CUDA(cudaFuncSetAttribute(knn_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
        EPC_CUDA(cudaFuncSetAttribute(knn_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
        attr_done = true;
    }
    {
        ScopedStage ss(EPC_STAGE_SORT, st);
        sort_kernel<<<B, 1024, sort_smem, st>>>(xyz, N, NP, sorted, perm);
        EPC_LAUNCH_CHECK();
    }
    ScopedStage ss(EPC_STAGE_KNN, st);
    dim3 grid((N + KNN_ROWS_PER_CTA - 1) / KNN_ROWS_PER_CTA, B);
    const int th = KNN_WARPS * 32;
    if (arith == EPC_KNN_ARITH_MULADD) {
        if (prune)
            knn_kernel<0, true><<<grid, th, knn_smem, st>>>(sorted, perm, N, nbr, kthd, cnt, idx_out, kth_out, count_out);
        else
            knn_kernel<0, false><<<grid, th, knn_smem, st>>>(sorted, perm, N, nbr, kthd, cnt, idx_out, kth_out, count_out);
    } else {
        if (prune)
            knn_kernel<1, true><<<grid, th, knn_smem, st>>>(sorted, perm, N, nbr, kthd, cnt, idx_out, kth_out, count_out);
        else
            knn_kernel<1, false><<<grid, th, knn_smem, st>>>(sorted, perm, N, nbr, kthd, cnt, idx_out, kth_out, count_out);
    }
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

size_t knn_state_bytes(int B, int N) {
    const size_t R = (size_t)B * N;
    return align_up(R * sizeof(float4)) + align_up(R * sizeof(int)) + align_up(R * KNN_K * sizeof(uint16_t)) +
           align_up(R * sizeof(float)) + align_up(R * sizeof(int));
}

KnnState knn_state_carve(Arena& ar, int B, int N) {
    const size_t R = (size_t)B * N;
    KnnState s;
    s.sorted = ar.take<float4>(R);
    s.perm = ar.take<int>(R);
    s.nbr = ar.take<uint16_t>(R * KNN_K);
    s.kthd = ar.take<float>(R);
    s.cnt = ar.take<int>(R);
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
