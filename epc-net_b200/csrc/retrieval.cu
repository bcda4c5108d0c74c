// K6 -- retrieval: exact Euclidean top-k of queries against a descriptor database.
// Replaces `KDTree(database_output).query(q, k=25)` of evaluate.get_recall (evaluate.py:463,481), which
// evaluates float64 Euclidean distances on the fp32 descriptors one query at a time on the CPU.
//
//   1. fp32-accurate scoring  s_ij = (|q_i|^2 + |d_j|^2) - 2 q_i.d_j on tcgen05 with bf16 operand pairs (retrieval_tc.cu:
//      q = qh + ql, d = dh + dl; qh.dh + ql.dh + qh.dl), the candidate filter fused into the TMEM-drain epilogue so that the
//      Q x D score matrix is never written (FFMA GEMM + dense scan when dim is not a multiple of 64 or > 256)
//   2. per query: the 32 smallest scores among the emitted candidates       -> candidates
//   3. float64 re-rank of the candidates, sequential sum_k (q_k - d_k)^2    -> top-k, ascending (dist, index)
//   4. proof of exactness per query: every non-candidate has score >= a32 (the 32nd candidate's score),
//      hence exact d^2 >= a32 - err;  if the exact k-th distance^2 is < a32 - err the answer equals a
//      float64 brute force.  Otherwise (rare: near-ties within fp32 resolution) the query is redone by an
//      exact float64 scan of the whole database.
#include "common.cuh"
#include "kernels.h"

namespace epc {

constexpr int RC = 32;   // candidates per query

__global__ void sq_norm_kernel(const float* __restrict__ X, int R, int dim, float* __restrict__ out) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    float ss = 0.f;
    for (int i = lane; i < dim; i += 32) {
        const float v = X[(size_t)r * dim + i];
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) out[r] = ss;
}

// lexicographic (value, index) "a before b"
__device__ __forceinline__ bool key_less(double av, long long ai, double bv, long long bi) {
    return (av < bv) || (av == bv && ai < bi);
}

// One warp per query: 32 smallest of score_j = dn[j] - 2 dot[j] (= |q - d_j|^2 - |q|^2), ascending; columns ascend so ties
// keep the lower index.
// RAW: `dots` already holds the scores (the tensor-core epilogue applied the norms)
template <bool RAW>
__global__ void candidates_kernel(const float* __restrict__ dots, int ld, const float* __restrict__ qn,
                                  const float* __restrict__ dn, int Qt, int D, int* __restrict__ cand,
                                  float* __restrict__ a32) {
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (qi >= Qt) return;
    const float* row = dots + (size_t)qi * ld;
    float val = INFINITY;
    int vi = -1;
    int filled = 0;
    float thr = INFINITY;
    constexpr int U = 8;                              // 32-column blocks loaded together: the scan is latency-bound otherwise
    for (int jb = 0; jb < D; jb += 32 * U) {
        float sv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = jb + 32 * u + lane;
            sv[u] = (j < D) ? (RAW ? __ldg(row + j) : fmaf(-2.0f, __ldg(row + j), __ldg(dn + j))) : INFINITY;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j0 = jb + 32 * u;
            const float s = sv[u];
            unsigned m = __ballot_sync(FULL, (j0 + lane < D) && (filled < RC || s < thr));
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float c = __shfl_sync(FULL, s, src);
                const bool before = (lane < filled) && (val <= c);
                const int pos = __popc(__ballot_sync(FULL, before));
                if (pos < RC) {
                    const float upv = __shfl_up_sync(FULL, val, 1);
                    const int upi = __shfl_up_sync(FULL, vi, 1);
                    if (lane == pos) {
                        val = c;
                        vi = j0 + src;
                    } else if (lane > pos) {
                        val = upv;
                        vi = upi;
                    }
                    filled = min(filled + 1, RC);
                    thr = (filled == RC) ? __shfl_sync(FULL, val, RC - 1) : INFINITY;
                }
            }
        }
    }
    cand[(size_t)qi * RC + lane] = (lane < filled) ? vi : -1;
    if (lane == 0) a32[qi] = thr;      // +inf when D < 32: every row is a candidate
}

__device__ __forceinline__ double exact_d2(const float* __restrict__ q, const float* __restrict__ d, int dim) {
    double acc = 0.0;
    int k = 0;
    if ((dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
        // same sequential order, operands fetched 16 bytes at a time with a batch of loads in flight
        const float4* q4 = reinterpret_cast<const float4*>(q);
        const float4* d4 = reinterpret_cast<const float4*>(d);
        for (; k + 16 <= dim; k += 16) {
            float4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a[u] = __ldg(q4 + (k >> 2) + u);
                b[u] = __ldg(d4 + (k >> 2) + u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double t0 = (double)a[u].x - (double)b[u].x, t1 = (double)a[u].y - (double)b[u].y;
                const double t2 = (double)a[u].z - (double)b[u].z, t3 = (double)a[u].w - (double)b[u].w;
                acc = __dadd_rn(acc, __dmul_rn(t0, t0));   // no FMA contraction: sklearn's rdist loop is mul then add
                acc = __dadd_rn(acc, __dmul_rn(t1, t1));
                acc = __dadd_rn(acc, __dmul_rn(t2, t2));
                acc = __dadd_rn(acc, __dmul_rn(t3, t3));
            }
        }
    }
    for (; k < dim; ++k) {
        const double t = (double)q[k] - (double)d[k];
        acc = __dadd_rn(acc, __dmul_rn(t, t));
    }
    return acc;
}

// the same sum with the query already converted to float64 (shared memory, one copy per warp): half the F2F conversions
__device__ __forceinline__ double exact_d2_staged(const double* __restrict__ qd, const float* __restrict__ d, int dim) {
    double acc = 0.0;
    int k = 0;
    if ((dim & 3) == 0 && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        const float4* d4 = reinterpret_cast<const float4*>(d);
        for (; k + 16 <= dim; k += 16) {
            float4 b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) b[u] = __ldg(d4 + (k >> 2) + u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double t0 = qd[k + 4 * u] - (double)b[u].x, t1 = qd[k + 4 * u + 1] - (double)b[u].y;
                const double t2 = qd[k + 4 * u + 2] - (double)b[u].z, t3 = qd[k + 4 * u + 3] - (double)b[u].w;
                acc = __dadd_rn(acc, __dmul_rn(t0, t0));
                acc = __dadd_rn(acc, __dmul_rn(t1, t1));
                acc = __dadd_rn(acc, __dmul_rn(t2, t2));
                acc = __dadd_rn(acc, __dmul_rn(t3, t3));
            }
        }
    }
    for (; k < dim; ++k) {
        const double t = qd[k] - (double)d[k];
        acc = __dadd_rn(acc, __dmul_rn(t, t));
    }
    return acc;
}

// bitonic sort of one (value,index) pair per lane, ascending
__device__ __forceinline__ void warp_sort_pairs(double& v, long long& i, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const double ov = __shfl_xor_sync(FULL, v, j);
            const long long oi = __shfl_xor_sync(FULL, i, j);
            const bool up = ((lane & k) == 0), lower = ((lane & j) == 0);
            const bool other_less = key_less(ov, oi, v, i);
            const bool take = (lower == up) ? other_less : (!other_less && !(ov == v && oi == i));
            if (take) {
                v = ov;
                i = oi;
            }
        }
    }
}

// MODE 0: each lane walks its candidate's row in global memory; MODE 1: the same with the query pre-converted to float64 in
// shared memory; MODE 2 (dim % 4 == 0, 16-byte aligned rows): the 32 candidate rows are brought in with COALESCED loads --
// 64 dimensions at a time, 16 lanes per row, two rows per load instruction -- into a padded shared-memory tile, from which each
// lane then sums its own row in the sequential order (a lane-per-row LDG.128 touches 32 half-used sectors per instruction and
// is L1-throughput bound: 36 us instead of ~12 for 3000 queries).  Dynamic shared memory: warps x dim doubles (MODE >= 1)
// + warps x 32 x RR_PITCH floats (MODE 2).
constexpr int RR_CHUNK = 64, RR_PITCH = RR_CHUNK + 4;
template <int MODE>
__global__ void rerank_kernel(const float* __restrict__ db, const float* __restrict__ q, const int* __restrict__ cand,
                              const float* __restrict__ a32, const float* __restrict__ qn, const float* __restrict__ dn_max_p, int Qt, int dim,
                              int k, long long id_offset, int64_t* __restrict__ idx, double* __restrict__ dist,
                              int* __restrict__ flags, float err_unit) {
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int qi = blockIdx.x * nw + w;
    const int lane = threadIdx.x & 31;
    if (qi >= Qt) return;
    const int c = cand[(size_t)qi * RC + lane];
    double v = INFINITY;
    long long gi = 0x7fffffffffffffffLL;
    extern __shared__ double s_q[];
    double* qd = s_q + (size_t)w * dim;
    if (MODE >= 1) {
        for (int i = lane; i < dim; i += 32) qd[i] = (double)q[(size_t)qi * dim + i];
        __syncwarp();
    }
    if (MODE == 2) {
        float* tile = reinterpret_cast<float*>(s_q + (size_t)nw * dim) + (size_t)w * 32 * RR_PITCH;
        const int sub = lane >> 4, col = (lane & 15) * 4;          // this lane loads floats [col, col + 4) of rows 2 j + sub
        double acc = 0.0;
        int cr[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) cr[j] = __shfl_sync(FULL, c, 2 * j + sub);
        float4 x[16];                                              // the next chunk, in flight while this one is summed
        auto fetch = [&](int c0) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                x[j] = (cr[j] >= 0 && c0 + col < dim) ? __ldg(reinterpret_cast<const float4*>(db + (size_t)cr[j] * dim + c0 + col))
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        fetch(0);
        for (int c0 = 0; c0 < dim; c0 += RR_CHUNK) {
#pragma unroll
            for (int j = 0; j < 16; ++j) *reinterpret_cast<float4*>(tile + (2 * j + sub) * RR_PITCH + col) = x[j];
            __syncwarp();
            if (c0 + RR_CHUNK < dim) fetch(c0 + RR_CHUNK);
            const float* mine = tile + lane * RR_PITCH;
            const int n = min(RR_CHUNK, dim - c0);
            for (int e = 0; e < n; e += 4) {
                const float4 b = *reinterpret_cast<const float4*>(mine + e);
                const double t0 = qd[c0 + e] - (double)b.x, t1 = qd[c0 + e + 1] - (double)b.y;
                const double t2 = qd[c0 + e + 2] - (double)b.z, t3 = qd[c0 + e + 3] - (double)b.w;
                acc = __dadd_rn(acc, __dmul_rn(t0, t0));   // no FMA contraction: sklearn's rdist loop is mul then add
                acc = __dadd_rn(acc, __dmul_rn(t1, t1));
                acc = __dadd_rn(acc, __dmul_rn(t2, t2));
                acc = __dadd_rn(acc, __dmul_rn(t3, t3));
            }
            __syncwarp();
        }
        if (c >= 0) {
            v = acc;
            gi = (long long)c + id_offset;
        }
    } else if (c >= 0) {
        v = (MODE == 1) ? exact_d2_staged(qd, db + (size_t)c * dim, dim) : exact_d2(q + (size_t)qi * dim, db + (size_t)c * dim, dim);
        gi = (long long)c + id_offset;
    }
    warp_sort_pairs(v, gi, lane);
    if (lane < k) {
        idx[(size_t)qi * k + lane] = (v == INFINITY && gi == 0x7fffffffffffffffLL) ? -1 : gi;
        dist[(size_t)qi * k + lane] = sqrt(v);
    }
    const double kth = __shfl_sync(FULL, v, k - 1);
    if (lane == 0) {
        // scoring error bound: |score - exact d^2| <= err = err_unit (|q|^2 + max|d|^2).  FFMA path: dim-term FMA chain + norm
        // sums, 1.2e-7 (dim + 8).  3xTF32 path: products exact in fp32, dropped ql.dl and split residues 3 x 2^-22, accumulation
        // of 3 dim terms at <= 2^-22 each (the tensor core may truncate): 2.4e-7 (3 dim + 8) -- both generous.
        const float err = err_unit * (qn[qi] + *dn_max_p);
        const float a = a32[qi] + qn[qi];              // scores omit the query's own |q|^2
        flags[qi] = (a == INFINITY) ? 0 : !((float)kth * (1.0f + 2e-7f) < a - err);
    }
}

// Exact float64 scan for flagged queries (one warp per query).
__global__ void exact_fallback_kernel(const float* __restrict__ db, const float* __restrict__ q,
                                      const int* __restrict__ flags, int Qt, int D, int dim, int k, long long id_offset,
                                      int64_t* __restrict__ idx, double* __restrict__ dist) {
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (qi >= Qt || !flags[qi]) return;
    double val = INFINITY;
    long long vi = 0x7fffffffffffffffLL;
    int filled = 0;
    double thr = INFINITY;
    for (int j0 = 0; j0 < D; j0 += 32) {
        const int j = j0 + lane;
        const double s = (j < D) ? exact_d2(q + (size_t)qi * dim, db + (size_t)j * dim, dim) : (double)INFINITY;
        unsigned m = __ballot_sync(FULL, (j < D) && (filled < k || s < thr));
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const double c = __shfl_sync(FULL, s, src);
            const bool before = (lane < filled) && (val <= c);
            const int pos = __popc(__ballot_sync(FULL, before));
            if (pos < k) {
                const double upv = __shfl_up_sync(FULL, val, 1);
                const long long upi = __shfl_up_sync(FULL, vi, 1);
                if (lane == pos) {
                    val = c;
                    vi = (long long)(j0 + src) + id_offset;
                } else if (lane > pos && lane < k) {
                    val = upv;
                    vi = upi;
                }
                filled = min(filled + 1, k);
                thr = (filled == k) ? __shfl_sync(FULL, val, k - 1) : (double)INFINITY;
            }
        }
    }
    if (lane < k) {
        idx[(size_t)qi * k + lane] = (lane < filled) ? vi : -1;
        dist[(size_t)qi * k + lane] = sqrt(val);
    }
}

// k > 32 (KDTree.query accepts any k; train.py:857-869 passes num_to_take): exact float64 scans in passes of 32 -- pass p
// collects the 32 smallest (d^2, row) keys that are lexicographically greater than the last key of pass p - 1.  One warp per
// query and D x dim float64 operations per pass: the rare path, exact by construction.
__global__ void exact_scan_pass_kernel(const float* __restrict__ db, const float* __restrict__ q, int Qt, int D, int dim, int kk,
                                       int col0, int ld, long long id_offset, int has_lb, double* __restrict__ lb_d2,
                                       long long* __restrict__ lb_row, int64_t* __restrict__ idx, double* __restrict__ dist) {
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (qi >= Qt) return;
    const double bd = has_lb ? lb_d2[qi] : -1.0;
    const long long br = has_lb ? lb_row[qi] : -1;
    double val = INFINITY;
    long long vi = 0x7fffffffffffffffLL;
    int filled = 0;
    double thr = INFINITY;
    for (int j0 = 0; j0 < D; j0 += 32) {
        const int j = j0 + lane;
        const double s = (j < D) ? exact_d2(q + (size_t)qi * dim, db + (size_t)j * dim, dim) : (double)INFINITY;
        const bool after_lb = (s > bd) || (s == bd && (long long)j > br);
        unsigned m = __ballot_sync(FULL, (j < D) && after_lb && (filled < kk || s < thr));
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const double c = __shfl_sync(FULL, s, src);
            const bool before = (lane < filled) && (val <= c);
            const int pos = __popc(__ballot_sync(FULL, before));
            if (pos < kk) {
                const double upv = __shfl_up_sync(FULL, val, 1);
                const long long upi = __shfl_up_sync(FULL, vi, 1);
                if (lane == pos) {
                    val = c;
                    vi = (long long)(j0 + src);
                } else if (lane > pos && lane < kk) {
                    val = upv;
                    vi = upi;
                }
                filled = min(filled + 1, kk);
                thr = (filled == kk) ? __shfl_sync(FULL, val, kk - 1) : (double)INFINITY;
            }
        }
    }
    if (lane < kk) {
        idx[(size_t)qi * ld + col0 + lane] = (lane < filled) ? vi + id_offset : -1;
        dist[(size_t)qi * ld + col0 + lane] = sqrt(val);
    }
    if (lane == kk - 1) {               // the next pass continues after this key (inf / max row when the database is exhausted)
        lb_d2[qi] = val;
        lb_row[qi] = vi;
    }
}

__global__ void max_reduce_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
    __shared__ float s[32];
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, x[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s[w]);
        *out = m;
    }
}

// One warp per query: an upper bound of the 32nd smallest of the query's S sample scores that at most ~40 of them reach
// (bisection on the value between the smallest lane minimum and the largest lane minimum -- each lane's minimum is a
// distinct sample element, so 32 of them lie at or below the largest).  S <= 32 * PER_LANE.
template <int PER_LANE>
__global__ void sample_threshold_kernel(const float* __restrict__ scores, int ld, int S, int Qt, float* __restrict__ thr) {
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (qi >= Qt) return;
    const float* row = scores + (size_t)qi * ld;
    float v[PER_LANE];
    float mn = INFINITY;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
        v[i] = (32 * i + lane < S) ? __ldg(row + 32 * i + lane) : INFINITY;
        mn = fminf(mn, v[i]);
    }
    float lo = warp_min(mn), hi = warp_max(mn);
    for (int it = 0; it < 16; ++it) {
        const float t = 0.5f * lo + 0.5f * hi;
        if (!(t > lo && t < hi)) break;
        int c = 0;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) c += (v[i] <= t) ? 1 : 0;
        c = __reduce_add_sync(FULL, c);
        if (c >= RC) {
            hi = t;
            if (c <= RC + 8) break;
        } else {
            lo = t;
        }
    }
    if (lane == 0) thr[qi] = hi;
}

// One warp per query: the 32 smallest (score, row) among the candidates the scoring epilogue emitted into the query's
// regions (retrieval_tc.cu).  The emission order is arbitrary, the (score, row) order makes the result deterministic.
__global__ void select_kernel(const uint2* __restrict__ cand_in, const int* __restrict__ counts, int n_regions, int cap, int Qt,
                              int* __restrict__ cand, float* __restrict__ a32) {
    extern __shared__ int s_counts[];                     // [warps per block][n_regions]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + w;
    if (qi >= Qt) return;
    int* cnt = s_counts + w * n_regions;
    bool overflow = false;
    for (int r = lane; r < n_regions; r += 32) {
        const int c = __ldg(counts + (size_t)qi * n_regions + r);
        overflow |= c < 0;
        cnt[r] = c < 0 ? 0 : c;
    }
    overflow = __any_sync(FULL, overflow);
    __syncwarp();
    const uint2* base = cand_in + (size_t)qi * n_regions * cap;
    const uint2 none = make_uint2(0x7f800000u, 0x7fffffffu);
    // chunks of 32 entries, walked region by region; the next chunk's load is in flight while this one is merged
    int r = 0, e0 = 0;
    while (r < n_regions && cnt[r] == 0) ++r;
    uint2 nxt = (r < n_regions && e0 + lane < cnt[r]) ? __ldg(base + (size_t)r * cap + e0 + lane) : none;
    float val = INFINITY;
    int vi = 0x7fffffff;
    int filled = 0;
    float thr = INFINITY;
    int thr_i = 0x7fffffff;
    while (r < n_regions) {
        const uint2 ent = nxt;
        e0 += 32;
        if (e0 >= cnt[r]) {
            e0 = 0;
            ++r;
            while (r < n_regions && cnt[r] == 0) ++r;
        }
        nxt = (r < n_regions && e0 + lane < cnt[r]) ? __ldg(base + (size_t)r * cap + e0 + lane) : none;
        const float s = __uint_as_float(ent.x);
        const int si = (int)ent.y;
        unsigned m = __ballot_sync(FULL, si != 0x7fffffff && (filled < RC || s < thr || (s == thr && si < thr_i)));
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cv = __shfl_sync(FULL, s, src);
            const int ci = __shfl_sync(FULL, si, src);
            const bool before = (lane < filled) && (val < cv || (val == cv && vi < ci));
            const int pos = __popc(__ballot_sync(FULL, before));
            if (pos < RC) {
                const float upv = __shfl_up_sync(FULL, val, 1);
                const int upi = __shfl_up_sync(FULL, vi, 1);
                if (lane == pos) {
                    val = cv;
                    vi = ci;
                } else if (lane > pos) {
                    val = upv;
                    vi = upi;
                }
                filled = min(filled + 1, RC);
                if (filled == RC) {
                    thr = __shfl_sync(FULL, val, RC - 1);
                    thr_i = __shfl_sync(FULL, vi, RC - 1);
                }
            }
        }
    }
    cand[(size_t)qi * RC + lane] = (lane < filled) ? vi : -1;
    // an overflowed region may hide a better candidate, and fewer than 32 emitted entries (cannot happen while the sample's
    // own rows score the same in both passes) leave no bound on the others: a32 = -inf sends the query to the exact fallback
    if (lane == 0) a32[qi] = (overflow || filled < RC) ? -INFINITY : thr;
}

__global__ void fill_empty_kernel(int64_t* __restrict__ idx, double* __restrict__ dist, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        idx[i] = -1;
        dist[i] = INFINITY;
    }
}

static int pad256(int D) { return (D + 255) / 256 * 256; }

constexpr int RETR_DENSE_MAX = 4096;      // databases up to this many rows: dense scores + scan (the matrix is small)
constexpr int RETR_QTILE = 8192;          // queries per pass (bounds the workspace)
constexpr int RETR_SAMPLE_TILES = 16;     // threshold sample: 16 tiles of 128 rows spread over the database

// entries of one candidate region (one query x one range x one column half).  The threshold lets 32..40 of the 2048 sample
// scores through, i.e. ~2 % of exchangeable rows; 3x headroom + 32.
static int region_cap(int tiles_per_range) {
    const long long rows = 64ll * tiles_per_range;
    return (int)((rows * 6 / 100 + 32 + 7) / 8 * 8);
}

// ---- the prepared database ("index"): what KDTree(database_output) (evaluate.py:463) is to the reference --------------------
struct RetrIndex {             // laid out at the start of the caller's index memory
    int D, dim, tensor;
    float dn_max;              // (device copy lives in dn_max_dev)
};
size_t retrieve_index_bytes(int D, int dim) {
    const int Dp = pad256(D);
    return 256 + align_up((size_t)Dp * 4) + align_up(16) + (retr_tc_supported(dim) ? align_up((size_t)(D > 0 ? D : 1) * 2 * dim * 2) : 0) + 256;
}
struct IndexView {
    float* dn;                 // [pad256(D)] |d|^2, +inf padding
    float* dn_max;             // [1]
    __nv_bfloat16* db2;        // [D, 2 dim] (tensor path)
};
static IndexView index_view(void* mem, int D, int dim) {
    Arena ar(mem, (size_t)1 << 60);
    IndexView v;
    v.dn = ar.take<float>(pad256(D));
    v.dn_max = ar.take<float>(4);
    v.db2 = retr_tc_supported(dim) ? ar.take<__nv_bfloat16>((size_t)(D > 0 ? D : 1) * 2 * dim) : nullptr;
    return v;
}
int retrieve_index_build(const float* db, int D, int dim, void* index_mem, size_t index_bytes, cudaStream_t st) {
    EPC_CHECK_ARG(D >= 0 && dim >= 1, "retrieve_index_build: bad sizes D=%d dim=%d", D, dim);
    EPC_CHECK_ARG(index_mem && index_bytes >= retrieve_index_bytes(D, dim), "retrieve_index_build: index memory %zu < required %zu",
                  index_bytes, retrieve_index_bytes(D, dim));
    if (D == 0) return EPC_OK;
    IndexView v = index_view(index_mem, D, dim);
    const int Dp = pad256(D);
    if (v.db2) {
        if (int rc = retr_split2(db, D, Dp, dim, v.db2, v.dn, st)) return rc;
    } else {
        sq_norm_kernel<<<(D + 7) / 8, 256, 0, st>>>(db, D, dim, v.dn);
        EPC_LAUNCH_CHECK();
    }
    max_reduce_kernel<<<1, 1024, 0, st>>>(v.dn, D, v.dn_max);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

size_t retrieve_workspace_bytes(int D, int Q, int dim, int k) {
    if (k > 32) return align_up((size_t)(Q > 0 ? Q : 1) * 8) * 2 + 512;
    const int qt = Q < RETR_QTILE ? (Q > 0 ? Q : 1) : RETR_QTILE;
    size_t s = retrieve_index_bytes(D, dim) + align_up((size_t)qt * 4) * 3 + align_up((size_t)qt * RC * 4) + 1024;
    const int Dp = pad256(D);
    if (retr_tc_supported(dim)) {
        s += align_up((size_t)qt * 2 * dim * 2);
        if (D <= RETR_DENSE_MAX) {
            s += align_up((size_t)qt * Dp * 4);
        } else {
            const int n_tiles = Dp / 128;
            const int nr = retr_ranges(qt, n_tiles);
            const int tpr = (n_tiles + nr - 1) / nr;
            s += align_up((size_t)qt * RETR_SAMPLE_TILES * 128 * 4) + align_up((size_t)qt * 2 * nr * region_cap(tpr) * 8) + align_up((size_t)qt * 2 * nr * 4);
        }
    } else {
        s += align_up((size_t)qt * Dp * 4);
    }
    return s;
}

int retrieve_topk(const float* db, int D, const float* q, int Q, int dim, int k, long long id_offset, const void* index,
                  int64_t* idx, double* dist, void* ws, size_t ws_bytes, cudaStream_t st) {
    EPC_CHECK_ARG(k >= 1, "retrieve_topk: k=%d", k);
    EPC_CHECK_ARG(D >= 0 && dim >= 1 && Q >= 0, "retrieve_topk: bad sizes D=%d Q=%d dim=%d", D, Q, dim);
    if (Q == 0) return EPC_OK;
    if (k > 32 && D > 0) {      // more neighbours than the candidate lists hold: exact float64 scans, 32 per pass
        if (ws_bytes < retrieve_workspace_bytes(D, Q, dim, k)) {
            set_error("retrieve_topk: workspace %zu < required %zu", ws_bytes, retrieve_workspace_bytes(D, Q, dim, k));
            return EPC_EWORKSPACE;
        }
        Arena ar(ws, ws_bytes);
        double* lb_d2 = ar.take<double>(Q);
        long long* lb_row = ar.take<long long>(Q);
        for (int col0 = 0; col0 < k; col0 += 32) {
            const int kk = (k - col0 < 32) ? (k - col0) : 32;
            exact_scan_pass_kernel<<<(Q + 7) / 8, 256, 0, st>>>(db, q, Q, D, dim, kk, col0, k, id_offset, col0 > 0, lb_d2, lb_row, idx, dist);
            EPC_LAUNCH_CHECK();
        }
        return EPC_OK;
    }
    if (D == 0) {               // an empty shard (more ranks than rows): all padding
        const long long n = (long long)Q * k;
        fill_empty_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(idx, dist, n);
        EPC_LAUNCH_CHECK();
        return EPC_OK;
    }
    if (ws_bytes < retrieve_workspace_bytes(D, Q, dim, k)) {
        set_error("retrieve_topk: workspace %zu < required %zu", ws_bytes, retrieve_workspace_bytes(D, Q, dim, k));
        return EPC_EWORKSPACE;
    }
    const bool tensor = retr_tc_supported(dim);
    const int Dp = pad256(D);
    const int qt = Q < RETR_QTILE ? Q : RETR_QTILE;
    Arena ar(ws, ws_bytes);
    void* own_index = ar.take<unsigned char>(retrieve_index_bytes(D, dim));
    float* qn = ar.take<float>(qt);
    float* a32 = ar.take<float>(qt);
    int* flags = ar.take<int>(qt);
    int* cand = ar.take<int>((size_t)qt * RC);
    if (!index) {
        ScopedStage ss(EPC_STAGE_RETRIEVE_SCORE, st);
        if (int rc = retrieve_index_build(db, D, dim, own_index, retrieve_index_bytes(D, dim), st)) return rc;
        index = own_index;
    }
    const IndexView iv = index_view(const_cast<void*>(index), D, dim);
    // scoring error bound: |score + |q|^2 - exact d^2| <= err_unit (|q|^2 + max |d|^2).  bf16-pair path: dropped ql.dl and split
    // residues 2^-16, accumulation of 3 dim products at <= 2^-22 each (the tensor core may truncate); FFMA path: dim-term FMA
    // chain + norm sums, 1.2e-7 (dim + 8) -- both generous.
    const float err_unit = tensor ? 2.4e-7f * (float)(3 * dim + 8) + 4e-5f : 1.2e-7f * (float)(dim + 8);

    for (int q0 = 0; q0 < Q; q0 += qt) {
        const int nq = (Q - q0 < qt) ? (Q - q0) : qt;
        const float* qp = q + (size_t)q0 * dim;
        if (tensor) {
            __nv_bfloat16* q2 = ar.take<__nv_bfloat16>((size_t)qt * 2 * dim);
            const int n_tiles = Dp / 128;
            {
                ScopedStage ss(EPC_STAGE_RETRIEVE_SCORE, st);
                if (int rc = retr_split2(qp, nq, nq, dim, q2, qn, st)) return rc;
            }
            if (D <= RETR_DENSE_MAX) {
                float* scores = ar.take<float>((size_t)qt * Dp);
                {
                    ScopedStage ss(EPC_STAGE_RETRIEVE_SCORE, st);
                    if (int rc = retr_scores(q2, nq, iv.db2, D, dim, n_tiles, 1, retr_ranges(nq, n_tiles), iv.dn, scores, Dp, nullptr, nullptr,
                                             nullptr, 0, st))
                        return rc;
                }
                ScopedStage ss(EPC_STAGE_RETRIEVE_SELECT, st);
                candidates_kernel<true><<<(nq + 7) / 8, 256, 0, st>>>(scores, Dp, qn, iv.dn, nq, D, cand, a32);
                EPC_LAUNCH_CHECK();
            } else {
                const int nr = retr_ranges(nq, n_tiles);
                const int tpr = (n_tiles + nr - 1) / nr;
                const int cap = region_cap(tpr);
                constexpr int S = RETR_SAMPLE_TILES * 128;
                float* sample = ar.take<float>((size_t)qt * S);
                uint2* regions = ar.take<uint2>((size_t)qt * 2 * nr * cap);
                int* counts = ar.take<int>((size_t)qt * 2 * nr);
                {
                    ScopedStage ss(EPC_STAGE_RETRIEVE_SCORE, st);
                    // scores of 16 tiles spread over the database -> an upper bound of every query's 32nd smallest score ...
                    if (int rc = retr_scores(q2, nq, iv.db2, D, dim, RETR_SAMPLE_TILES, n_tiles / RETR_SAMPLE_TILES,
                                             retr_ranges(nq, RETR_SAMPLE_TILES), iv.dn, sample, S, nullptr, nullptr, nullptr, 0, st))
                        return rc;
                    sample_threshold_kernel<S / 32><<<(nq + 7) / 8, 256, 0, st>>>(sample, S, S, nq, a32);
                    EPC_LAUNCH_CHECK();
                    // ... then every row scoring at or below it, straight from the accumulators
                    if (int rc = retr_scores(q2, nq, iv.db2, D, dim, n_tiles, 1, nr, iv.dn, nullptr, 0, a32, regions, counts, cap, st))
                        return rc;
                }
                ScopedStage ss(EPC_STAGE_RETRIEVE_SELECT, st);
                select_kernel<<<(nq + 7) / 8, 256, (size_t)8 * 2 * nr * sizeof(int), st>>>(regions, counts, 2 * nr, cap, nq, cand, a32);
                EPC_LAUNCH_CHECK();
            }
        } else {
            const int ld = pad256(D);
            float* dots = ar.take<float>((size_t)qt * ld);
            {
                ScopedStage ss(EPC_STAGE_RETRIEVE_SCORE, st);
                sq_norm_kernel<<<(nq + 7) / 8, 256, 0, st>>>(qp, nq, dim, qn);
                EPC_LAUNCH_CHECK();
                GemmArgs g = {};
                g.A = qp;  g.sAm = dim; g.sAk = 1;
                g.B = db;  g.sBk = 1;   g.sBn = dim;
                g.C = dots; g.ldc = ld; g.M = nq; g.N = D; g.K = dim; g.batch = 1; g.splitk = 1;
                if (int rc = sgemm(g, st)) return rc;
            }
            ScopedStage ss(EPC_STAGE_RETRIEVE_SELECT, st);
            candidates_kernel<false><<<(nq + 7) / 8, 256, 0, st>>>(dots, ld, qn, iv.dn, nq, D, cand, a32);
            EPC_LAUNCH_CHECK();
        }
        ScopedStage ss(EPC_STAGE_RETRIEVE_RERANK, st);
        if (dim <= 512 && dim % 4 == 0 && (reinterpret_cast<uintptr_t>(db) & 15) == 0) {
            const size_t smem = (size_t)8 * dim * sizeof(double) + (size_t)8 * 32 * RR_PITCH * sizeof(float);
            static PerDeviceSize attr;
            EPC_CUDA(ensure_dyn_smem(rerank_kernel<2>, smem, attr));
            rerank_kernel<2><<<(nq + 7) / 8, 256, smem, st>>>(db, qp, cand, a32, qn, iv.dn_max, nq, dim, k, id_offset,
                                                             idx + (size_t)q0 * k, dist + (size_t)q0 * k, flags, err_unit);
        } else if (dim <= 512) {
            rerank_kernel<1><<<(nq + 7) / 8, 256, (size_t)8 * dim * sizeof(double), st>>>(
                db, qp, cand, a32, qn, iv.dn_max, nq, dim, k, id_offset, idx + (size_t)q0 * k, dist + (size_t)q0 * k, flags, err_unit);
        } else {
            rerank_kernel<0><<<(nq + 7) / 8, 256, 0, st>>>(db, qp, cand, a32, qn, iv.dn_max, nq, dim, k, id_offset,
                                                          idx + (size_t)q0 * k, dist + (size_t)q0 * k, flags, err_unit);
        }
        EPC_LAUNCH_CHECK();
        exact_fallback_kernel<<<(nq + 7) / 8, 256, 0, st>>>(db, qp, flags, nq, D, dim, k, id_offset, idx + (size_t)q0 * k,
                                                            dist + (size_t)q0 * k);
        EPC_LAUNCH_CHECK();
        ar.off = (size_t)((char*)cand - ar.base) + align_up((size_t)qt * RC * 4);      // the per-pass buffers are reused by the next pass
    }
    if (!ar.ok()) {
        set_error("retrieve_topk: internal workspace accounting error");
        return EPC_EWORKSPACE;
    }
    return EPC_OK;
}

// Radius search: the reference builds its evaluation pickles with sklearn's
// KDTree(db[['northing','easting']]).query_radius(coor, r=25) (generating_queries/generate_test_sets.py:70-104), one query per
// call.  Here: one warp per query, float64 (UTM coordinates need it), membership d^2 = sum_k (q_k - d_k)^2 <= r^2 evaluated in
// sklearn's order; two passes (count, then fill at the caller's exclusive offsets), indices ascending.
template <bool FILL>
__global__ void radius_kernel(const double* __restrict__ db, int D, const double* __restrict__ q, int Q, int dim, double r2,
                              int32_t* __restrict__ counts, const int64_t* __restrict__ offsets, int32_t* __restrict__ indices) {
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (qi >= Q) return;
    const double* qq = q + (size_t)qi * dim;
    int n = 0;
    for (int j0 = 0; j0 < D; j0 += 32) {
        const int j = j0 + lane;
        bool in = false;
        if (j < D) {
            double acc = 0.0;
            for (int k = 0; k < dim; ++k) {
                const double t = qq[k] - db[(size_t)j * dim + k];
                acc = __dadd_rn(acc, __dmul_rn(t, t));
            }
            in = (acc <= r2);
        }
        const unsigned m = __ballot_sync(FULL, in);
        if (FILL && in) indices[offsets[qi] + n + __popc(m & ((1u << lane) - 1u))] = j;
        n += __popc(m);
    }
    if (!FILL && lane == 0) counts[qi] = n;
}

int radius_search(const double* db, int D, const double* q, int Q, int dim, double r, int32_t* counts, const int64_t* offsets,
                  int32_t* indices, cudaStream_t st) {
    EPC_CHECK_ARG(D >= 0 && Q >= 0 && dim >= 1 && r >= 0.0, "radius_search: bad sizes D=%d Q=%d dim=%d r=%g", D, Q, dim, r);
    if (Q == 0) return EPC_OK;
    const double r2 = r * r;
    if (indices)
        radius_kernel<true><<<(Q + 7) / 8, 256, 0, st>>>(db, D, q, Q, dim, r2, nullptr, offsets, indices);
    else
        radius_kernel<false><<<(Q + 7) / 8, 256, 0, st>>>(db, D, q, Q, dim, r2, counts, nullptr, nullptr);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// Merge R shard lists per query by (distance, index): one warp per query, lists staged in shared memory.
__global__ void merge_topk_kernel(const double* __restrict__ dist, const int64_t* __restrict__ idx, long long rank_stride, int R,
                                  int Q, int k, double* __restrict__ out_dist, int64_t* __restrict__ out_idx) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int qi = blockIdx.x, lane = threadIdx.x;
    const int n = R * k;
    double* sd = reinterpret_cast<double*>(sm);
    long long* si = reinterpret_cast<long long*>(sd + n);
    for (int t = lane; t < n; t += 32) {
        const int r = t / k, j = t % k;
        const size_t src = (size_t)r * rank_stride + (size_t)qi * k + j;
        long long id = idx[src];
        sd[t] = (id < 0) ? (double)INFINITY : dist[src];
        si[t] = (id < 0) ? 0x7fffffffffffffffLL : id;
    }
    __syncwarp();
    for (int o = 0; o < k; ++o) {
        double bv = INFINITY;
        long long bi = 0x7fffffffffffffffLL;
        int bt = -1;
        for (int t = lane; t < n; t += 32) {
            if (key_less(sd[t], si[t], bv, bi)) {
                bv = sd[t];
                bi = si[t];
                bt = t;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(FULL, bv, off);
            const long long oi = __shfl_xor_sync(FULL, bi, off);
            const int ot = __shfl_xor_sync(FULL, bt, off);
            if (key_less(ov, oi, bv, bi)) {
                bv = ov;
                bi = oi;
                bt = ot;
            }
        }
        if (lane == 0) {
            out_dist[(size_t)qi * k + o] = bv;
            out_idx[(size_t)qi * k + o] = (bt < 0) ? -1 : bi;
            if (bt >= 0) {
                sd[bt] = INFINITY;
                si[bt] = 0x7fffffffffffffffLL;
            }
        }
        __syncwarp();
    }
}

int merge_topk(const double* dist, const int64_t* idx, long long rank_stride, int R, int Q, int k, double* out_dist,
               int64_t* out_idx, cudaStream_t st) {
    EPC_CHECK_ARG(R >= 1 && k >= 1 && (size_t)R * k * 16 <= 48 * 1024, "merge_topk: R=%d k=%d unsupported", R, k);
    EPC_CHECK_ARG(rank_stride >= (long long)Q * k, "merge_topk: rank_stride %lld < Q k", rank_stride);
    if (Q == 0) return EPC_OK;
    merge_topk_kernel<<<Q, 32, (size_t)R * k * 16, st>>>(dist, idx, rank_stride, R, Q, k, out_dist, out_idx);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

}  // namespace epc
