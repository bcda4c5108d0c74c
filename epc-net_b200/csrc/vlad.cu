// K4 -- G_VLAD / NetVLAD aggregation tail and the EPC-Net-L head (loupe.py:233-333, 61-101;
// models/epc-net.py:147-155; models/epc-net-l.py:88-100).  Small, bandwidth-bound pieces that sit
// around the three dense contractions (cluster assignment, VLAD accumulate, hidden FC).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace epc {

// inv[r] = 1 / sqrt(max(sum_f X[r,f]^2, 1e-12))     (tf.nn.l2_normalize, models/epc-net.py:147)
__global__ void row_inv_norm_kernel(const float* __restrict__ X, long long R, int F, float* __restrict__ inv) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float4* x4 = reinterpret_cast<const float4*>(X + r * F);
    float ss = 0.f;
    for (int i = lane; i < F / 4; i += 32) {
        const float4 v = __ldg(x4 + i);
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    if (lane == 0) inv[r] = 1.0f / sqrtf(fmaxf(ss, L2_EPS));
}

int row_inv_norm(const float* X, long long R, int F, float* inv, cudaStream_t st) {
    EPC_CHECK_ARG(F % 4 == 0, "row_inv_norm: F=%d must be a multiple of 4", F);
    if (R == 0) return EPC_OK;
    row_inv_norm_kernel<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(X, R, F, inv);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// VLAD finalise (loupe.py:284-298): r[f,c] = V[f,c] - a_sum[c] * Wc2[f,c]; L2 over f per (b,c); flatten f-major
// (index f*K + c); global L2.  Two passes over (cloud, 128-feature slice) CTAs:
//   pass 1: r -> v (unnormalised), partial column sums of squares -> colss [B, F/128, K]
//   pass 2: every CTA re-derives the 64 column scales and the global scale from colss (fixed order) and scales its slice.
constexpr int VF_ROWS = 128;

template <int NSLAB>
__global__ void __launch_bounds__(256)
vlad_residual_kernel(const float* __restrict__ V, long long slab, const float* __restrict__ vscale, const float* __restrict__ a_sum,
                     int a_parts, const float* __restrict__ Wc2, int F, int K, float* __restrict__ v, float* __restrict__ colss) {
    __shared__ float s_part[4][64];
    const int b = blockIdx.y, sl = blockIdx.x, tid = threadIdx.x;
    const float vs = vscale ? __ldg(vscale + b) : 1.0f;      // fp8 head: V carries the cloud's power-of-two S'' scale (head_fp8.cu)
    const int c = tid & 63, grp = tid >> 6;
    float as = 0.f;                                  // fixed summation order; the loads are issued eight at a time
    if (c < K) {
        for (int p0 = 0; p0 < a_parts; p0 += 8) {
            float t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = (p0 + u < a_parts) ? __ldg(a_sum + ((size_t)b * a_parts + p0 + u) * K + c) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) as += t[u];
        }
    }
    const float* Vb = V + (size_t)b * F * K;
    float* vb = v + (size_t)b * F * K;
    float ss = 0.f;
    if (c < K) {
        const int f0 = sl * VF_ROWS + grp, f1 = min((sl + 1) * VF_ROWS, F);
        constexpr int U = 8;                       // independent loads in flight per thread (the kernel is pure streaming)
        for (int fb = f0; fb < f1; fb += 4 * U) {
            float part[U][NSLAB], w2[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int f = fb + 4 * u;
                w2[u] = 0.f;
#pragma unroll
                for (int s = 0; s < NSLAB; ++s) part[u][s] = 0.f;
                if (f < f1) {
#pragma unroll
                    for (int s = 0; s < NSLAB; ++s) part[u][s] = __ldg(Vb + s * slab + (size_t)f * K + c);
                    w2[u] = __ldg(Wc2 + (size_t)f * K + c);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int f = fb + 4 * u;
                if (f < f1) {
                    float acc = 0.f;                 // split-K slabs summed in a fixed order
#pragma unroll
                    for (int s = 0; s < NSLAB; ++s) acc += part[u][s];
                    const float r = acc * vs - as * w2[u];
                    vb[(size_t)f * K + c] = r;
                    ss += r * r;
                }
            }
        }
    }
    s_part[grp][c] = ss;
    __syncthreads();
    if (tid < K) colss[((size_t)b * gridDim.x + sl) * K + tid] = (s_part[0][tid] + s_part[1][tid]) + (s_part[2][tid] + s_part[3][tid]);
}

__global__ void __launch_bounds__(256)
vlad_scale_kernel(float* __restrict__ v, const float* __restrict__ colss, int F, int K) {
    __shared__ float s_inv[64];
    __shared__ float s_g[64];
    __shared__ float s_ginv;
    const int b = blockIdx.y, sl = blockIdx.x, tid = threadIdx.x;
    if (tid < 64) {
        float tot = 0.f;
        if (tid < K) {
            float t[8];                              // gridDim.x == F / 128 == 8 partials: all loads in flight, fixed order
            for (int p0 = 0; p0 < (int)gridDim.x; p0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = (p0 + u < (int)gridDim.x) ? __ldg(colss + ((size_t)b * gridDim.x + p0 + u) * K + tid) : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) tot += t[u];
            }
        }
        const float iv = 1.0f / sqrtf(fmaxf(tot, L2_EPS));
        s_inv[tid] = iv;
        s_g[tid] = (tid < K) ? tot * iv * iv : 0.f;          // squared norm of the normalised column
    }
    __syncthreads();
    if (tid < 32) {
        float g = warp_sum(s_g[tid] + s_g[tid + 32]);
        if (tid == 0) s_ginv = 1.0f / sqrtf(fmaxf(g, L2_EPS));
    }
    __syncthreads();
    const int c = tid & 63, grp = tid >> 6;
    if (c < K) {
        const float sc = s_inv[c] * s_ginv;
        float* vb = v + (size_t)b * F * K;
        for (int f = sl * VF_ROWS + grp; f < (sl + 1) * VF_ROWS && f < F; f += 4) vb[(size_t)f * K + c] *= sc;
    }
}

// Split-K slabs of the VLAD accumulate.  More slabs = smaller work items for the VLAD CTAs of the fused assignment + VLAD
// launch, which therefore stay within ~1 cloud of the assignment front (head_fused.cu) at the price of 256 KB of fp32
// partial sums per slab and cloud.
int vlad_splitk() {
    static const int v = [] {
        const char* e = getenv("EPC_VLAD_SPLITK");
        const int x = e ? atoi(e) : 0;
        return (x == 1 || x == 2 || x == 4 || x == 8) ? x : 2;
    }();
    return v;
}

int vlad_finalize(const float* V, int nslab, long long slab, const float* vscale, const float* a_sum, int a_parts, const float* Wc2,
                  int B, int F, int K, float* v, float* colss, cudaStream_t st) {
    EPC_CHECK_ARG(K >= 1 && K <= 64, "vlad_finalize: cluster_size=%d unsupported (1..64)", K);
    if (B == 0) return EPC_OK;
    dim3 grid((F + VF_ROWS - 1) / VF_ROWS, B);
    if (nslab == 1)
        vlad_residual_kernel<1><<<grid, 256, 0, st>>>(V, slab, vscale, a_sum, a_parts, Wc2, F, K, v, colss);
    else if (nslab == 2)
        vlad_residual_kernel<2><<<grid, 256, 0, st>>>(V, slab, vscale, a_sum, a_parts, Wc2, F, K, v, colss);
    else if (nslab == 4)
        vlad_residual_kernel<4><<<grid, 256, 0, st>>>(V, slab, vscale, a_sum, a_parts, Wc2, F, K, v, colss);
    else if (nslab == 8)
        vlad_residual_kernel<8><<<grid, 256, 0, st>>>(V, slab, vscale, a_sum, a_parts, Wc2, F, K, v, colss);
    else {
        set_error("vlad_finalize: %d split-K slabs unsupported (1, 2, 4 or 8)", nslab);
        return EPC_EUNSUPPORTED;
    }
    EPC_LAUNCH_CHECK();
    vlad_scale_kernel<<<grid, 256, 0, st>>>(v, colss, F, K);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// Tail (loupe.py:320-331, 61-101; models/epc-net.py:153): Y [B*G, D] raw hidden products ->
//   y = sum_g BN_bn(Y[b,g,:]);  z = y * sigmoid(BN_gating(y Wg));  out = l2 ? z/|z| : z.   One CTA per cloud, D threads.
__global__ void vlad_tail_kernel(const float* __restrict__ Y, int nslab, size_t slab, int G, int D, const float* __restrict__ bn_scale,
                                 const float* __restrict__ bn_shift, const float* __restrict__ Wg,
                                 const float* __restrict__ g_scale, const float* __restrict__ g_shift, int gating,
                                 int l2, float* __restrict__ out) {
    extern __shared__ float sy[];          // [D] + [32]
    float* sred = sy + D;
    const int b = blockIdx.x, d = threadIdx.x;
    float y = 0.f;
    if (d < D) {
        for (int g = 0; g < G; ++g) {
            float h = 0.f;                    // split-K partial slabs, summed in a fixed order; loads issued eight at a time
            for (int s0 = 0; s0 < nslab; s0 += 8) {
                float t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = (s0 + u < nslab) ? __ldg(Y + (s0 + u) * slab + ((size_t)b * G + g) * D + d) : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) h += t[u];
            }
            y += h * bn_scale[d] + bn_shift[d];
        }
        sy[d] = y;
    }
    __syncthreads();
    float z = y;
    if (gating && d < D) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;       // four independent chains: the loop is L2-latency bound
        int k = 0;
        for (; k + 4 <= D; k += 4) {
            a0 = fmaf(sy[k], __ldg(Wg + (size_t)k * D + d), a0);
            a1 = fmaf(sy[k + 1], __ldg(Wg + (size_t)(k + 1) * D + d), a1);
            a2 = fmaf(sy[k + 2], __ldg(Wg + (size_t)(k + 2) * D + d), a2);
            a3 = fmaf(sy[k + 3], __ldg(Wg + (size_t)(k + 3) * D + d), a3);
        }
        for (; k < D; ++k) a0 = fmaf(sy[k], __ldg(Wg + (size_t)k * D + d), a0);
        const float acc = (a0 + a1) + (a2 + a3);
        const float gt = acc * g_scale[d] + g_shift[d];
        z = y * (1.0f / (1.0f + expf(-gt)));
    }
    if (l2) {
        float ss = (d < D) ? z * z : 0.f;
        ss = warp_sum(ss);
        if ((d & 31) == 0) sred[d >> 5] = ss;
        __syncthreads();
        float tot = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sred[w];
        z *= 1.0f / sqrtf(fmaxf(tot, L2_EPS));
    }
    if (d < D) out[(size_t)b * D + d] = z;
}

int vlad_tail(const float* Y, int nslab, int B, int G, int D, const float* bn_scale, const float* bn_shift, const float* Wg,
              const float* g_scale, const float* g_shift, int gating, int l2, float* out, cudaStream_t st) {
    EPC_CHECK_ARG(D >= 1 && D <= 1024, "vlad_tail: output_dim=%d unsupported (1..1024)", D);
    if (B == 0) return EPC_OK;
    const int threads = (D + 31) / 32 * 32;
    vlad_tail_kernel<<<B, threads, (D + 32) * sizeof(float), st>>>(Y, nslab, (size_t)B * G * D, G, D, bn_scale, bn_shift, Wg, g_scale, g_shift,
                                                                   gating, l2, out);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// g[b,f] = max_n H[b,n,f]   (tf_util.max_pool2d over [N,1], models/epc-net-l.py:91).  H >= 0 is NOT assumed:
// a float atomic max via the ordered-int trick.
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (__float_as_int(v) >= 0)          // sign BIT, not value: -0.0f has the int pattern INT_MIN and must take the negative branch
        atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void fill_kernel(float* p, long long n, float v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void col_max_kernel(const float* __restrict__ H, int N, int F, int rows_per_cta, float* __restrict__ g) {
    const int b = blockIdx.z;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int n0 = blockIdx.y * rows_per_cta, n1 = min(N, n0 + rows_per_cta);
    const float* h = H + ((size_t)b * N) * F + f;
    float m = -INFINITY;
    for (int n = n0; n < n1; ++n) m = fmaxf(m, __ldg(h + (size_t)n * F));
    if (n1 > n0) atomic_max_float(&g[(size_t)b * F + f], m);
}

int col_max(const float* H, int B, int N, int F, float* g, cudaStream_t st) {
    if (B == 0) return EPC_OK;
    const long long tot = (long long)B * F;
    fill_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, tot, -INFINITY);
    EPC_LAUNCH_CHECK();
    const int rows = 128;
    dim3 grid((F + 127) / 128, (N + rows - 1) / rows, B);
    col_max_kernel<<<grid, 128, 0, st>>>(H, N, F, rows, g);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// out[r,:] = X[r,:] / sqrt(max(|X[r,:]|^2, 1e-12)); one warp per row
__global__ void row_l2_normalize_kernel(const float* __restrict__ X, int R, int D, float* __restrict__ out) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    float ss = 0.f;
    for (int i = lane; i < D; i += 32) {
        const float v = X[(size_t)r * D + i];
        ss += v * v;
    }
    ss = warp_sum(ss);
    const float iv = 1.0f / sqrtf(fmaxf(ss, L2_EPS));
    for (int i = lane; i < D; i += 32) out[(size_t)r * D + i] = X[(size_t)r * D + i] * iv;
}

int row_l2_normalize(const float* X, int R, int D, float* out, cudaStream_t st) {
    if (R == 0) return EPC_OK;
    row_l2_normalize_kernel<<<(R + 7) / 8, 256, 0, st>>>(X, R, D, out);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// KD feature (models/kd_epc-net.py:158): feat[b, perm[pos], :] = H[b,pos,:] * inv[b,pos]  (back to original order)
__global__ void kd_feat_kernel(const float* __restrict__ H, const float* __restrict__ inv, const int* __restrict__ perm,
                               int N, int F, float* __restrict__ feat) {
    const size_t row = blockIdx.x;     // b*N + pos
    const size_t b = row / N;
    const size_t orow = b * N + perm[row];
    const float iv = inv[row];
    const float4* src = reinterpret_cast<const float4*>(H + row * F);
    float4* dst = reinterpret_cast<float4*>(feat + orow * F);
    for (int i = threadIdx.x; i < F / 4; i += blockDim.x) {
        float4 v = __ldg(src + i);
        v.x *= iv; v.y *= iv; v.z *= iv; v.w *= iv;
        dst[i] = v;
    }
}

int kd_feat(const float* H, const float* inv, const int* perm, int B, int N, int F, float* feat, cudaStream_t st) {
    if (B == 0) return EPC_OK;
    kd_feat_kernel<<<(unsigned)((size_t)B * N), 256, 0, st>>>(H, inv, perm, N, F, feat);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// X fp32 [R,F] -> bf16, and rowss[r] := 1 (the stand-alone loupe API uses the caller's rows as given)
__global__ void f32_to_bf16_rows_kernel(const float* __restrict__ X, long long n4, __nv_bfloat16* __restrict__ Y,
                                        float* __restrict__ rowss, long long R) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(X) + i);
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        reinterpret_cast<uint2*>(Y)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
    if (i < R) rowss[i] = 1.0f;
}

int f32_to_bf16_rows(const float* X, long long R, int F, __nv_bfloat16* Y, float* rowss, cudaStream_t st) {
    EPC_CHECK_ARG(F % 4 == 0, "f32_to_bf16_rows: F=%d must be a multiple of 4", F);
    if (R == 0) return EPC_OK;
    const long long n4 = R * F / 4;
    f32_to_bf16_rows_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(X, n4, Y, rowss, R);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

}  // namespace epc
