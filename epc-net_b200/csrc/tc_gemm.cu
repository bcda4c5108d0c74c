// Instantiations of the tcgen05 GEMM template (tc_gemm.cuh) for the dense contractions of the EPC-Net head.
#include <stdlib.h>
#include "tc_gemm.cuh"
#include "kernels.h"

namespace epc {

// conv5 (models/epc-net.py:136-139): H = relu(Xc W5 + b5) as bf16 [R,1024] + rowss [R, 1024/256] partial |H_n|^2
int tc_conv5_bf16(const __nv_bfloat16* Xc, long long R, int cin, const __nv_bfloat16* W5t, const float* b5,
                  __nv_bfloat16* H, float* rowss, cudaStream_t st) {
    tc::GemmParams p = {};
    p.M = (int)R; p.N = 1024; p.K = cin; p.splitk = 1; p.C = H; p.ldc = 1024; p.bias = b5; p.relu = 1; p.aux = rowss;
    { static const int pf = getenv("EPC_CONV5_PREFETCH") ? atoi(getenv("EPC_CONV5_PREFETCH")) : 0; p.l2_prefetch_tiles = pf; }
    Operand<__nv_bfloat16> a{Xc, R, cin, cin}, b{W5t, 1024, cin, cin};
    // Variant with the A tiles TMA-multicast to a cluster of the 4 N-tile CTAs (L2->SM operand traffic / 4): correct, but
    // measured slower on B200 (2.72 vs 2.47 us/cloud) -- the kernel is bound by the 8 MiB/cloud H write, not by operand
    // reads, and the cluster couples four CTAs' pipelines.  Kept selectable for experiments.
    static const bool mc = getenv("EPC_CONV5_MULTICAST") != nullptr;
    if (mc) return tc_gemm_bres_launch<__nv_bfloat16, 256, tc::EPI_CONV5_BF16, 8, 4>(a, b, p, st);
    return tc_gemm_bres_launch<__nv_bfloat16, 256, tc::EPI_CONV5_BF16, 8>(a, b, p, st);
}
int conv5_rowss_parts() { return 8; }     // 1024 / BN(256) N tiles x 2 epilogue warps per lane quarter

// cluster assignment (loupe.py:255-276): S' = softmax(BN((H Wc)/|H|))/|H| as bf16 [R,64]; a_part [R/128, 64]
int tc_assign(const __nv_bfloat16* H, long long R, const __nv_bfloat16* Wct, const float* rowss, int parts,
              const float* bn_scale, const float* bn_shift, __nv_bfloat16* S, float* a_part, cudaStream_t st) {
    tc::GemmParams p = {};
    p.M = (int)R; p.N = 64; p.K = 1024; p.splitk = 1; p.C = S; p.ldc = 64; p.aux = a_part; p.rowss = rowss;
    p.rowss_parts = parts; p.bn_scale = bn_scale; p.bn_shift = bn_shift;
    p.reverse_m = getenv("EPC_ASSIGN_FORWARD") ? 0 : 1;     // conv5 wrote the last rows of H last: start with what is still in L2
    Operand<__nv_bfloat16> a{H, R, 1024, 1024}, b{Wct, 64, 1024, 1024};
    return tc_gemm_bres_launch<__nv_bfloat16, 64, tc::EPI_ASSIGN>(a, b, p, st);
}

// VLAD accumulate (loupe.py:286-291): V[b] = H[b]^T S'[b]  -> fp32 slabs [splitk][B,1024,64]
int tc_vlad(const __nv_bfloat16* H, const __nv_bfloat16* S, int B, int N, float* V, int splitk, long long slab,
            cudaStream_t st) {
    tc::GemmParams p = {};
    p.M = 1024; p.N = 64; p.K = N / splitk; p.k_batch_rows = N; p.splitk = splitk; p.C = V; p.ldc = 64;
    p.c_batch = 1024ll * 64; p.c_slab = slab;
    Operand<__nv_bfloat16> a{H, (long long)B * N, 1024, 1024}, b{S, (long long)B * N, 64, 64};
    return tc_gemm_launch<__nv_bfloat16, 64, true, true, tc::EPI_STORE_F32>(a, b, p, B, st, 1);
}

// EPC-Net-L (models/epc-net-l.py:84-91): g[b,:] = max_n relu(Xc W5 + b5) -- H is never written.  TF32.
int tc_conv5_colmax(const float* Xc, long long R, int cin, int rows_per_cloud, const float* W5t, const float* b5,
                    float* g, int clouds, const int* cloud_mask, cudaStream_t st) {
    if (!cloud_mask) EPC_CUDA(cudaMemsetAsync(g, 0, sizeof(float) * (size_t)clouds * 1024, st));
    tc::GemmParams p = {};
    p.M = (int)R; p.N = 1024; p.K = cin; p.splitk = 1; p.bias = b5; p.aux = g; p.rows_per_cloud = rows_per_cloud; p.cloud_mask = cloud_mask;
    Operand<float> a{Xc, R, cin, cin}, b{W5t, 1024, cin, cin};
    return tc_gemm_bres_launch<float, 256, tc::EPI_COLMAX, 8>(a, b, p, st);
}

// the same with bf16 operands (fp32 accumulation): half the operand bytes through shared memory, twice the tensor rate
int tc_conv5_colmax_bf16(const __nv_bfloat16* Xc, long long R, int cin, int rows_per_cloud, const __nv_bfloat16* W5t, const float* b5,
                         float* g, int clouds, cudaStream_t st) {
    EPC_CUDA(cudaMemsetAsync(g, 0, sizeof(float) * (size_t)clouds * 1024, st));
    tc::GemmParams p = {};
    p.M = (int)R; p.N = 1024; p.K = cin; p.splitk = 1; p.bias = b5; p.aux = g; p.rows_per_cloud = rows_per_cloud;
    Operand<__nv_bfloat16> a{Xc, R, cin, cin}, b{W5t, 1024, cin, cin};
    return tc_gemm_bres_launch<__nv_bfloat16, 256, tc::EPI_COLMAX, 8>(a, b, p, st);
}

// fp16 operands: the same 10-bit mantissa as TF32 at twice the tensor rate and half the operand bytes.  Clouds whose values left
// the fp16 range (flags) hold garbage here; the caller re-does them with the masked TF32 kernel on the fp32 safe-pass rows.
int tc_conv5_colmax_f16(const __half* Xc, long long R, int cin, int rows_per_cloud, const __half* W5t, const float* b5, float* g, int clouds,
                        cudaStream_t st) {
    EPC_CUDA(cudaMemsetAsync(g, 0, sizeof(float) * (size_t)clouds * 1024, st));
    tc::GemmParams p = {};
    p.M = (int)R; p.N = 1024; p.K = cin; p.splitk = 1; p.bias = b5; p.aux = g; p.rows_per_cloud = rows_per_cloud;
    Operand<__half> a{Xc, R, cin, cin}, b{W5t, 1024, cin, cin};
    return tc_gemm_bres_launch<__half, 256, tc::EPI_COLMAX, 8>(a, b, p, st);
}

__global__ void reset_rows_flagged_kernel(float* __restrict__ g, const int* __restrict__ flags, int cols) {
    if (!flags[blockIdx.x]) return;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) g[(size_t)blockIdx.x * cols + i] = 0.f;
}
int reset_rows_flagged(float* g, const int* flags, int clouds, int cols, cudaStream_t st) {
    if (clouds == 0) return EPC_OK;
    reset_rows_flagged_kernel<<<clouds, 256, 0, st>>>(g, flags, cols);
    EPC_LAUNCH_CHECK();
    return EPC_OK;
}

// fp32-output conv5 on TF32 tensor cores (KD feature export, models/kd_epc-net.py:158)
int tc_conv5_f32(const float* Xc, long long R, int cin, const float* W5t, const float* b5, float* H, cudaStream_t st) {
    tc::GemmParams p = {};
    p.M = (int)R; p.N = 1024; p.K = cin; p.splitk = 1; p.C = H; p.ldc = 1024; p.bias = b5; p.relu = 1;
    Operand<float> a{Xc, R, cin, cin}, b{W5t, 1024, cin, cin};
    if (cin > 128)      // fp32 operands: a 256-column slice of W5 (256 KB at K = 256) would not fit in shared memory
        return tc_gemm_bres_launch<float, 128, tc::EPI_STORE_F32, 8>(a, b, p, st);
    return tc_gemm_bres_launch<float, 256, tc::EPI_STORE_F32, 8>(a, b, p, st);
}

// hidden FC of the VLAD head (loupe.py:302-320): Y[s] = v[:, ks] Wh[ks, :]   (rows = B*G, K = hidden_in), TF32, split-K slabs
int tc_hidden(const float* v, int rows, int hidden_in, const float* Wht /*[D, hidden_in]*/, int D, float* Y, int splitk,
              cudaStream_t st) {
    tc::GemmParams p = {};
    p.M = rows; p.N = D; p.K = hidden_in / splitk; p.splitk = splitk; p.C = Y; p.ldc = D; p.c_slab = (long long)rows * D;
    Operand<float> a{v, rows, hidden_in, hidden_in}, b{Wht, D, hidden_in, hidden_in};
    // N tile = the widest of 256 / 128 / 64 that divides FEATURE_OUTPUT_DIM (256 in every shipped config)
    if (D % 256 == 0) return tc_gemm_launch<float, 256, false, false, tc::EPI_STORE_F32>(a, b, p, 1, st, 1);
    if (D % 128 == 0) return tc_gemm_launch<float, 128, false, false, tc::EPI_STORE_F32>(a, b, p, 1, st, 1);
    EPC_CHECK_ARG(D % 64 == 0, "tc_hidden: output_dim=%d must be a multiple of 64", D);
    return tc_gemm_launch<float, 64, false, false, tc::EPI_STORE_F32>(a, b, p, 1, st, 1);
}

// EPC-Net-L head FC (models/epc-net-l.py:95): o[B, D] = relu(g[B,1024] . Wfc + b) on TF32 tensor cores (BN folded into W, b)
int tc_fc_relu(const float* g, int rows, int K, const float* Wt /*[D, K]*/, const float* bias, int D, float* out, cudaStream_t st) {
    tc::GemmParams p = {};
    p.M = rows; p.N = D; p.K = K; p.splitk = 1; p.C = out; p.ldc = D; p.bias = bias; p.relu = 1;
    Operand<float> a{g, rows, K, K}, b{Wt, D, K, K};
    if (D % 256 == 0) return tc_gemm_launch<float, 256, false, false, tc::EPI_STORE_F32>(a, b, p, 1, st, 1);
    if (D % 128 == 0) return tc_gemm_launch<float, 128, false, false, tc::EPI_STORE_F32>(a, b, p, 1, st, 1);
    EPC_CHECK_ARG(D % 64 == 0, "tc_fc_relu: output_dim=%d must be a multiple of 64", D);
    return tc_gemm_launch<float, 64, false, false, tc::EPI_STORE_F32>(a, b, p, 1, st, 1);
}

}  // namespace epc
