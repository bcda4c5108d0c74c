// Internal (C++) interface between the translation units of libepc_b200.so.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace epc {

// ---- knn.cu ------------------------------------------------------------------------------------
struct KnnState {          // per-cloud kNN graph in *sorted* (Morton) point order
    float4* sorted;        // [B,N]   (x,y,z,|p|^2)
    int* perm;             // [B,N]   sorted position -> original index
    uint16_t* perm16;      // [B,N]   the same as u16 (N <= 8192), bulk-copied into the kNN kernel's shared memory
    uint16_t* nbr;         // [B,N,20] neighbour positions (sorted space); those outside the row's 128-point tile come first
    float* kthd;           // [B,N]   20th smallest d
    int* cnt;              // [B,N]   |{j : d_ij <= kthd_i}| (>= 20) in the low 24 bits | (# neighbours outside the row's 128-point tile) << 24
    float4* aabb;          // [B][2][N/32] per 32-point block: (lo.xyz, max |p|^2), then (hi.xyz, -)
    uint32_t* tie;         // [B][TIE_WORDS] per cloud: count, pad, then up to TIE_CAP entries (row << 16 | neighbour, d bits):
                           // the members of thresholded sets beyond the 20 listed ones (ties at the 20th distance)
    float* U;              // [B,N]   pass A: upper bound of the row's 20th smallest canonical distance
    int* slow;             // [1 + B*N] counter, then the rows whose candidate list overflowed (mass ties): exact warp-per-row path
    uint16_t* glist;       // [B,N,40] pass B: the row's listed candidates (canonical d may be <= U)
    int* gcount;           // [B,N]    ... their number, -1 = the list overflowed
};
constexpr uint32_t TIE_CAP = 510;                  // entries per cloud; more (degenerate clouds) -> the gather re-scans the cloud
constexpr uint32_t TIE_WORDS = 2 + 2 * TIE_CAP;    // 4 KB per cloud
int knn_check_n(int N);
size_t knn_state_bytes(int B, int N);
KnnState knn_state_carve(Arena& ar, int B, int N);
int knn_build(const float* xyz, int B, int N, int arith, bool prune, const KnnState& s, int32_t* idx_out, float* kth_out,
              int32_t* count_out, cudaStream_t st);
int knn_dense(const float* xyz, int B, int N, int arith, const float* kth, float* mask, float* dist, cudaStream_t st);
int rows_topk_smallest(const float* adj, long long R, int M, int k, int32_t* idx, cudaStream_t st);

// ---- backbone.cu -------------------------------------------------------------------------------
struct DenseDev {          // BN-folded pointwise layer on the device: y = act(x W + b)
    const float* W;        // [cin, cout]
    const float* b;        // [cout]
    int cin, cout;
    const void* Wimg;      // 64->64 layers: 8 KB shared-memory image (fp16, swizzled) of the tensor-core B operand
    const float* Wimg32;   // ... and the 16 KB TF32 image used by the range-safe fp32 pass (backbone_f32.cu)
    const float* b_host;   // host copy of b (64->64 layers): passed to the ProxyConv kernel by value, i.e. through the constant bank
};
void make_w64_image(const float* W /*[64][64] folded*/, uint16_t* img /*[4096] fp16 bits*/);
void make_w64_image_f32(const float* W /*[64][64] folded, TF32-rounded*/, float* img /*[4096]*/);
struct BlockDev {          // one ProxyConv block (models/epc-net.py:66-81)
    DenseDev conv, conv_a, conv_b;
};
// fp16 fast pass (backbone.cu): flags[b] is set when an activation of cloud b left the fp16 range
int conv_in(const float4* sorted, int B, int N, const DenseDev& L, uint16_t* x, int* flags, cudaStream_t st);
// concat32 / concat16: the block's 64-channel output is written into column slice [coff, coff+64) of the fp32
// and/or bf16 concat buffer (either may be nullptr)
int proxy_block(const uint16_t* x, const KnnState& g, int B, int N, int arith, float divisor, const DenseDev& conv_a,
                const DenseDev& conv_b, const DenseDev* conv_next, float* concat32, __nv_bfloat16* concat16, int ctot,
                int coff, uint16_t* xnext, int* flags, float* cloud_absmax, int concat_f16, cudaStream_t st);
// range-safe fp32/TF32 pass over the flagged clouds only (backbone_f32.cu)
int conv_in_f32(const float4* sorted, int B, int N, const DenseDev& L, float* x, const int* flags, cudaStream_t st);
int proxy_block_f32(const int* flags, const float* x, const KnnState& g, int B, int N, int arith, float divisor, const DenseDev& conv_a,
                    const DenseDev& conv_b, const DenseDev* conv_next, float* concat32, __nv_bfloat16* concat16, int ctot,
                    int coff, float* xnext, float* cloud_absmax, cudaStream_t st);

// ---- gemm.cu -----------------------------------------------------------------------------------
struct GemmArgs {
    const float* A;        // element (m,k) at A[m*sAm + k*sAk]
    const float* B;        // element (k,n) at B[k*sBk + n*sBn]
    float* C;              // row-major [M, ldc]
    int M, N, K;
    long long sAm, sAk, sBk, sBn;
    int ldc;
    int batch;             // grid.z batches with the strides below
    long long bA, bB, bC;
    const float* bias;     // [N] or nullptr
    int relu;
    int splitk;            // >1: split s writes its partial sums to the slab C + s*slab (no bias/relu; no atomics)
    long long slab;        // elements between split-K slabs
};
int sgemm(const GemmArgs& g, cudaStream_t st);

// ---- vlad.cu -----------------------------------------------------------------------------------
int row_inv_norm(const float* X, long long R, int F, float* inv, cudaStream_t st);
// V: nslab split-K slabs of [B,F,K] (slab elements apart); a_sum: [B, a_parts, K] partial column sums
int vlad_finalize(const float* V, int nslab, long long slab, const float* vscale, const float* a_sum, int a_parts, const float* Wc2,
                  int B, int F, int K, float* v, float* colss, cudaStream_t st);
constexpr int HIDDEN_SPLITK = 32;  // hidden FC split-K slabs [HIDDEN_SPLITK, B*G, D]
int vlad_splitk();                 // VLAD accumulate split-K slabs (api.cu; EPC_VLAD_SPLITK = 1 | 2 | 4 | 8)

// ---- tc_gemm.cu (tcgen05 tensor-core contractions) ---------------------------------------------------
int tc_conv5_bf16(const __nv_bfloat16* Xc, long long R, int cin, const __nv_bfloat16* W5t, const float* b5,
                  __nv_bfloat16* H, float* rowss, cudaStream_t st);
constexpr int CONV5_ROWSS_PARTS = 8;   // 1024 / BN(256) N tiles x 2 epilogue warps per lane quarter
int conv5_rowss_parts();
int tc_assign(const __nv_bfloat16* H, long long R, const __nv_bfloat16* Wct, const float* rowss, int parts,
              const float* bn_scale, const float* bn_shift, __nv_bfloat16* S, float* a_part, cudaStream_t st);
// head_fused.cu: both of the above in one launch (VLAD's read of H is served by the L2)
int tc_assign_vlad(const __nv_bfloat16* H, int clouds, int N, const __nv_bfloat16* Wct, const float* rowss, int parts,
                   const float* bn_scale, const float* bn_shift, __nv_bfloat16* S, float* a_part, float* V, int splitk, long long slab,
                   int* ready, cudaStream_t st);
// head_fp8.cu: the same head on an fp8 (e4m3) copy of H with exact power-of-two scales
int tc_conv5_fp8(const __nv_bfloat16* Xc, long long R, int cin, int rows_per_cloud, const __nv_bfloat16* W5t, const float* b5,
                 const float* b5_host, const float* cloud_absmax_dev, float l1max, float bmax, uint8_t* H8, float* rowss, cudaStream_t st);
int sprime_scale(const float* rowss, int parts, int clouds, int N, float* t, float* t_inv, cudaStream_t st);
int f32_to_fp8_rows(const float* X, long long R, int F, int rows_per_cloud, uint8_t* Y, float* rowss, cudaStream_t st);
int tc_assign_vlad_fp8(const uint8_t* H8, int clouds, int N, const uint8_t* Wct8, const float* rowss, int parts, const float* bn_scale,
                       const float* bn_shift, const float* sscale, uint8_t* S8, float* a_part, float* V, int splitk, long long slab,
                       int* ready, cudaStream_t st);
int tc_vlad(const __nv_bfloat16* H, const __nv_bfloat16* S, int B, int N, float* V, int splitk, long long slab,
            cudaStream_t st);
int tc_conv5_colmax_bf16(const __nv_bfloat16* Xc, long long R, int cin, int rows_per_cloud, const __nv_bfloat16* W5t, const float* b5,
                         float* g, int clouds, cudaStream_t st);
// fp16 operands (10-bit mantissa = TF32 precision, twice the tensor rate); cloud_mask: only clouds with a non-zero entry
int tc_conv5_colmax_f16(const __half* Xc, long long R, int cin, int rows_per_cloud, const __half* W5t, const float* b5, float* g, int clouds,
                        cudaStream_t st);
int reset_rows_flagged(float* g, const int* flags, int clouds, int cols, cudaStream_t st);
// cloud_mask != NULL: only the clouds with a non-zero entry are processed, and g is NOT cleared (reset_rows_flagged does that)
int tc_conv5_colmax(const float* Xc, long long R, int cin, int rows_per_cloud, const float* W5t, const float* b5,
                    float* g, int clouds, const int* cloud_mask, cudaStream_t st);
int tc_hidden(const float* v, int rows, int hidden_in, const float* Wht, int D, float* Y, int splitk, cudaStream_t st);
int tc_fc_relu(const float* g, int rows, int K, const float* Wt, const float* bias, int D, float* out, cudaStream_t st);
int tc_conv5_f32(const float* Xc, long long R, int cin, const float* W5t, const float* b5, float* H, cudaStream_t st);
int vlad_tail(const float* Y, int nslab, int B, int G, int D, const float* bn_scale, const float* bn_shift, const float* Wg,
              const float* g_scale, const float* g_shift, int gating, int l2, float* out, cudaStream_t st);
int col_max(const float* H, int B, int N, int F, float* g, cudaStream_t st);
int row_l2_normalize(const float* X, int R, int D, float* out, cudaStream_t st);
int f32_to_bf16_rows(const float* X, long long R, int F, __nv_bfloat16* Y, float* rowss, cudaStream_t st);
int kd_feat(const float* H, const float* inv, const int* perm, int B, int N, int F, float* feat, cudaStream_t st);

// ---- retrieval.cu ------------------------------------------------------------------------------
size_t retrieve_workspace_bytes(int D, int Q, int dim, int k);
size_t retrieve_index_bytes(int D, int dim);
int retrieve_index_build(const float* db, int D, int dim, void* index_mem, size_t index_bytes, cudaStream_t st);
// index: memory prepared by retrieve_index_build for this (db, D, dim), or nullptr (built in the workspace)
int retrieve_topk(const float* db, int D, const float* q, int Q, int dim, int k, long long id_offset, const void* index,
                  int64_t* idx, double* dist, void* ws, size_t ws_bytes, cudaStream_t st);
// ---- retrieval_tc.cu ---------------------------------------------------------------------------
bool retr_tc_supported(int dim);
int retr_ranges(int Q, int n_tiles);
int retr_split2(const float* X, int R, int Rpad, int dim, __nv_bfloat16* out, float* norms, cudaStream_t st);
int retr_scores(const __nv_bfloat16* q2, int Q, const __nv_bfloat16* db2, int D, int dim, int n_tiles, int tile_stride, int n_ranges,
                const float* dn, float* scores, int ld, const float* thr, uint2* cand, int* cand_count, int cap, cudaStream_t st);
int radius_search(const double* db, int D, const double* q, int Q, int dim, double r, int32_t* counts, const int64_t* offsets,
                  int32_t* indices, cudaStream_t st);
int merge_topk(const double* dist, const int64_t* idx, long long rank_stride, int R, int Q, int k, double* out_dist, int64_t* out_idx,
               cudaStream_t st);

}  // namespace epc
