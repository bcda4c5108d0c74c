// C ABI of libepc_b200.so (include/epc_b200.h): argument checking, BN folding / weight upload,
// workspace carving and the kernel sequence of one embedding call.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"

namespace epc {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

// ---- per-stage timing -------------------------------------------------------------------------------
struct StageRec {
    int id;
    cudaEvent_t a, b;
};
static thread_local bool g_prof_on = false;
static thread_local std::vector<StageRec>* g_prof = nullptr;
static thread_local double g_stage_ms[EPC_STAGE_COUNT] = {};
static thread_local long long g_stage_n[EPC_STAGE_COUNT] = {};

ScopedStage::ScopedStage(int stage_id, cudaStream_t stream) : id(stage_id), st(stream), on(g_prof_on) {
    if (!on) return;
    if (!g_prof) g_prof = new std::vector<StageRec>();
    StageRec r;
    r.id = id;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) {
        on = false;
        return;
    }
    cudaEventRecord(r.a, st);
    g_prof->push_back(r);
}
ScopedStage::~ScopedStage() {
    if (!on) return;
    cudaEventRecord(g_prof->back().b, st);
}

static void prof_collect() {
    if (!g_prof) return;
    for (StageRec& r : *g_prof) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            g_stage_ms[r.id] += ms;
            g_stage_n[r.id] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof->clear();
}

}  // namespace epc

using namespace epc;

struct EpcModel {
    int arch = 0, n_blocks = 0, K = 64, D = 256, G = 4, pooling = 0, gating = 1;
    float divisor = 20.f;
    bool vlad_head = true;
    float* blob = nullptr;            // one device allocation holding every folded parameter
    DenseDev conv[12];
    const float *W5t = nullptr, *b5 = nullptr;         // BN-folded conv5, transposed [1024, cin] (fp32: TF32 operand)
    __nv_bfloat16* blob16 = nullptr;                    // bf16 operands of the EPC-Net head
    const __nv_bfloat16 *W5t16 = nullptr, *Wct16 = nullptr;   // [1024, cin], [64, 1024]
    const float *cbn_scale = nullptr, *cbn_shift = nullptr, *Wc2 = nullptr, *Wh = nullptr,
                *hbn_scale = nullptr, *hbn_shift = nullptr, *Wg = nullptr, *gbn_scale = nullptr, *gbn_shift = nullptr;
    __half* blobf16 = nullptr;                          // conv5 as an fp16 operand [1024, cin] (EPC-Net-L)
    uint8_t* blob8 = nullptr;                           // fp8 (e4m3) operands of the fp8 head (head_fp8.cu)
    const uint8_t* Wct8 = nullptr;                      // 2^w Wc^T [64, 1024]
    const float* cbn_scale8 = nullptr;                  // cluster-BN scale x 2^-w
    float conv_b_host[12][64] = {};                     // host copies of the 64-channel conv biases (ProxyConv kernel arguments)
    float b5_host[1024] = {};                           // host copy of the conv5 bias: passed by value to the fp8 conv5 kernel (constant bank)
    float l1max = 0.f, bmax = 0.f;                      // max_f sum_c |W5[c,f]| (bf16 operand values), max_f |b5[f]|
    int hidden_in = 0;                // rows of hidden1_weights
    DenseDev fc1;
    const float* fc1_Wt = nullptr;      // fc1 weights transposed [D, 1024], TF32-rounded: K-major B operand of the tensor-core FC
};

namespace {

// inference BN -> per-channel affine:  y = x*scale + shift  (utils/tf_util.py:490; FusedBatchNorm is the same map)
void bn_affine(const EpcBN& bn, int C, std::vector<float>& scale, std::vector<float>& shift) {
    scale.resize(C);
    shift.resize(C);
    for (int c = 0; c < C; ++c) {
        const float inv = (1.0f / std::sqrt(bn.var_host[c] + BN_EPS)) * bn.gamma_host[c];
        scale[c] = inv;
        shift[c] = bn.beta_host[c] - bn.mean_host[c] * inv;
    }
}

bool bn_ok(const EpcBN& bn) { return bn.beta_host && bn.gamma_host && bn.mean_host && bn.var_host; }

// y = relu(BN(xW+b)) = relu(x (W*scale) + (b*scale + shift))
void fold_dense(const EpcDense& L, std::vector<float>& W, std::vector<float>& b) {
    std::vector<float> sc, sh;
    bn_affine(L.bn, L.cout, sc, sh);
    W.resize((size_t)L.cin * L.cout);
    b.resize(L.cout);
    for (int k = 0; k < L.cin; ++k)
        for (int n = 0; n < L.cout; ++n) W[(size_t)k * L.cout + n] = L.weights_host[(size_t)k * L.cout + n] * sc[n];
    for (int n = 0; n < L.cout; ++n) b[n] = L.biases_host[n] * sc[n] + sh[n];
}

// round-to-nearest-even to TF32 precision (the tensor core would otherwise truncate)
float host_round_tf32(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u += 0x0fffu + ((u >> 13) & 1u);
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}

struct Packer {
    std::vector<float> host;
    size_t add(const float* p, size_t n) {
        size_t off = (host.size() + 63) / 64 * 64;   // 256-byte alignment
        host.resize(off + n);
        memcpy(host.data() + off, p, n * sizeof(float));
        return off;
    }
    size_t add(const std::vector<float>& v) { return add(v.data(), v.size()); }
};

// libepc_b200 links cudart statically; make the device that owns `p` current for this thread so that
// launches go to the same device as the caller's (e.g. torch's) allocations.
int ensure_device(const void* p) {
    if (!p) return EPC_OK;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        set_error("cudaPointerGetAttributes: %s (is a CUDA device present?)", cudaGetErrorString(e));
        return EPC_ECUDA;
    }
    if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) {
        cudaGetLastError();
        set_error("pointer %p is not device memory (libepc_b200 has no CPU path)", p);
        return EPC_EINVAL;
    }
    e = cudaSetDevice(at.device);
    if (e != cudaSuccess) {
        set_error("cudaSetDevice(%d): %s", at.device, cudaGetErrorString(e));
        return EPC_ECUDA;
    }
    return EPC_OK;
}

bool dense_ok(const EpcDense& L) { return L.weights_host && L.biases_host && bn_ok(L.bn) && L.cin > 0 && L.cout > 0; }

}  // namespace

extern "C" {

const char* epc_last_error(void) { return g_err; }
int epc_abi_version(void) { return EPC_ABI_VERSION; }
long long epc_launch_count(void) { return g_launches; }
void epc_launch_count_reset(void) { g_launches = 0; }

void epc_profile_enable(int on) { g_prof_on = (on != 0); }
void epc_profile_reset(void) {
    prof_collect();
    for (int i = 0; i < EPC_STAGE_COUNT; ++i) {
        g_stage_ms[i] = 0.0;
        g_stage_n[i] = 0;
    }
}
int epc_profile_read(int stage, double* ms, long long* launches) {
    EPC_CHECK_ARG(stage >= 0 && stage < EPC_STAGE_COUNT, "bad stage %d", stage);
    prof_collect();
    if (ms) *ms = g_stage_ms[stage];
    if (launches) *launches = g_stage_n[stage];
    return EPC_OK;
}
const char* epc_stage_name(int stage) {
    static const char* names[EPC_STAGE_COUNT] = {"sort", "knn", "conv_in", "proxy_block", "conv5", "rownorm",
                                                 "assign_gemm", "assign_softmax", "vlad_gemm", "vlad_finalize",
                                                 "hidden_gemm", "tail", "colmax", "fc", "kd_feat", "retrieve_score",
                                                 "retrieve_select", "retrieve_rerank", "proxy_block_safe", "assign_vlad"};
    return (stage >= 0 && stage < EPC_STAGE_COUNT) ? names[stage] : "?";
}

int epc_set_device(int device) {
    EPC_CUDA(cudaSetDevice(device));
    return EPC_OK;
}
int epc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// -------------------------------------------------------------------------------------------------
// kNN
// -------------------------------------------------------------------------------------------------
size_t epc_knn_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 256;
    return knn_state_bytes(B, N) + align_up((size_t)B * N * sizeof(float)) + 256;
}

int epc_knn(const float* xyz, int B, int N, int arith, int32_t* idx, float* kth, int32_t* count, void* workspace,
            size_t workspace_bytes, void* stream) {
    EPC_CHECK_ARG(B >= 0 && xyz != nullptr || B == 0, "epc_knn: xyz is NULL");
    if (int rc = ensure_device(xyz)) return rc;
    if (int rc = knn_check_n(N)) return rc;
    if (workspace_bytes < epc_knn_workspace_bytes(B, N) || !workspace) {
        set_error("epc_knn: workspace %zu < required %zu", workspace_bytes, epc_knn_workspace_bytes(B, N));
        return EPC_EWORKSPACE;
    }
    Arena ar(workspace, workspace_bytes);
    KnnState s = knn_state_carve(ar, B, N);
    return knn_build(xyz, B, N, arith, true, s, idx, kth, count, static_cast<cudaStream_t>(stream));
}

// test hook: same as epc_knn with the AABB pruning switched off (results must be bit-identical)
int epc_knn_noprune(const float* xyz, int B, int N, int arith, int32_t* idx, float* kth, int32_t* count, void* workspace,
                    size_t workspace_bytes, void* stream) {
    if (int rc = ensure_device(xyz)) return rc;
    if (int rc = knn_check_n(N)) return rc;
    if (workspace_bytes < epc_knn_workspace_bytes(B, N) || !workspace) {
        set_error("epc_knn_noprune: workspace too small");
        return EPC_EWORKSPACE;
    }
    Arena ar(workspace, workspace_bytes);
    KnnState s = knn_state_carve(ar, B, N);
    return knn_build(xyz, B, N, arith, false, s, idx, kth, count, static_cast<cudaStream_t>(stream));
}

int epc_knn_dense(const float* xyz, int B, int N, int arith, float* mask, float* dist, void* workspace,
                  size_t workspace_bytes, void* stream) {
    if (int rc = ensure_device(xyz)) return rc;
    if (int rc = knn_check_n(N)) return rc;
    if (workspace_bytes < epc_knn_workspace_bytes(B, N) || !workspace) {
        set_error("epc_knn_dense: workspace %zu < required %zu", workspace_bytes, epc_knn_workspace_bytes(B, N));
        return EPC_EWORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Arena ar(workspace, workspace_bytes);
    KnnState s = knn_state_carve(ar, B, N);
    float* kth = ar.take<float>((size_t)B * N);
    if (mask) {
        if (int rc = knn_build(xyz, B, N, arith, true, s, nullptr, kth, nullptr, st))
            return rc;
    }
    return knn_dense(xyz, B, N, arith, kth, mask, dist, st);
}

int epc_rows_topk_smallest(const float* adj, long long R, int M, int k, int32_t* idx, void* stream) {
    if (int rc = ensure_device(adj)) return rc;
    return rows_topk_smallest(adj, R, M, k, idx, static_cast<cudaStream_t>(stream));
}

// -------------------------------------------------------------------------------------------------
// stand-alone layers
// -------------------------------------------------------------------------------------------------
int epc_dense_forward(const EpcDense* layer, const float* x, long long R, float* y, int relu, void* stream) {
    EPC_CHECK_ARG(layer && dense_ok(*layer), "epc_dense_forward: incomplete layer description");
    EPC_CHECK_ARG(R >= 0 && R < (1ll << 31), "epc_dense_forward: bad row count");
    if (int rc = ensure_device(x)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<float> W, b;
    fold_dense(*layer, W, b);
    float* dW = nullptr;
    EPC_CUDA(cudaMalloc(&dW, (W.size() + b.size()) * sizeof(float)));
    float* db = dW + W.size();
    cudaError_t e = cudaMemcpyAsync(dW, W.data(), W.size() * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice, st);
    int rc = EPC_OK;
    if (e != cudaSuccess) {
        set_error("epc_dense_forward: upload failed: %s", cudaGetErrorString(e));
        rc = EPC_ECUDA;
    } else {
        GemmArgs g = {};
        g.A = x; g.sAm = layer->cin; g.sAk = 1;
        g.B = dW; g.sBk = layer->cout; g.sBn = 1;
        g.C = y; g.ldc = layer->cout; g.M = (int)R; g.N = layer->cout; g.K = layer->cin;
        g.bias = db; g.relu = relu; g.batch = 1; g.splitk = 1;
        rc = sgemm(g, st);
    }
    cudaStreamSynchronize(st);   // W/b are pageable host vectors that die with this frame
    cudaFree(dW);
    return rc;
}

int epc_max_pool_points(const float* x, int B, int N, int C, float* y, void* stream) {
    EPC_CHECK_ARG(x && y && B >= 0 && N > 0 && C > 0, "epc_max_pool_points: bad arguments");
    if (int rc = ensure_device(x)) return rc;
    return col_max(x, B, N, C, y, static_cast<cudaStream_t>(stream));
}

// -------------------------------------------------------------------------------------------------
// model
// -------------------------------------------------------------------------------------------------
int epc_model_create(const EpcWeights* w, EpcModel** out) {
    EPC_CHECK_ARG(w && out, "epc_model_create: NULL argument");
    EPC_CHECK_ARG(w->arch >= EPC_ARCH_EPC_NET && w->arch <= EPC_ARCH_KD_EPC_NET_L, "unknown arch %d", w->arch);
    const bool vlad = (w->arch == EPC_ARCH_EPC_NET || w->arch == EPC_ARCH_KD_EPC_NET);
    const int nb = w->n_blocks;
    EPC_CHECK_ARG(nb >= 1 && nb <= 4, "n_blocks=%d unsupported (1..4)", nb);
    EPC_CHECK_ARG(w->knn_k > 0, "knn_k (the divisor of the neighbour mean) must be positive");
    EPC_CHECK_ARG(w->output_dim >= 1 && w->output_dim <= 1024, "output_dim=%d unsupported", w->output_dim);
    for (int i = 0; i < 3 * nb; ++i) {
        EPC_CHECK_ARG(dense_ok(w->conv[i]), "conv layer %d incomplete", i);
        EPC_CHECK_ARG(w->conv[i].cout == 64 && w->conv[i].cin == (i == 0 ? 3 : 64), "conv layer %d: shape %d->%d", i,
                      w->conv[i].cin, w->conv[i].cout);
    }
    EPC_CHECK_ARG(dense_ok(w->conv5) && w->conv5.cin == 64 * nb && w->conv5.cout == 1024, "conv5 must be %d->1024",
                  64 * nb);

    EpcModel* m = new EpcModel();
    m->arch = w->arch; m->n_blocks = nb; m->K = w->cluster_size; m->D = w->output_dim; m->G = w->groups;
    m->pooling = w->pooling; m->gating = w->gating; m->divisor = (float)w->knn_k; m->vlad_head = vlad;

    Packer pk;
    std::vector<float> W, b, sc, sh;
    size_t offW[13], offb[13], offI[13];
    std::vector<uint16_t> img(4096);
    std::vector<float> img32(4096);
    size_t offI32[13];
    for (int i = 0; i < 3 * nb; ++i) {
        fold_dense(w->conv[i], W, b);
        offW[i] = pk.add(W); offb[i] = pk.add(b);
        if (i < 12 && b.size() == 64) memcpy(m->conv_b_host[i], b.data(), sizeof(float) * 64);
        offI[i] = 0;
        offI32[i] = 0;
        if (i > 0) {
            make_w64_image(W.data(), img.data());
            offI[i] = pk.add(reinterpret_cast<const float*>(img.data()), img.size() / 2);
            for (auto& x : W) x = host_round_tf32(x);
            make_w64_image_f32(W.data(), img32.data());
            offI32[i] = pk.add(img32);
        }
    }
    fold_dense(w->conv5, W, b);
    const int c5 = 64 * nb;
    std::vector<float> W5t((size_t)1024 * c5);                 // [cout, cin]: the K-major B operand of the tensor-core conv5
    for (int k = 0; k < c5; ++k)
        for (int n = 0; n < 1024; ++n) W5t[(size_t)n * c5 + k] = host_round_tf32(W[(size_t)k * 1024 + n]);
    offW[12] = pk.add(W5t); offb[12] = pk.add(b);
    std::vector<__nv_bfloat16> h16;                              // bf16 operands (EPC-Net head)
    size_t o16_W5 = 0, o16_Wc = 0;
    std::vector<uint8_t> h8;                                     // fp8 operands (fp8 head)
    size_t oCs8 = 0;
    size_t oCs = 0, oCh = 0, oWc2 = 0, oWh = 0, oHs = 0, oHh = 0, oWg = 0, oGs = 0, oGh = 0, oFW = 0, oFb = 0, oFWt = 0;
    h16.resize((size_t)1024 * c5 + (vlad ? (size_t)64 * 1024 : 0));    // conv5 as a bf16 operand: both heads
    for (size_t i = 0; i < W5t.size(); ++i) h16[i] = __float2bfloat16(W5t[i]);
    if (vlad) {
        const int K = w->cluster_size, D = w->output_dim;
        if (K != 64) { delete m; set_error("cluster_size=%d unsupported: the tensor-core assignment kernel is built for 64", K); return EPC_EUNSUPPORTED; }
        const bool gv = (w->pooling == EPC_POOL_G_VLAD);
        if (gv && !(w->groups >= 1 && (1024 * K) % w->groups == 0)) { delete m; set_error("bad groups=%d", w->groups); return EPC_EINVAL; }
        if (!gv) m->G = 1;
        m->hidden_in = 1024 * K / m->G;
        if (!(w->cluster_weights_host && bn_ok(w->cluster_bn) && w->cluster_weights2_host && w->hidden1_weights_host &&
              bn_ok(w->hidden_bn) && (!w->gating || (w->gating_weights_host && bn_ok(w->gating_bn))))) {
            delete m; set_error("VLAD head weights incomplete"); return EPC_EINVAL;
        }
        o16_Wc = W5t.size();
        for (int f = 0; f < 1024; ++f)                          // Wc^T [K, 1024]: K-major B operand of the assignment GEMM
            for (int c = 0; c < K; ++c) h16[o16_Wc + (size_t)c * 1024 + f] = __float2bfloat16(w->cluster_weights_host[(size_t)f * K + c]);
        bn_affine(w->cluster_bn, K, sc, sh); oCs = pk.add(sc); oCh = pk.add(sh);
        {   // fp8 head: Wc' = 2^wexp Wc as e4m3 (|Wc'| <= 256), 2^-wexp folded into the cluster-BN scale; conv5 output bound
            float wmax = 0.f;
            for (size_t i = 0; i < (size_t)1024 * K; ++i) wmax = std::max(wmax, std::fabs(w->cluster_weights_host[i]));
            int ex = 0;
            std::frexp(wmax + 1e-30f, &ex);
            const float ws = std::ldexp(1.0f, 8 - ex);
            h8.resize((size_t)64 * 1024);
            for (int f = 0; f < 1024; ++f)
                for (int c = 0; c < K; ++c)
                    h8[(size_t)c * 1024 + f] = (uint8_t)__nv_cvt_float_to_fp8(w->cluster_weights_host[(size_t)f * K + c] * ws, __NV_SATFINITE, __NV_E4M3);
            std::vector<float> sc8(K);
            for (int c = 0; c < K; ++c) sc8[c] = sc[c] / ws;
            oCs8 = pk.add(sc8);
            for (int n = 0; n < 1024; ++n) {
                float l1 = 0.f;
                for (int k = 0; k < c5; ++k) l1 += std::fabs(__bfloat162float(h16[(size_t)n * c5 + k]));
                m->l1max = std::max(m->l1max, l1);
                m->bmax = std::max(m->bmax, std::fabs(b[n]));
                m->b5_host[n] = b[n];
            }
        }
        oWc2 = pk.add(w->cluster_weights2_host, (size_t)1024 * K);
        {   // hidden1_weights [hidden_in, D] -> transposed [D, hidden_in], TF32-rounded: K-major B operand of the tensor-core FC
            std::vector<float> Wht((size_t)D * m->hidden_in);
            for (int k = 0; k < m->hidden_in; ++k)
                for (int n = 0; n < D; ++n) Wht[(size_t)n * m->hidden_in + k] = host_round_tf32(w->hidden1_weights_host[(size_t)k * D + n]);
            oWh = pk.add(Wht);
        }
        bn_affine(w->hidden_bn, D, sc, sh); oHs = pk.add(sc); oHh = pk.add(sh);
        if (w->gating) {
            oWg = pk.add(w->gating_weights_host, (size_t)D * D);
            bn_affine(w->gating_bn, D, sc, sh); oGs = pk.add(sc); oGh = pk.add(sh);
        }
    } else {
        if (!(dense_ok(w->fc1) && w->fc1.cin == 1024 && w->fc1.cout == w->output_dim)) {
            delete m; set_error("fc1 must be 1024->%d", w->output_dim); return EPC_EINVAL;
        }
        fold_dense(w->fc1, W, b);
        oFW = pk.add(W); oFb = pk.add(b);
        {
            const int D = w->output_dim;
            std::vector<float> Wt((size_t)D * 1024);
            for (int k = 0; k < 1024; ++k)
                for (int n = 0; n < D; ++n) Wt[(size_t)n * 1024 + k] = host_round_tf32(W[(size_t)k * D + n]);
            oFWt = pk.add(Wt);
        }
    }
    cudaError_t e = cudaMalloc(&m->blob, pk.host.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(m->blob, pk.host.data(), pk.host.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !h16.empty()) {
        e = cudaMalloc(&m->blob16, h16.size() * sizeof(__nv_bfloat16));
        if (e == cudaSuccess) e = cudaMemcpy(m->blob16, h16.data(), h16.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && !vlad) {       // EPC-Net-L: conv5 weights as fp16 (10-bit mantissa like TF32; |W| is far inside the range)
        std::vector<__half> hf(W5t.size());
        for (size_t i = 0; i < W5t.size(); ++i) hf[i] = __float2half_rn(W5t[i]);
        e = cudaMalloc(&m->blobf16, hf.size() * sizeof(__half));
        if (e == cudaSuccess) e = cudaMemcpy(m->blobf16, hf.data(), hf.size() * sizeof(__half), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && !h8.empty()) {
        e = cudaMalloc(&m->blob8, h8.size());
        if (e == cudaSuccess) e = cudaMemcpy(m->blob8, h8.data(), h8.size(), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        set_error("epc_model_create: %s", cudaGetErrorString(e));
        if (m->blob) cudaFree(m->blob);
        if (m->blob16) cudaFree(m->blob16);
        if (m->blob8) cudaFree(m->blob8);
        if (m->blobf16) cudaFree(m->blobf16);
        delete m;
        return EPC_ECUDA;
    }
    for (int i = 0; i < 3 * nb; ++i)
        m->conv[i] = DenseDev{m->blob + offW[i], m->blob + offb[i], w->conv[i].cin, 64, i > 0 ? m->blob + offI[i] : nullptr,
                              i > 0 ? m->blob + offI32[i] : nullptr, i < 12 ? m->conv_b_host[i] : nullptr};
    m->W5t = m->blob + offW[12];
    m->b5 = m->blob + offb[12];
    m->W5t16 = m->blob16 + o16_W5;
    if (vlad) {
        m->Wct16 = m->blob16 + o16_Wc;
        m->cbn_scale = m->blob + oCs; m->cbn_shift = m->blob + oCh; m->Wc2 = m->blob + oWc2;
        m->Wct8 = m->blob8; m->cbn_scale8 = m->blob + oCs8;
        m->Wh = m->blob + oWh; m->hbn_scale = m->blob + oHs; m->hbn_shift = m->blob + oHh;
        if (w->gating) { m->Wg = m->blob + oWg; m->gbn_scale = m->blob + oGs; m->gbn_shift = m->blob + oGh; }
    } else {
        m->fc1 = DenseDev{m->blob + oFW, m->blob + oFb, 1024, w->output_dim, nullptr, nullptr, nullptr};
        m->fc1_Wt = m->blob + oFWt;
    }
    *out = m;
    return EPC_OK;
}

void epc_model_destroy(EpcModel* m) {
    if (!m) return;
    if (m->blob) cudaFree(m->blob);
    if (m->blob16) cudaFree(m->blob16);
    if (m->blob8) cudaFree(m->blob8);
    if (m->blobf16) cudaFree(m->blobf16);
    delete m;
}

namespace {

// clouds per head sub-batch: the bf16 conv5 output H (8 MiB per cloud) of one sub-batch is produced and consumed by
// the GEMM launches back to back.  Measured on B200 (us/cloud for conv5+assign+VLAD): 8 -> 10.2, 32 -> 8.2, 64 -> 7.5,
// 128 -> 7.1 (round 1, three kernels); conv5 + fused assignment/VLAD, round 2: 64 -> 5.72, 128 -> 5.45, 256 -> 5.49:
// launch/prologue/wave-quantisation costs outweigh keeping H inside the L2.  With the fp8 head and the faster conv5 of the end of
// round 2: 64 -> 4.54, 128 -> 4.14, 256 -> 4.00 (conv5 + assignment/VLAD), so the fp8 head takes a whole 256-cloud call at once
// and the bf16 head keeps 128.  EPC_HEAD_SUB overrides both (tuning aid).
static int head_sub_init() {
    const char* e = getenv("EPC_HEAD_SUB");
    int v = e ? atoi(e) : 0;
    return (v >= 1 && v <= 256) ? v : 256;
}
static const int HEAD_SUB = head_sub_init();       // upper bound (workspace sizing); the step actually used: head_step(N)
// the head's storage format of the per-point features: e4m3 with exact power-of-two scales (head_fp8.cu) unless EPC_HEAD_FP8=0
// -- for clouds of at least 2048 points: the quantisation errors average out over the points of a cloud (measured descriptor
// error at N = 4096: 6e-5, the bf16 head's own level), so the precision budget is spent where the sums are long; smaller clouds
// (unit tests, the stand-alone loupe API at small max_samples) keep the bf16 head
static bool head_fp8(int N) {
    static const bool v = !(getenv("EPC_HEAD_FP8") && atoi(getenv("EPC_HEAD_FP8")) == 0);
    return v && N >= 2048;
}
static int head_step(int N) {
    static const bool from_env = getenv("EPC_HEAD_SUB") != nullptr;
    return (head_fp8(N) || from_env || HEAD_SUB < 128) ? HEAD_SUB : 128;
}


struct HeadWs {                 // buffers of the G_VLAD / NetVLAD head
    __nv_bfloat16* H16;         // [sub*N, 1024]   conv5 output (operand of the assignment and VLAD GEMMs)
    float* rowss;               // [sub*N, 4]      partial |H_n|^2
    __nv_bfloat16* S16;         // [sub*N, 64]     S' = softmax/|H|
    float* a_part;              // [B*N/128, 64]   partial column sums of the soft assignment
    float* V;                   // [vlad_splitk()][B,1024,64]
    float* v;                   // [B, 65536]
    float* Y;                   // [HIDDEN_SPLITK][B*G, D]
    float* colss;               // [B, 8, 64] partial column sums of squares of the VLAD residuals
    int* ready;                 // [sub] per-cloud tile counters of the fused assignment + VLAD launch
    float* absmax;              // [sub] fp8 head: max |conv5 input| per cloud
    float* tscale;              // [sub] fp8 head: power-of-two scale of the cloud's S''
    float* tinv;                // [B]   its inverse, applied when V is finalised
};

size_t head_bytes(const EpcModel* m, int B, int N) {
    const size_t sub = (size_t)(B < HEAD_SUB ? B : HEAD_SUB) * N;
    return align_up(sub * 1024 * 2) + align_up(sub * CONV5_ROWSS_PARTS * 4) + align_up(sub * 64 * 2) +
           align_up((size_t)B * (N / 128) * 64 * 4) + align_up((size_t)vlad_splitk() * B * 1024 * 64 * 4) +
           align_up((size_t)B * 1024 * 64 * 4) + align_up((size_t)HIDDEN_SPLITK * B * m->G * m->D * 4) +
           align_up((size_t)B * 8 * 64 * 4) + 3 * align_up((size_t)(B < HEAD_SUB ? B : HEAD_SUB) * 4) + align_up((size_t)B * 4);
}

HeadWs head_carve(Arena& ar, const EpcModel* m, int B, int N) {
    const size_t sub = (size_t)(B < HEAD_SUB ? B : HEAD_SUB) * N;
    HeadWs h;
    h.H16 = ar.take<__nv_bfloat16>(sub * 1024);
    h.rowss = ar.take<float>(sub * CONV5_ROWSS_PARTS);
    h.S16 = ar.take<__nv_bfloat16>(sub * 64);
    h.a_part = ar.take<float>((size_t)B * (N / 128) * 64);
    h.V = ar.take<float>((size_t)vlad_splitk() * B * 1024 * 64);
    h.v = ar.take<float>((size_t)B * 1024 * 64);
    h.Y = ar.take<float>((size_t)HIDDEN_SPLITK * B * m->G * m->D);
    h.colss = ar.take<float>((size_t)B * 8 * 64);
    h.ready = ar.take<int>((size_t)(B < HEAD_SUB ? B : HEAD_SUB));
    h.absmax = ar.take<float>((size_t)(B < HEAD_SUB ? B : HEAD_SUB));
    h.tscale = ar.take<float>((size_t)(B < HEAD_SUB ? B : HEAD_SUB));
    h.tinv = ar.take<float>((size_t)B);
    return h;
}

// assignment + VLAD accumulate for clouds [b0, b0+nb) whose bf16 features and rowss are in h.H16 / h.rowss
int head_assign_vlad(const EpcModel* m, int B, int N, int b0, int nb, const HeadWs& h, int rowss_parts, cudaStream_t st) {
    const long long Rs = (long long)nb * N;
    // one launch for both GEMMs: VLAD's read of H comes from the L2 the assignment CTAs filled (head_fused.cu);
    // EPC_HEAD_FUSED=0 selects the two separate kernels (same results bit for bit: identical tiles and summation order)
    static const bool fused = !(getenv("EPC_HEAD_FUSED") && atoi(getenv("EPC_HEAD_FUSED")) == 0);
    if (head_fp8(N)) {      // H16 / S16 hold the e4m3 tensors (head_fp8.cu): rowss are those of the scaled H'
        ScopedStage ss(EPC_STAGE_ASSIGN_VLAD, st);
        if (int rc = sprime_scale(h.rowss, rowss_parts, nb, N, h.tscale, h.tinv + b0, st)) return rc;
        return tc_assign_vlad_fp8(reinterpret_cast<const uint8_t*>(h.H16), nb, N, m->Wct8, h.rowss, rowss_parts, m->cbn_scale8, m->cbn_shift,
                                  h.tscale, reinterpret_cast<uint8_t*>(h.S16), h.a_part + (size_t)b0 * (N / 128) * 64,
                                  h.V + (size_t)b0 * 1024 * 64, vlad_splitk(), (long long)B * 1024 * 64, h.ready, st);
    }
    if (fused) {
        ScopedStage ss(EPC_STAGE_ASSIGN_VLAD, st);
        return tc_assign_vlad(h.H16, nb, N, m->Wct16, h.rowss, rowss_parts, m->cbn_scale, m->cbn_shift, h.S16,
                              h.a_part + (size_t)b0 * (N / 128) * 64, h.V + (size_t)b0 * 1024 * 64, vlad_splitk(),
                              (long long)B * 1024 * 64, h.ready, st);
    }
    {
        ScopedStage ss(EPC_STAGE_ASSIGN_GEMM, st);
        if (int rc = tc_assign(h.H16, Rs, m->Wct16, h.rowss, rowss_parts, m->cbn_scale, m->cbn_shift, h.S16,
                               h.a_part + (size_t)b0 * (N / 128) * 64, st))
            return rc;
    }
    ScopedStage ss(EPC_STAGE_VLAD_GEMM, st);
    return tc_vlad(h.H16, h.S16, nb, N, h.V + (size_t)b0 * 1024 * 64, vlad_splitk(), (long long)B * 1024 * 64, st);
}

// finalise + hidden FC + gating (+ L2) for all B clouds
int head_tail(const EpcModel* m, int B, int N, const HeadWs& h, int l2, float* out, cudaStream_t st) {
    const int D = m->D;
    {
        ScopedStage ss(EPC_STAGE_VLAD_FINALIZE, st);
        if (int rc = vlad_finalize(h.V, vlad_splitk(), (long long)B * 1024 * 64, head_fp8(N) ? h.tinv : nullptr, h.a_part, N / 128, m->Wc2, B, 1024, 64, h.v, h.colss, st))
            return rc;
    }
    {   // hidden FC (loupe.py:302-320): rows of length hidden_in, G per cloud; TF32 tensor cores, split-K slabs
        ScopedStage ss(EPC_STAGE_HIDDEN_GEMM, st);
        if (D % 64 == 0 && m->hidden_in % (HIDDEN_SPLITK * 32) == 0) {
            if (int rc = tc_hidden(h.v, B * m->G, m->hidden_in, m->Wh, D, h.Y, HIDDEN_SPLITK, st)) return rc;
        } else {
            set_error("hidden FC: output_dim=%d (must be a multiple of 64) / hidden_in=%d unsupported by the tensor-core path", D, m->hidden_in);
            return EPC_EUNSUPPORTED;
        }
    }
    ScopedStage ss(EPC_STAGE_TAIL, st);
    return vlad_tail(h.Y, HIDDEN_SPLITK, B, m->G, D, m->hbn_scale, m->hbn_shift, m->Wg, m->gbn_scale, m->gbn_shift,
                     m->gating, l2, out, st);
}

int check_embed_n(int N) {
    if (int rc = knn_check_n(N)) return rc;
    if (N % 128 != 0) {
        set_error("N=%d unsupported by the embedding path: the tensor-core tiles need a multiple of 128 points", N);
        return EPC_EINVAL;
    }
    return EPC_OK;
}

}  // namespace

size_t epc_embed_workspace_bytes(const EpcModel* m, int B, int N) {
    if (!m || B <= 0 || N <= 0) return 256;
    const size_t R = (size_t)B * N;
    const size_t sub = (size_t)(B < HEAD_SUB ? B : HEAD_SUB) * N;
    const int ctot = 64 * m->n_blocks;
    size_t s = knn_state_bytes(B, N) + 2 * align_up(R * 64 * 2) + 2 * align_up(R * 64 * 4) + 2 * align_up((size_t)B * 4) + align_up(R * ctot * 4) /*concat32 (L, or KD export)*/ +
               align_up(R * ctot * 2) /*concat16*/ + align_up(sub * 1024 * 4) /*H32 (KD export)*/ + align_up(sub * 4);
    if (m->vlad_head)
        s += head_bytes(m, B, N);
    else
        s += align_up((size_t)B * 1024 * 4) + align_up((size_t)B * m->D * 4);
    return s + 256;
}

int epc_embed(const EpcModel* m, const float* xyz, int B, int N, int knn_arith, float* out, float* feat, void* workspace,
              size_t workspace_bytes, void* stream) {
    EPC_CHECK_ARG(m && out && (xyz || B == 0), "epc_embed: NULL argument");
    EPC_CHECK_ARG(B >= 0, "epc_embed: B=%d", B);
    if (int rc = ensure_device(xyz)) return rc;
    if (int rc = check_embed_n(N)) return rc;
    if (B == 0) return EPC_OK;
    if (!workspace || workspace_bytes < epc_embed_workspace_bytes(m, B, N)) {
        set_error("epc_embed: workspace %zu < required %zu", workspace_bytes, epc_embed_workspace_bytes(m, B, N));
        return EPC_EWORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t R = (size_t)B * N;
    EPC_CHECK_ARG(R < (1ull << 31), "too many points in one call (B*N = %zu)", R);
    const int nb = m->n_blocks, ctot = 64 * nb;
    const int subB = B < HEAD_SUB ? B : HEAD_SUB;
    // EPC-Net-L conv5 operands: fp16 (default; EPC_L_F16=0 -> TF32 on an fp32 concat).  fp16 keeps TF32's 10-bit mantissa, so the
    // precision is unchanged, but the blocks write a 16-bit concat instead of an fp32 one and the tensor rate doubles; clouds that
    // leave the fp16 range (flags) are re-done below by the TF32 kernel on the fp32 rows of the range-safe pass.
    // EPC_L_BF16=1 selects bf16 operands: a little faster still, but the descriptor error grows 1.6e-4 -> 5e-4 and one
    // configuration leaves the tolerance (1.04e-3): the max-pool picks extremes, nothing averages the rounding out.  Rejected.
    static const bool l_bf16 = getenv("EPC_L_BF16") && atoi(getenv("EPC_L_BF16")) != 0;
    static const bool l_f16 = !l_bf16 && !(getenv("EPC_L_F16") && atoi(getenv("EPC_L_F16")) == 0);
    const bool lite16 = !m->vlad_head && (l_bf16 || l_f16);
    const bool want32 = (!m->vlad_head && !lite16) || feat != nullptr;     // fp32 concat of the fast pass: TF32 conv5 and the KD feature export
    const bool want16 = m->vlad_head || lite16;
    const int concat_f16 = (!m->vlad_head && l_f16) ? 1 : 0;
    Arena ar(workspace, workspace_bytes);
    KnnState ks = knn_state_carve(ar, B, N);
    uint16_t* xa = ar.take<uint16_t>(R * 64);
    uint16_t* xb = ar.take<uint16_t>(R * 64);
    float* xa32 = ar.take<float>(R * 64);                      // fp32 activations of the range-safe pass (flagged clouds only)
    float* xb32 = ar.take<float>(R * 64);
    int* flags = ar.take<int>(B);                              // clouds whose activations left the fp16 range
    float* cabs_buf = ar.take<float>(B);                       // fp8 head: per-cloud max of the conv5 input
    float* concat32 = ar.take<float>(R * ctot);
    __nv_bfloat16* concat16 = ar.take<__nv_bfloat16>(R * ctot);
    float* H32 = ar.take<float>((size_t)subB * N * 1024);
    float* inv = ar.take<float>((size_t)subB * N);

    if (int rc = knn_build(xyz, B, N, knn_arith, true, ks, nullptr, nullptr, nullptr, st))
        return rc;
    // ProxyConv chain: fp16 fast pass over every cloud, then the range-safe fp32 pass over the clouds it flagged
    EPC_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)B, st));
    float* cabsmax = (m->vlad_head && head_fp8(N)) ? cabs_buf : nullptr;     // per-cloud max of the conv5 input, tracked by the blocks
    if (cabsmax) EPC_CUDA(cudaMemsetAsync(cabsmax, 0, sizeof(float) * (size_t)B, st));
    {
        {
            ScopedStage ss(EPC_STAGE_CONV_IN, st);
            if (int rc = conv_in(ks.sorted, B, N, m->conv[0], xa, flags, st)) return rc;
        }
        uint16_t *cur = xa, *nxt = xb;
        for (int blk = 0; blk < nb; ++blk) {
            const DenseDev* next = (blk + 1 < nb) ? &m->conv[3 * (blk + 1)] : nullptr;
            {
                ScopedStage ss(EPC_STAGE_BLOCK, st);
                if (int rc = proxy_block(cur, ks, B, N, knn_arith, m->divisor, m->conv[3 * blk + 1], m->conv[3 * blk + 2],
                                         next, want32 ? concat32 : nullptr, want16 ? concat16 : nullptr, ctot,
                                         64 * blk, nxt, flags, cabsmax, concat_f16, st))
                    return rc;
            }
            uint16_t* t = cur; cur = nxt; nxt = t;
        }
    }
    {
        ScopedStage ss(EPC_STAGE_BLOCK_SAFE, st);
        if (int rc = conv_in_f32(ks.sorted, B, N, m->conv[0], xa32, flags, st)) return rc;
        float *cur = xa32, *nxt = xb32;
        for (int blk = 0; blk < nb; ++blk) {
            const DenseDev* next = (blk + 1 < nb) ? &m->conv[3 * (blk + 1)] : nullptr;
            if (int rc = proxy_block_f32(flags, cur, ks, B, N, knn_arith, m->divisor, m->conv[3 * blk + 1], m->conv[3 * blk + 2],
                                         next, (want32 || concat_f16) ? concat32 : nullptr, (want16 && !concat_f16) ? concat16 : nullptr, ctot,
                                         64 * blk, nxt, cabsmax, st))
                return rc;
            float* t = cur; cur = nxt; nxt = t;
        }
    }
    // KD feature export (models/kd_epc-net.py:158): l2norm(relu(BN(conv5))) per point, fp32 via TF32 tensor cores
    auto export_feat = [&](int b0, int nbs) -> int {
        const long long Rs = (long long)nbs * N;
        const float* xc = concat32 + (size_t)b0 * N * ctot;
        {
            ScopedStage ss(EPC_STAGE_CONV5, st);
            if (int rc = tc_conv5_f32(xc, Rs, ctot, m->W5t, m->b5, H32, st)) return rc;
        }
        {
            ScopedStage ss(EPC_STAGE_ROWNORM, st);
            if (int rc = row_inv_norm(H32, Rs, 1024, inv, st)) return rc;
        }
        ScopedStage ss(EPC_STAGE_KD_FEAT, st);
        return kd_feat(H32, inv, ks.perm + (size_t)b0 * N, nbs, N, 1024, feat + (size_t)b0 * N * 1024, st);
    };
    if (m->vlad_head) {
        HeadWs h = head_carve(ar, m, B, N);
        const int step = head_step(N);
        for (int b0 = 0; b0 < B; b0 += step) {
            const int nbs = (B - b0 < step) ? (B - b0) : step;
            {   // conv5 (models/epc-net.py:136-139) on bf16 tensor cores; H stays in L2 for the next two GEMMs
                ScopedStage ss(EPC_STAGE_CONV5, st);
                if (head_fp8(N)) {
                    if (int rc = tc_conv5_fp8(concat16 + (size_t)b0 * N * ctot, (long long)nbs * N, ctot, N, m->W5t16, m->b5, m->b5_host, cabsmax + b0,
                                              m->l1max, m->bmax, reinterpret_cast<uint8_t*>(h.H16), h.rowss, st))
                        return rc;
                } else if (int rc = tc_conv5_bf16(concat16 + (size_t)b0 * N * ctot, (long long)nbs * N, ctot, m->W5t16, m->b5, h.H16,
                                                  h.rowss, st))
                    return rc;
            }
            if (int rc = head_assign_vlad(m, B, N, b0, nbs, h, conv5_rowss_parts(), st)) return rc;
            if (feat)
                if (int rc = export_feat(b0, nbs)) return rc;
        }
        if (int rc = head_tail(m, B, N, h, /*l2=*/1, out, st)) return rc;
    } else {
        float* gmax = ar.take<float>((size_t)B * 1024);
        float* o = ar.take<float>((size_t)B * m->D);
        {   // conv5 + global max-pool fused (models/epc-net-l.py:84-91): the 16 MiB/cloud activation is never written
            ScopedStage ss(EPC_STAGE_CONV5, st);
            if (l_bf16) {
                if (int rc = tc_conv5_colmax_bf16(concat16, (long long)R, ctot, N, m->W5t16, m->b5, gmax, B, st)) return rc;
            } else if (l_f16) {
                if (int rc = tc_conv5_colmax_f16(reinterpret_cast<const __half*>(concat16), (long long)R, ctot, N, m->blobf16, m->b5, gmax, B, st))
                    return rc;
                // clouds outside the fp16 range: their fp16 rows are garbage -> TF32 on the fp32 rows of the safe pass (usually none)
                if (int rc = reset_rows_flagged(gmax, flags, B, 1024, st)) return rc;
                if (int rc = tc_conv5_colmax(concat32, (long long)R, ctot, N, m->W5t, m->b5, gmax, B, flags, st)) return rc;
            } else if (int rc = tc_conv5_colmax(concat32, (long long)R, ctot, N, m->W5t, m->b5, gmax, B, nullptr, st)) return rc;
        }
        {
            ScopedStage ss(EPC_STAGE_FC, st);
            if (m->D % 64 == 0) {
                if (int rc = tc_fc_relu(gmax, B, 1024, m->fc1_Wt, m->fc1.b, m->D, o, st)) return rc;
            } else {
                GemmArgs g = {};
                g.A = gmax; g.sAm = 1024; g.sAk = 1;
                g.B = m->fc1.W; g.sBk = m->D; g.sBn = 1;
                g.C = o; g.ldc = m->D; g.M = B; g.N = m->D; g.K = 1024; g.bias = m->fc1.b; g.relu = 1; g.batch = 1; g.splitk = 1;
                if (int rc = sgemm(g, st)) return rc;
            }
            if (int rc = row_l2_normalize(o, B, m->D, out, st)) return rc;
        }
        if (feat)
            for (int b0 = 0; b0 < B; b0 += HEAD_SUB)
                if (int rc = export_feat(b0, (B - b0 < HEAD_SUB) ? (B - b0) : HEAD_SUB)) return rc;
    }
    if (!ar.ok()) {
        set_error("epc_embed: internal workspace accounting error");
        return EPC_EWORKSPACE;
    }
    return EPC_OK;
}

size_t epc_vlad_workspace_bytes(const EpcModel* m, int B, int N) {
    if (!m || !m->vlad_head || B <= 0 || N <= 0) return 256;
    return head_bytes(m, B, N) + 256;
}

int epc_vlad_forward(const EpcModel* m, const float* X, int B, int N, float* out, void* workspace, size_t workspace_bytes,
                     void* stream) {
    EPC_CHECK_ARG(m && X && out, "epc_vlad_forward: NULL argument");
    EPC_CHECK_ARG(m->vlad_head, "epc_vlad_forward: this model (EPC-Net-L) has no VLAD head");
    EPC_CHECK_ARG(B >= 0 && N > 0 && N % 128 == 0, "epc_vlad_forward: max_samples=%d must be a positive multiple of 128", N);
    if (int rc = ensure_device(X)) return rc;
    if (B == 0) return EPC_OK;
    if (!workspace || workspace_bytes < epc_vlad_workspace_bytes(m, B, N)) {
        set_error("epc_vlad_forward: workspace %zu < required %zu", workspace_bytes, epc_vlad_workspace_bytes(m, B, N));
        return EPC_EWORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Arena ar(workspace, workspace_bytes);
    HeadWs h = head_carve(ar, m, B, N);
    const int step = head_step(N);
    for (int b0 = 0; b0 < B; b0 += step) {
        const int nbs = (B - b0 < step) ? (B - b0) : step;
        // the caller's rows are used as given (loupe.py does not normalise): bf16 operands, |row| := 1
        if (head_fp8(N)) {
            if (int rc = f32_to_fp8_rows(X + (size_t)b0 * N * 1024, (long long)nbs * N, 1024, N, reinterpret_cast<uint8_t*>(h.H16), h.rowss, st)) return rc;
        } else if (int rc = f32_to_bf16_rows(X + (size_t)b0 * N * 1024, (long long)nbs * N, 1024, h.H16, h.rowss, st)) return rc;
        if (int rc = head_assign_vlad(m, B, N, b0, nbs, h, 1, st)) return rc;
    }
    return head_tail(m, B, N, h, /*l2=*/0, out, st);
}

// -------------------------------------------------------------------------------------------------
// measurement aid
// -------------------------------------------------------------------------------------------------
}  // extern "C"
namespace epc {
__global__ void __launch_bounds__(256) ffma_peak_kernel(int iters, float seed, float* sink) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i);
    const float m = 1.0f + seed * 1e-9f, c = seed * 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;          // never true: keeps the chains alive
}
}  // namespace epc
extern "C" {
int epc_microbench_ffma(int iters, double* flops, void* stream) {
    EPC_CHECK_ARG(iters >= 1 && flops, "epc_microbench_ffma: bad arguments");
    int dev = 0, sms = 148;
    EPC_CUDA(cudaGetDevice(&dev));
    EPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8;
    epc::ffma_peak_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(iters, 1.0f, nullptr);
    EPC_LAUNCH_CHECK();
    *flops = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
    return EPC_OK;
}

// -------------------------------------------------------------------------------------------------
// retrieval
// -------------------------------------------------------------------------------------------------
size_t epc_retrieve_workspace_bytes(int D, int Q, int dim, int k) { return retrieve_workspace_bytes(D, Q, dim, k); }

int epc_retrieve_topk(const float* db, int D, const float* q, int Q, int dim, int k, long long id_offset, int64_t* idx,
                      double* dist, void* workspace, size_t workspace_bytes, void* stream) {
    EPC_CHECK_ARG((db || D == 0) && (q || Q == 0) && idx && dist && (workspace || D == 0), "epc_retrieve_topk: NULL argument");
    if (int rc = ensure_device(idx)) return rc;
    return retrieve_topk(db, D, q, Q, dim, k, id_offset, nullptr, idx, dist, workspace, workspace_bytes,
                         static_cast<cudaStream_t>(stream));
}

size_t epc_retrieve_index_bytes(int D, int dim) { return retrieve_index_bytes(D, dim); }

int epc_retrieve_index_build(const float* db, int D, int dim, void* index, size_t index_bytes, void* stream) {
    EPC_CHECK_ARG((db || D == 0) && index, "epc_retrieve_index_build: NULL argument");
    if (int rc = ensure_device(index)) return rc;
    return retrieve_index_build(db, D, dim, index, index_bytes, static_cast<cudaStream_t>(stream));
}

int epc_retrieve_topk_indexed(const float* db, int D, const void* index, const float* q, int Q, int dim, int k,
                              long long id_offset, int64_t* idx, double* dist, void* workspace, size_t workspace_bytes,
                              void* stream) {
    EPC_CHECK_ARG((db || D == 0) && (index || D == 0) && (q || Q == 0) && idx && dist && (workspace || D == 0),
                  "epc_retrieve_topk_indexed: NULL argument");
    if (int rc = ensure_device(idx)) return rc;
    return retrieve_topk(db, D, q, Q, dim, k, id_offset, index, idx, dist, workspace, workspace_bytes,
                         static_cast<cudaStream_t>(stream));
}

int epc_radius_count(const double* db, int D, const double* q, int Q, int dim, double r, int32_t* counts, void* stream) {
    EPC_CHECK_ARG(db && (q || Q == 0) && counts, "epc_radius_count: NULL argument");
    if (int rc = ensure_device(db)) return rc;
    return radius_search(db, D, q, Q, dim, r, counts, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int epc_radius_fill(const double* db, int D, const double* q, int Q, int dim, double r, const int64_t* offsets, int32_t* indices,
                    void* stream) {
    EPC_CHECK_ARG(db && (q || Q == 0) && offsets && indices, "epc_radius_fill: NULL argument");
    if (int rc = ensure_device(db)) return rc;
    return radius_search(db, D, q, Q, dim, r, nullptr, offsets, indices, static_cast<cudaStream_t>(stream));
}

int epc_merge_topk(const double* dist, const int64_t* idx, int R, int Q, int k, double* out_dist, int64_t* out_idx,
                   void* stream) {
    EPC_CHECK_ARG(dist && idx && out_dist && out_idx, "epc_merge_topk: NULL argument");
    if (int rc = ensure_device(dist)) return rc;
    return merge_topk(dist, idx, (long long)Q * k, R, Q, k, out_dist, out_idx, static_cast<cudaStream_t>(stream));
}

int epc_merge_topk_strided(const double* dist, const int64_t* idx, long long rank_stride, int R, int Q, int k, double* out_dist,
                           int64_t* out_idx, void* stream) {
    EPC_CHECK_ARG(dist && idx && out_dist && out_idx, "epc_merge_topk_strided: NULL argument");
    if (int rc = ensure_device(dist)) return rc;
    return merge_topk(dist, idx, rank_stride, R, Q, k, out_dist, out_idx, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
