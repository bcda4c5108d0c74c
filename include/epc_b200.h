/* epc_b200.h -- C ABI of the B200-native EPC-Net embedding-and-retrieval hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI: its "plugin" is a
 * Python module name (`MODEL = importlib.import_module(args["ARCH"])`, evaluate.py:119) exposing
 * `forward(point_cloud, is_training, bn_decay, params)` (models/epc-net.py:29), below which sit the
 * operator wrappers of utils/tf_util.py and loupe.py.  Each entry point here replaces the TF sub-graph
 * that the cited reference function builds; the Python shims in `epc-net_b200/` bind them with ctypes
 * (INTEGRATION.md shows the binding a reference maintainer would add).
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer on the current CUDA device unless the
 *     parameter name ends in `_host`;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = default stream); calls are asynchronous
 *     with respect to the host unless stated otherwise;
 *   - return value: 0 on success, negative `EPC_E*` code otherwise; `epc_last_error()` returns a
 *     thread-local message.  Nothing throws; nothing allocates device memory except
 *     `epc_model_create` (weights) and the API-parity helper `epc_dense_forward` (see there) --
 *     scratch comes from the caller (`*_workspace_bytes`);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns EPC_ECUDA.
 *   - tensors are dense row-major fp32 unless noted.  N (points per cloud) must be a multiple of 32,
 *     32 <= N <= 8192 (the reference fixes N = 4096: utils/loading_pointclouds.py:32).
 */
#ifndef EPC_B200_H_
#define EPC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPC_ABI_VERSION 1

enum {
    EPC_OK = 0,
    EPC_EINVAL = -1,   /* bad argument (shape, NULL pointer, unsupported size) */
    EPC_ECUDA = -2,    /* CUDA runtime error / no device */
    EPC_EWORKSPACE = -3, /* workspace too small */
    EPC_EUNSUPPORTED = -4
};

/* ARCH values: the module names accepted by importlib.import_module(args["ARCH"]) (evaluate.py:119). */
enum {
    EPC_ARCH_EPC_NET = 0,      /* models/epc-net.py      4 ProxyConv blocks + G_VLAD         */
    EPC_ARCH_EPC_NET_L = 1,    /* models/epc-net-l.py    2 blocks + global max-pool + FC     */
    EPC_ARCH_KD_EPC_NET = 2,   /* models/kd_epc-net.py   = EPC_NET, also returns per-point features */
    EPC_ARCH_KD_EPC_NET_L = 3  /* models/kd_epc-net-l.py = EPC_NET_L (scope BACKBONE), ditto */
};

/* fp32 evaluation order of p_i . p_j inside pairwise_distance_mask (utils/tf_util.py:651-656).
 * TensorFlow is not runnable in the build environment, so this is a documented reconstruction
 * (DESIGN.md "kNN arithmetic"): MULADD = ((x x' + y y') + z z') separately rounded (TF-1.12 CPU
 * wheel: AVX, no FMA); FMA = fma(z,z',fma(y,y',x x')) (cuBLAS sgemm, i.e. the TF GPU path). */
enum { EPC_KNN_ARITH_MULADD = 0, EPC_KNN_ARITH_FMA = 1 };

/* Pooling head of loupe.py */
enum { EPC_POOL_G_VLAD = 0 /* loupe.py:216-333 */, EPC_POOL_NETVLAD = 1 /* loupe.py:103-214 */ };

const char* epc_last_error(void);
int epc_abi_version(void);
/* number of kernel launches issued by this library on the calling thread since the last reset */
long long epc_launch_count(void);
void epc_launch_count_reset(void);
/* Device selection for calls that take no device pointer (epc_model_create).  Every other entry point
 * runs on the device that owns its first device-pointer argument. */
int epc_set_device(int device);
int epc_device_count(void);   /* 0 when no CUDA device/driver is present */

/* Per-stage device timing: when enabled, every stage's kernels are bracketed by CUDA events on the launching
 * stream; epc_profile_read synchronises those events and returns the accumulated milliseconds and the number of
 * bracketed launches since the last reset.  Used by bench.py for the roofline of the dominant kernel. */
enum {
    EPC_STAGE_SORT = 0, EPC_STAGE_KNN, EPC_STAGE_CONV_IN, EPC_STAGE_BLOCK, EPC_STAGE_CONV5, EPC_STAGE_ROWNORM,
    EPC_STAGE_ASSIGN_GEMM, EPC_STAGE_ASSIGN_SOFTMAX, EPC_STAGE_VLAD_GEMM, EPC_STAGE_VLAD_FINALIZE,
    EPC_STAGE_HIDDEN_GEMM, EPC_STAGE_TAIL, EPC_STAGE_COLMAX, EPC_STAGE_FC, EPC_STAGE_KD_FEAT,
    EPC_STAGE_RETRIEVE_SCORE, EPC_STAGE_RETRIEVE_SELECT, EPC_STAGE_RETRIEVE_RERANK, EPC_STAGE_BLOCK_SAFE, EPC_STAGE_ASSIGN_VLAD,
    EPC_STAGE_COUNT
};
void epc_profile_enable(int on);
void epc_profile_reset(void);
int epc_profile_read(int stage, double* ms, long long* launches);
const char* epc_stage_name(int stage);
/* Measurement aid for bench.py's roofline (not on the data path): launches a register-only FFMA loop on every SM,
 * `iters` x 16 independent fused multiply-adds per thread, and returns the FLOP count of the launch (2 per FMA) in
 * *flops.  The caller times it with events on `stream`: FLOP / time = the FP32 pipe's achievable peak at the clocks
 * of the moment, which is the roofline of the ALU-bound kNN kernels. */
int epc_microbench_ffma(int iters, double* flops, void* stream);

/* ---------------------------------------------------------------------------------------------
 * kNN graph  (replaces tf_util.pairwise_distance_mask, utils/tf_util.py:647-666; and
 *             tf_util.pairwise_distance / tf_util.knn, utils/tf_util.py:577-610)
 *
 * a_ij = -((s_i + (-2 p_i.p_j)) + s_j); the reference's neighbour set of row i is
 * {j : a_ij >= kth_i}, kth_i = 20th largest a_ij (literal 20, :660).  Ties at the 20th value enlarge
 * the set (count > 20) while the divisor of the mean stays 20 (models/epc-net.py:71).
 *
 * epc_knn outputs, all in ORIGINAL point order:
 *   idx   [B,N,20] int32  the 20 best j in tf.nn.top_k order (descending a; ties -> lower j first)
 *   kth   [B,N]    fp32   kth_i (a value, i.e. minus the distance expression); bit-exact
 *   count [B,N]    int32  |{j : a_ij >= kth_i}|  (>= 20)
 * any of the three may be NULL.  workspace: epc_knn_workspace_bytes(B,N).
 * ------------------------------------------------------------------------------------------- */
size_t epc_knn_workspace_bytes(int B, int N);
int epc_knn(const float* xyz /*[B,N,3]*/, int B, int N, int arith,
            int32_t* idx, float* kth, int32_t* count,
            void* workspace, size_t workspace_bytes, void* stream);

/* Test hook: epc_knn with the exact AABB block pruning disabled (must be bit-identical to epc_knn). */
int epc_knn_noprune(const float* xyz, int B, int N, int arith, int32_t* idx, float* kth, int32_t* count,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Dense outputs for API parity/testing: mask [B,N,N] fp32 0/1 (pairwise_distance_mask's return
 * value) and/or dist [B,N,N] fp32 (pairwise_distance's return value, = -a).  Either may be NULL. */
int epc_knn_dense(const float* xyz /*[B,N,3]*/, int B, int N, int arith,
                  float* mask, float* dist,
                  void* workspace, size_t workspace_bytes, void* stream);

/* tf_util.knn(adj_matrix, k) (utils/tf_util.py:599-610): indices of the k smallest entries of each
 * row of adj [R,M], ascending value, ties -> lower index first.  k <= 32. */
int epc_rows_topk_smallest(const float* adj /*[R,M]*/, long long R, int M, int k,
                           int32_t* idx /*[R,k]*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Layer wrappers (utils/tf_util.py)
 * ------------------------------------------------------------------------------------------- */
/* inference batch norm parameters of one layer (tf_util.batch_norm_template :454-491, or
 * slim/contrib batch_norm); all HOST pointers to [C] fp32; eps is 1e-3 everywhere in the reference */
typedef struct EpcBN {
    const float* beta_host;
    const float* gamma_host;
    const float* mean_host;
    const float* var_host;
} EpcBN;

/* tf_util.conv1d(kernel 1, bn=True, relu) (:52-107) and tf_util.fully_connected(bn=True, relu)
 * (:310-346): y = relu(BN(x W + b)).  W [cin,cout], b [cout] (HOST pointers). */
typedef struct EpcDense {
    const float* weights_host;
    const float* biases_host;
    EpcBN bn;
    int cin, cout;
} EpcDense;

/* Stand-alone pointwise layer on device data: y[R,cout] = act(BN(x[R,cin] W + b)); relu != 0 applies
 * the ReLU.  Weights are folded, uploaded into a temporary device buffer (cudaMalloc/cudaFree) and the
 * call SYNCHRONISES the stream before returning -- this entry point exists for API parity of
 * tf_util.conv1d / fully_connected and for tests, not for the fused path (epc_embed never calls it). */
int epc_dense_forward(const EpcDense* layer, const float* x, long long R, float* y, int relu, void* stream);

/* tf_util.max_pool2d(x[B,N,1,C], [N,1]) (:349-372, models/epc-net-l.py:91): y[B,C] = max_n x[B,n,C] */
int epc_max_pool_points(const float* x /*[B,N,C]*/, int B, int N, int C, float* y /*[B,C]*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Model  (replaces MODEL.forward: models/epc-net.py:29-157, models/epc-net-l.py:29-102,
 *         models/kd_epc-net.py:157-158, models/kd_epc-net-l.py:44,102)
 * ------------------------------------------------------------------------------------------- */
typedef struct EpcWeights {
    int arch;                 /* EPC_ARCH_*                                                     */
    int knn_k;                /* params["KNN"]: ONLY the divisor of the neighbour mean (20)     */
    int cluster_size;         /* params["CLUSTER_SIZE"]: 64 only (EPC_EUNSUPPORTED otherwise) -- */
                              /* deliberate: every shipped config uses 64 and the assignment      */
                              /* epilogue keeps a point's 64 logits in one thread's registers     */
    int output_dim;           /* params["FEATURE_OUTPUT_DIM"] (256): any multiple of 64 <= 1024  */
    int groups;               /* params["GROUPS"] (4); ignored for NETVLAD / -L                 */
    int pooling;              /* EPC_POOL_*                                                     */
    int gating;               /* loupe gating flag (1)                                          */
    int n_blocks;             /* 4 (EPC-Net) or 2 (EPC-Net-L)                                   */
    EpcDense conv[12];        /* conv1,conv1_a,conv1_b,conv2,... (3 per block)                  */
    EpcDense conv5;           /* [64*n_blocks -> 1024]                                          */
    /* G_VLAD / NetVLAD head (loupe.py:249-331) */
    const float* cluster_weights_host;   /* [1024, K]                                           */
    EpcBN cluster_bn;                    /* [K]                                                 */
    const float* cluster_weights2_host;  /* [1, 1024, K]                                        */
    const float* hidden1_weights_host;   /* [1024*K/groups, D]  (NetVLAD: [1024*K, D])          */
    EpcBN hidden_bn;                     /* 'bn' [D]                                            */
    const float* gating_weights_host;    /* [D, D]                                              */
    EpcBN gating_bn;                     /* [D]                                                 */
    /* EPC-Net-L head (models/epc-net-l.py:95) */
    EpcDense fc1;                        /* [1024 -> D]                                         */
} EpcWeights;

typedef struct EpcModel EpcModel;   /* opaque, immutable after creation => usable from many streams */

int epc_model_create(const EpcWeights* w, EpcModel** out);
void epc_model_destroy(EpcModel* m);

size_t epc_embed_workspace_bytes(const EpcModel* m, int B, int N);
/* Embeds B clouds: xyz [B,N,3] -> out [B,D] (L2-normalised descriptors, the `last_output` tensor of
 * models/epc-net.py:155 flattened over (Bq,P)).  If feat != NULL it receives the KD feature
 * l2norm(conv5 per-point features) [B*N,1024] (models/kd_epc-net.py:158) in ORIGINAL point order. */
int epc_embed(const EpcModel* m, const float* xyz, int B, int N, int knn_arith,
              float* out, float* feat, void* workspace, size_t workspace_bytes, void* stream);

/* loupe.G_VLAD(...).forward(reshaped_input) / loupe.NetVLAD(...).forward (loupe.py:233-333,119-214)
 * on caller-provided per-point features X [B*max_samples, 1024] (rows need not be normalised --
 * exactly like the reference, the caller does that): out [B,D] (NOT L2-normalised, as in loupe). */
size_t epc_vlad_workspace_bytes(const EpcModel* m, int B, int N);
int epc_vlad_forward(const EpcModel* m, const float* X, int B, int N, float* out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Retrieval  (replaces KDTree(database).query(q, k) of evaluate.get_recall, evaluate.py:463,481)
 *
 * Exact Euclidean k-NN of each query among `D` database rows: fp32 tensor/FFMA scoring to pick
 * candidates, float64 re-ranking so that the result equals a float64 brute force (what the KD-tree
 * returns) -- ascending distance, ties -> lower index first.
 *   idx  [Q,k] int64 (database row + id_offset: lets a shard report global row ids)
 *   dist [Q,k] float64 Euclidean distances
 * k <= 32 takes the tensor-core path; k > 32 (any k, as KDTree.query allows) runs exact float64 scans of the whole
 * database in passes of 32 neighbours -- correct, but meant for the occasional small call (train.py:857-869).
 * ------------------------------------------------------------------------------------------- */
size_t epc_retrieve_workspace_bytes(int D, int Q, int dim, int k);
int epc_retrieve_topk(const float* db /*[D,dim]*/, int D, const float* q /*[Q,dim]*/, int Q, int dim, int k,
                      long long id_offset, int64_t* idx, double* dist,
                      void* workspace, size_t workspace_bytes, void* stream);
/* The prepared database -- what the `database_nbrs = KDTree(database_output)` object of evaluate.py:463 is to the
 * reference: built once per database set, queried by every query set of the m != n pair loop (evaluate.py:291-300).
 * It holds |d|^2 per row and the bf16 (hi | lo) operand pairs of the scoring GEMM in caller-owned device memory of
 * epc_retrieve_index_bytes(D, dim) bytes; `db` itself must stay alive (the float64 re-rank reads it).
 * epc_retrieve_topk_indexed == epc_retrieve_topk minus the per-call database preparation.  D == 0 is legal
 * (an empty shard): every idx is -1 and every dist +inf. */
size_t epc_retrieve_index_bytes(int D, int dim);
int epc_retrieve_index_build(const float* db /*[D,dim]*/, int D, int dim, void* index, size_t index_bytes, void* stream);
int epc_retrieve_topk_indexed(const float* db /*[D,dim]*/, int D, const void* index, const float* q /*[Q,dim]*/, int Q,
                              int dim, int k, long long id_offset, int64_t* idx, double* dist,
                              void* workspace, size_t workspace_bytes, void* stream);
/* Merge R per-shard candidate lists (e.g. after an NCCL all-gather): dist/idx [R,Q,k] -> [Q,k],
 * ordered by (distance, index) so the result does not depend on the shard count. */
int epc_merge_topk(const double* dist /*[R,Q,k]*/, const int64_t* idx /*[R,Q,k]*/, int R, int Q, int k,
                   double* out_dist /*[Q,k]*/, int64_t* out_idx /*[Q,k]*/, void* stream);
/* The same merge on lists whose rank-r slab starts rank_stride elements after rank r-1's (>= Q*k): lets every rank send
 * ONE packed (dist | idx) buffer through a single all-gather (dist = buf, idx = buf + Q*k, rank_stride = 2*Q*k). */
int epc_merge_topk_strided(const double* dist, const int64_t* idx, long long rank_stride, int R, int Q, int k,
                           double* out_dist /*[Q,k]*/, int64_t* out_idx /*[Q,k]*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Radius search  (replaces KDTree(db[['northing','easting']]).query_radius(coor, r) of
 * generating_queries/generate_test_sets.py:70-104 and generate_training_tuples_baseline.py:52-62, which
 * build the true_neighbors lists consumed by evaluate.get_recall)
 *
 * All rows j of db [D,dim] (float64) with sum_k (q_k - d_k)^2 <= r^2, per query, as a CSR pair: first
 * epc_radius_count -> counts [Q]; the caller turns them into exclusive offsets [Q] (int64) and sizes
 * `indices`; then epc_radius_fill writes each query's row ids in ascending order.
 * ------------------------------------------------------------------------------------------- */
int epc_radius_count(const double* db /*[D,dim]*/, int D, const double* q /*[Q,dim]*/, int Q, int dim, double r,
                     int32_t* counts /*[Q]*/, void* stream);
int epc_radius_fill(const double* db, int D, const double* q, int Q, int dim, double r, const int64_t* offsets /*[Q]*/,
                    int32_t* indices, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EPC_B200_H_ */
