/*
 * graphtrans_b200 — C ABI of the B200 (sm_100a) GraphTrans forward/backward hot path.
 *
 * The reference (ucbrise/graphtrans) is pure Python: it has no FFI of its own.  Every entry
 * point below replaces the third-party op that the cited reference call site reaches
 * (SURVEY.md §2.2 / §8a); the Python host mirror in graphtrans_b200/{models,modules} binds them
 * with ctypes (graphtrans_b200/_lib.py) exactly as INTEGRATION.md shows.
 *
 * Conventions
 *  - every export returns int: 0 ok, <0 argument/shape/alignment error, >0 cudaError_t;
 *    gt_last_error() gives the thread-local message.  Nothing throws, exits or prints.
 *  - all pointers are DEVICE pointers owned by the caller (PyTorch's allocator); the library
 *    borrows them for the duration of the enqueue and keeps no global mutable state.
 *  - everything is asynchronous on `stream` (a cudaStream_t), performs no allocation and no
 *    synchronisation, and is CUDA-graph capturable.
 *  - feature matrices are row-major [rows, ld] with logical width d <= ld; ld % 4 == 0 and
 *    16-byte aligned bases.  Columns d..ld-1 are kept zero by every kernel.
 *  - `dt` is the activation dtype (GT_F32 / GT_BF16); parameters, statistics and parameter
 *    gradients are always fp32 (statistics accumulators fp64).
 */
#ifndef GRAPHTRANS_B200_H
#define GRAPHTRANS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { GT_F32 = 0, GT_BF16 = 1 };
/* edge-encoder kinds (reference dataset/tud.py:67-71, dataset/code.py:117, dataset/mol.py:84) */
enum { GT_EDGE_NONE = 0, GT_EDGE_LINEAR = 1, GT_EDGE_TABLE = 2 };
/* aggregation kinds */
enum { GT_CONV_GCN = 0, GT_CONV_GIN = 1 };
/* GEMM epilogue flags */
enum { GT_EPI_RELU = 1, GT_EPI_ACCUM = 2, GT_EPI_OUT_F32 = 4, GT_EPI_RESID_F32 = 8 };

int gt_version(void);
const char* gt_last_error(void);

/* ---- graph preparation (integer; replaces PyG MessagePassing's index handling and
 *      torch_geometric.utils.degree, reference modules/conv.py:28,57,63) --------------------
 * edge_index: int64 [2,E] (row 0 = source, row 1 = target).  Builds, with a stable counting
 * sort, the target-sorted CSR (rowptr_dst[N+1], src_by_dst[E], eid_by_dst[E]) used by the
 * forward aggregation and the source-sorted CSR (rowptr_src, dst_by_src, eid_by_src) used by
 * its adjoint.  out-degree(i) = rowptr_src[i+1]-rowptr_src[i]; in-degree likewise.
 * work: int32 [2*(N+1)] scratch. */
int gt_csr_build(const int64_t* edge_index, int64_t E, int64_t N,
                 int32_t* rowptr_dst, int32_t* src_by_dst, int32_t* eid_by_dst,
                 int32_t* rowptr_src, int32_t* dst_by_src, int32_t* eid_by_src,
                 int32_t* work, void* stream);

/* combined edge-type id for table edge encoders: etype[e] = sum_c attr[e,c] * mult[c]
 * (ogb BondEncoder, reference dataset/mol.py:84: 5*6*2 = 60 combinations). */
int gt_edge_type(const int64_t* edge_attr, int64_t E, int32_t ncol, const int32_t* mult_host,
                 int32_t* etype, void* stream);

/* ---- dropout RNG state -------------------------------------------------------------------
 * rng_state: DEVICE uint64[2] = {seed, step}.  Every kernel that takes (drop_p, rng_state, salt)
 * derives its keep mask from (seed, step, salt, element index) with a counter-based hash, so
 * forward and backward agree by construction and CUDA-graph replays see fresh masks once
 * gt_rng_advance (step += 1) is part of the graph.  drop_p == 0 or rng_state == NULL: no dropout.
 * Replaces nn.Dropout / F.dropout (reference modules/gnn_module.py:86-90,205-209,227-229;
 * the four dropout sites of nn.TransformerEncoderLayer, modules/transformer_encoder.py:28-32). */
int gt_rng_advance(uint64_t* rng_state, void* stream);
/* y = x * keep / (1 - p), n % 4 == 0 elements; the same call on dy is the backward */
int gt_dropout(int dt, const void* x, int64_t n, void* y, float drop_p, const uint64_t* rng_state,
               uint64_t salt, void* stream);

/* ---- batch plan (integer, bit-exact; closed form of reference modules/utils.py:5-29) ------
 * batch: int64 [N] sorted graph ids, B graphs, L = max_input_len, cls = 1 when a <CLS> row is
 * appended per graph (graph_pooling == "cls", reference modules/transformer_encoder.py:50-55).
 * node_off[B+1]; kept[B] = min(n_i, L); tok_off[B+1] = exclusive scan of kept+cls (packed
 * token layout: graph i owns token rows [tok_off[i], tok_off[i+1]), its last row is <CLS> or,
 * with cls = 0, its last node);
 * tok2node[N+B]: node id, -1 for <CLS>, -2 for unused tail rows; tok_graph[N+B]: graph id or
 * -1; node_graph[N]: int32 copy of batch; node2tok[N]: token row of a node or -1 when truncated
 * away; cls_rows[B]: token row of each graph's pooled position (<CLS>, or last node if cls = 0);
 * scalars[4] = {S = min(max n_i, L), n_tok, max n_i, 0}. */
int gt_batch_plan(const int64_t* batch, int64_t N, int64_t B, int64_t L, int32_t cls,
                  int32_t* node_off, int32_t* kept, int32_t* tok_off, int32_t* tok2node,
                  int32_t* tok_graph, int32_t* node_graph, int32_t* node2tok, int32_t* cls_rows,
                  int32_t* scalars, void* stream);

/* ---- node encoders (reference dataset/utils.py:28-30 ASTNodeEncoder, ogb AtomEncoder) ------
 * out[i,:] = sum_c table_c[min(idx_c[i*stride_c], clamp_c), :].  ncol <= 12.  All arrays of
 * per-column descriptors are HOST arrays. */
int gt_embed_sum_fwd(int dt, void* out, int64_t N, int32_t d, int32_t ld, int32_t ncol,
                     const int64_t* const* idx_host, const int64_t* stride_host,
                     const int64_t* clamp_host, const float* const* table_host, void* stream);
int gt_embed_sum_bwd(int dt, const void* dout, int64_t N, int32_t d, int32_t ld, int32_t ncol,
                     const int64_t* const* idx_host, const int64_t* stride_host,
                     const int64_t* clamp_host, float* const* dtable_host, void* stream);

/* Embedding gradients of small tables as a contraction (backward of dataset/utils.py:28-30 / ogb AtomEncoder):
 * gt_onehot writes the bf16 one-hot operand [N, r_pad] (1 at base_c + min(idx_c[i], clamp_c) for every column with
 * base_c >= 0); d_tables = OneHot^T . dout is then one gt_gemm (a_mn = b_mn = 1, GT_EPI_ACCUM | GT_EPI_OUT_F32) into
 * an fp32 [R, ld_t] buffer that gt_embed_unpack adds to the per-table gradients (rows [base_c, base_c + rows_c)). */
int gt_onehot(int64_t N, int32_t ncol, const int64_t* const* idx_host, const int64_t* stride_host,
              const int64_t* clamp_host, const int32_t* base_host, int32_t r_pad, void* out, void* stream);
int gt_embed_unpack(const float* temp, int32_t ld_t, int32_t d, int32_t R, int32_t ncol, const int32_t* base_host,
                    const int32_t* rows_host, float* const* dtable_host, void* stream);

/* ---- stage 1: message-passing aggregation (reference modules/conv.py:26-33, 50-68) ---------
 * GCN: out[i] = sum_{e:(j->i)} rsqrt(deg_j) rsqrt(deg_i) relu(x[j]+ee_e) + relu(x[i]+root)/deg_i,
 *      deg = out-degree + 1 (conv.py:57);   GIN: out[i] = (1+eps) x[i] + sum relu(x[j]+ee_e).
 * Edge embedding ee_e is recomputed in-kernel: LINEAR: ee = b + W[:, :kdim] attr_e (W is the
 * nn.Linear weight [d,kdim], attr fp32 [E,kdim], kdim <= 4); TABLE: ee = table[etype_e] (fp32
 * [ntypes, ld]); NONE: 0.  `self_param`: root_emb [d] (GCN) or eps [1] (GIN), fp32. */
int gt_aggregate_fwd(int dt, int conv, const void* x, void* out, int64_t N, int32_t d, int32_t ld,
                     const int32_t* rowptr_dst, const int32_t* src_by_dst, const int32_t* eid_by_dst,
                     const int32_t* rowptr_src,
                     int edge_kind, const float* edge_attr, int32_t kdim, const float* edge_w,
                     const float* edge_b, const int32_t* etype, const float* table,
                     const float* self_param, const float* norm_slot, const int32_t* etype_slot,
                     const float* attr_slot, void* stream);
/* adjoint: dx (same dtype), and fp32 accumulators (must be zeroed by the caller):
 * d_edge_w [d,kdim], d_edge_b [d] (LINEAR) or d_table [ntypes, ld] (TABLE; NULL = skipped here, see
 * gt_aggregate_table_grad), d_self ([d] or [1]).  th_scratch: optional bf16 [ntypes + 1, ld] work buffer (GIN, TABLE,
 * bf16, d_table == NULL): the call first fills it with the bf16 ReLU thresholds of this layer's table and then runs the
 * packed-mask adjoint (same results bit for bit: x + e > 0 <=> x > round-down-to-bf16(-e) for a bf16 x); NULL = fp32 mask. */
int gt_aggregate_bwd(int dt, int conv, const void* x, const void* dout, void* dx, int64_t N,
                     int32_t d, int32_t ld,
                     const int32_t* rowptr_dst, const int32_t* rowptr_src, const int32_t* dst_by_src,
                     const int32_t* eid_by_src,
                     int edge_kind, const float* edge_attr, int32_t kdim, const float* edge_w,
                     const float* edge_b, const int32_t* etype, const float* table, int32_t ntypes,
                     const float* self_param,
                     float* d_edge_w, float* d_edge_b, float* d_table, float* d_self, const float* norm_slot,
                     const int32_t* etype_slot, const float* attr_slot, void* gm_out, void* th_scratch, void* stream);
/* gm_out (optional, activation dtype [E, ld], source-sorted slot order = dst_by_src order): per-edge masked message
 * gradient norm_e * dout[dst_e] * 1[x[src_e] + ee_e > 0].  With it the edge-TABLE gradient is the contraction
 * d_table = OneHot(etype_slot)^T . gm (gt_onehot + gt_gemm) instead of gt_aggregate_table_grad's second gather pass. */
/* per-CSR-slot copies of the per-edge data, computed once per batch and reused by every layer (optional; NULL
 * arguments to gt_aggregate_* mean "follow the indirection in-kernel").  For the CSR given by (rowptr_slot, nbr_slot,
 * eid_slot) - target-sorted for the forward, source-sorted for the adjoint: norm_slot[p] = GCN norm of the edge in
 * slot p (deg = out-degree + 1 from rowptr_src), etype_slot[p] = etype[eid], attr_slot[p,:] = edge_attr[eid,:]. */
int gt_edge_slots(const int32_t* rowptr_slot, const int32_t* nbr_slot, const int32_t* eid_slot,
                  const int32_t* rowptr_src, int64_t E, int64_t N, const int32_t* etype,
                  const float* edge_attr, int32_t kdim, float* norm_slot, int32_t* etype_slot,
                  float* attr_slot, void* stream);

/* Edge-table gradient of gt_aggregate split off the adjoint (leaf gradient, off the critical path):
 * d_table[ty,:] += sum over edges e of type ty of norm_e * dout[dst_e,:] * 1[x[src_e,:] + table[ty,:] > 0]
 * (norm_e = 1 for GIN).  Pass d_table = NULL to gt_aggregate_bwd when this is used.  The edges come SORTED BY TYPE
 * from gt_edges_by_type (once per batch): type_ptr[ntypes+1] (run boundaries), src_t / dst_t / type_t [E] = endpoints
 * and type of the edge in sorted slot p (order inside a type run unspecified); work = int32[2*ntypes] scratch;
 * every etype must lie in [0, ntypes), ntypes <= 1024. */
int gt_edges_by_type(const int64_t* edge_index, const int32_t* etype, int64_t E, int32_t ntypes,
                     int32_t* type_ptr, int32_t* src_t, int32_t* dst_t, int32_t* type_t, int32_t* work,
                     void* stream);
int gt_aggregate_table_grad(int dt, int conv, const void* x, const void* dout, int64_t N, int32_t d, int32_t ld,
                            const int32_t* rowptr_src, int64_t E, const int32_t* src_t, const int32_t* dst_t,
                            const int32_t* type_t, const float* table, int32_t ntypes, float* d_table,
                            void* stream);

/* ---- per-graph segment ops (PyG global_add_pool / vn[batch], reference
 *      modules/gnn_module.py:199,219) --------------------------------------------------------
 * segment_sum: out[g,:] (fp32 [B,ld], pre-zeroed unless init given) += sum_{i in g} x[i,:] */
int gt_segment_sum(int dt, const void* x, const int32_t* node_graph, int64_t N, int32_t ld,
                   float* out, void* stream);
/* same for graphs that own contiguous row ranges [node_off[g], node_off[g+1]) (PyG batches are sorted): one warp per
 * (graph, 128-channel chunk), single writer, no atomics (deterministic);
 * out[g,:] = (init ? init[g,:] : 0) + sum_{i in g} x[i,:]  (fp32 [B, ld]; out needs no zero fill, init may be NULL) */
int gt_segment_sum_sorted(int dt, const void* x, const int32_t* node_off, int64_t B, int32_t ld,
                          const float* init, float* out, void* stream);
/* y[i,:] = x[i,:] + v[node_graph[i],:]   (v fp32 [B,ld]); x may be NULL (pure broadcast) */
int gt_add_graph_vec(int dt, const void* x, const float* v, const int32_t* node_graph, int64_t N,
                     int32_t ld, void* y, void* stream);

/* ---- BatchNorm1d, train/eval (reference modules/gnn_module.py:58,84,164,167; conv.py:19) ---
 * colstats: stats[0:ld] += sum_rows x, stats[ld:2ld] += sum_rows x^2 (fp64, pre-zeroed).
 * m_valid (optional DEVICE int32[1], all BatchNorm entry points): only the leading m_valid[0] rows are real; the rest is
 * shape-bucket slack (batches padded up to a bucket so that CUDA-graph signatures repeat, graphtrans_b200/graphed.py):
 * slack rows stay out of the batch statistics (divisor m_valid[0]) and receive a zero gradient.  NULL = all M rows. */
int gt_colstats(int dt, const void* x, int64_t M, int32_t ld, double* stats, const int32_t* m_valid, void* stream);
/* finalize: from stats (train) or running stats (eval) produce scale/shift (y = x*scale+shift)
 * and mean/rstd; in train mode update running_mean/var (momentum, unbiased var) and ++nbt. */
int gt_bn_finalize(const double* stats, int64_t M, int32_t d, int32_t ld, const float* gamma,
                   const float* beta, float* running_mean, float* running_var, int64_t* nbt,
                   float momentum, float eps, int training, float* scale_shift_mean_rstd,
                   void* stream);
/* y = drop(act(x*scale+shift)) [+ resid] [+ gvec[node_graph]] ; act = relu if relu!=0 */
int gt_bn_apply_fwd(int dt, const void* x, int64_t M, int32_t d, int32_t ld, const float* ssmr,
                    int relu, const void* resid, const float* gvec, const int32_t* node_graph,
                    void* y, float drop_p, const uint64_t* rng_state, uint64_t salt, void* stream);
/* gt_bn_finalize + gt_bn_apply_fwd in one launch (the path the modules use): scale/shift are derived from `stats`
 * (train, from gt_colstats) or the running statistics (eval) inside the kernel; ssmr [4*ld] is written for the
 * backward, running statistics / num_batches_tracked are updated in train mode. */
int gt_bn_norm_fwd(int dt, const void* x, int64_t M, int32_t d, int32_t ld, const double* stats,
                   const float* gamma, const float* beta, float* running_mean, float* running_var,
                   int64_t* nbt, float momentum, float eps, int training, int relu, const void* resid,
                   const float* gvec, const int32_t* node_graph, void* y, float* ssmr, float drop_p,
                   const uint64_t* rng_state, uint64_t salt, const int32_t* m_valid, void* stream);
/* backward pass 1: g = dy * keep/(1-p) * relu'(x*scale+shift); red[0:ld] += sum g,
 * red[ld:2ld] += sum g*xhat */
int gt_bn_bwd_reduce(int dt, const void* x, const void* dy, int64_t M, int32_t d, int32_t ld,
                     const float* ssmr, int relu, double* red, float drop_p, const uint64_t* rng_state,
                     uint64_t salt, const int32_t* m_valid, void* stream);
/* backward pass 2: dx = gamma*rstd*(g - red0/M - xhat*red1/M) (train) or g*scale (eval);
 * dgamma += red1, dbeta += red0 (fp32 [d], accumulated so that gradients can be summed in place) */
int gt_bn_bwd_apply(int dt, const void* x, const void* dy, int64_t M, int32_t d, int32_t ld,
                    const float* ssmr, const float* gamma, int relu, int training, const double* red,
                    void* dx, float* dgamma, float* dbeta, float drop_p, const uint64_t* rng_state,
                    uint64_t salt, const int32_t* m_valid, void* stream);

/* ---- dense contraction (replaces nn.Linear -> cuBLAS, reference modules/conv.py:18-20,44;
 *      modules/gnn_module.py:161-170; models/gnn_transformer.py:70,85-88; the in/out
 *      projections and FFN inside nn.TransformerEncoderLayer) --------------------------------
 * C[m,n] = sum_k A(m,k) B(n,k) (+ bias[n]) (+ resid[m,n]) (relu) ; columns N..n_fill-1 := 0.
 * a_mn / b_mn: operand stored "MN-major" (element (m,k) at A[k*lda+m]) instead of K-major
 * (A[m*lda+k]).  A/B dtype = dt; C dtype = dt unless GT_EPI_OUT_F32; GT_EPI_ACCUM adds into C
 * (fp32 C only; used with splits > 1).  drop_p/rng_state/salt: dropout applied after the activation
 * (drop(relu(x W^T + b)), the FFN of nn.TransformerEncoderLayer); needs ldc % 4 == 0.  impl: 0 = auto (tcgen05 when eligible), 1 = CUDA-core
 * reference kernel, 2 = tcgen05 only (error when not eligible). */
int gt_gemm(int dt, const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb,
            void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int64_t n_fill,
            const float* bias, const void* resid, int64_t ldr, int flags, float drop_p,
            const uint64_t* rng_state, uint64_t salt, int impl, void* stream);
/* gt_gemm + the BatchNorm column statistics of the stored C in one pass: col_stats fp64 [2][ldc] (pre-zeroed) receives
 * += sum_m C[m, c] and += sum_m C[m, c]^2 exactly as gt_colstats(C) would (of the ROUNDED stored values).  In the tcgen05
 * kernel the sums are taken from the shared-memory slab of the TMA-store epilogue; other paths run gt_colstats after
 * the contraction.  Needs a non-accumulating C with n_fill >= ldc (every column written). */
int gt_gemm_stats(int dt, const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb,
                  void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int64_t n_fill,
                  const float* bias, const void* resid, int64_t ldr, int flags, float drop_p,
                  const uint64_t* rng_state, uint64_t salt, int impl, double* col_stats, const int32_t* m_valid,
                  void* stream);
/* dz = dy * (y > 0) * scale: backward of a ReLU (+ dropout: a dropped element has y == 0, scale = 1/(1-p)) that
 * was fused into a GEMM epilogue; n % 4 == 0 */
int gt_relu_bwd(int dt, const void* dy, const void* y, int64_t n, void* dz, float scale, void* stream);
/* the same over a [M, ld] matrix together with the column sums of dz: colsum fp32 [N] += sum_m dz[m, 0:N] (the bias
 * gradient of the Linear whose epilogue applied the ReLU, reference nn.TransformerEncoderLayer linear1 built at
 * modules/transformer_encoder.py:28-32); rows 16-byte aligned */
int gt_relu_bwd_colsum(int dt, const void* dy, const void* y, int64_t M, int64_t N, int64_t ld, void* dz, float scale,
                       float* colsum, void* stream);
/* fp32 parity mode ON the tensor cores: dst bf16 [3][rows][ld_dst] = the three-term split x = p0 + p1 + p2 of src fp32
 * [rows, cols] (row pitch ld_src; columns cols..ld_dst-1 := 0).  The host sums the six products p_i . q_j with i + j <= 2
 * through gt_gemm (fp32 accumulation in TMEM), which reproduces an fp32 contraction to ~2^-22 (ops._gemm_raw, GT_GEMM_TC_PARITY). */
int gt_split3(const float* src, int64_t rows, int64_t cols, int64_t ld_src, void* dst, int64_t ld_dst, void* stream);
/* out[n] += sum_m X[m,n]  (bias gradients) ; out fp32 [N] is ACCUMULATED into (caller zeroes a fresh buffer) */
int gt_colsum(int dt, const void* X, int64_t M, int64_t N, int64_t ld, float* out, void* stream);
/* dst[r, 0:cols_out] = cast(src[r, 0:cols_in]) zero padded to cols_out; rows_out >= rows_in zero padded;
 * ld_in == 0 broadcasts the single source row to rows_in rows */
int gt_cast_pad(int dt_in, const void* src, int64_t rows_in, int64_t cols_in, int64_t ld_in,
                int dt_out, void* dst, int64_t rows_out, int64_t cols_out, int64_t ld_out, void* stream);

/* every fp32 master weight of the model -> its zero-padded bf16 operand copy in ONE launch (once per step; replaces
 * what torch autocast does per nn.Linear call).  desc_dev: DEVICE int64[n][8] = {src fp32 [rows, cols], dst bf16
 * [rows, ld_dst] (ld_dst < 0: fp32 destination with row pitch -ld_dst), rows, cols, ld_dst, first block, ld_src (0 =
 * cols), width (0 = |ld_dst|: destination columns written per row; columns cols..width-1 := 0)}; a block covers 2048
 * (row, column < width) elements and total_blocks = sum over tensors of ceil(rows*width / 2048).  width < ld_dst
 * writes a sub-block of a larger matrix (the diagonal blocks of the PNA tower operands, modules/pna_layer.py:102-118). */
int gt_cast_multi(const int64_t* desc_dev, int32_t n, int64_t total_blocks, void* stream);
/* dst_b[r, c] += src[r0_b + r, c0_b + c] for n <= 16 rectangular blocks of the fp32 matrix src [*, ld_src]: gradients of
 * the diagonal blocks of a block-diagonal operand added into the per-tower parameter gradients (all arrays HOST). */
int gt_add_blocks(const float* src, int32_t ld_src, int32_t n, float* const* dst_host, const int32_t* ld_dst_host,
                  const int32_t* r0_host, const int32_t* c0_host, const int32_t* rows_host, const int32_t* cols_host,
                  void* stream);

/* ---- token packing + LayerNorm (reference modules/utils.py:5-29 pad_batch,
 *      modules/transformer_encoder.py:50-57 CLS append + norm_input) --------------------------
 * layernorm over rows of [M,d] (ld == d): y = LN(drop(x) [+ resid]) * gamma + beta (dropout on the sub-layer output
 * as in norm1(src + dropout1(src2)), reference nn.TransformerEncoderLayer) ; saves the
 * pre-norm sum (needed by backward) only implicitly through mean/rstd + xhat recompute.
 * in_rows (optional int32 [M]): row r reads x[in_rows[r]] (negative -> row of zeros/cls). */
int gt_layernorm_fwd(int dt, const void* x, const void* resid, const int32_t* in_rows,
                     const float* cls, int64_t M, int32_t d, const float* gamma, const float* beta,
                     float eps, void* y, void* presum, float* mean_rstd, float drop_p,
                     const uint64_t* rng_state, uint64_t salt, void* stream);
/* dx_sum = dLN/d(presum); dx_drop (optional, needs out_rows == NULL) = dx_sum * keep/(1-p) = gradient of the dropped
 * operand x; dgamma/dbeta accumulated (pre-zeroed fp32 [d]).
 * out_rows (optional): scatter row r of dx to dx[out_rows[r]] (rows < 0: -1 accumulates into
 * dcls (fp32 [d], pre-zeroed), -2 dropped); untouched rows of dx must be pre-zeroed by caller. */
int gt_layernorm_bwd(int dt, const void* dy, const void* presum, const float* mean_rstd,
                     const int32_t* out_rows, int64_t M, int32_t d, const float* gamma,
                     void* dx, float* dgamma, float* dbeta, float* dcls, void* dx_drop, float drop_p,
                     const uint64_t* rng_state, uint64_t salt, void* stream);
/* plain row gather/scatter for tokens when no input LayerNorm is configured, and for the
 * public pad_batch API: dst[r,:] = src[rows[r],:] (rows<0 -> cls or zeros). */
int gt_gather_rows(int dt, const void* src, const int32_t* rows, const float* cls, int64_t M,
                   int32_t ld, void* dst, void* stream);
int gt_scatter_rows(int dt, const void* dsrc, const int32_t* rows, int64_t M, int32_t ld,
                    void* ddst, float* dcls, void* stream);
/* public pad_batch layout: padded [S,B,ld] and mask uint8 [B,S] from node features */
int gt_pad_batch_fwd(int dt, const void* h, const int32_t* node_off, int64_t B, int64_t S, int32_t ld,
                     void* padded, uint8_t* mask, void* stream);
int gt_pad_batch_bwd(int dt, const void* dpadded, const int32_t* node_off, const int32_t* node_graph,
                     int64_t B, int64_t S, int64_t N, int32_t ld, void* dh, void* stream);

/* ---- stage 2: masked multi-head self-attention over packed tokens (replaces
 *      F.multi_head_attention_forward, reference modules/transformer_encoder.py:28-32,59) -----
 * qkv [n_rows, 3*d] (q | k | v, heads contiguous inside each), per-row graph id tok_graph and
 * per-graph token ranges tok_off; keys of a row = all token rows of its graph (padding never
 * exists in the packed layout, so the -inf key mask of the reference is implicit).
 * key_start (optional int32 [B]): first valid key row of each graph when the rows of a graph
 * begin with padding (dense left-padded layout of the public TransformerNodeEncoder API);
 * NULL = tok_off[g].  out [n_rows, d]; lse fp32 [nhead, n_rows].  Dropout on the attention
 * probabilities (after softmax, as F.multi_head_attention_forward does) with drop_p/rng/salt.
 * impl: 0 auto, 1 CUDA-core, 2 tcgen05. */
/* optional per-batch metadata for the tcgen05 kernels (computed once per step, reused by every layer and head):
 * row_bounds int32 [n_rows][2] = key-row range [lo, hi) of every token row; tile_bounds int32 [ceil(n_rows/128)][2]
 * = (first row, number of rows) of the row range interacting with each 128-row tile.  NULL = derive in-kernel. */
int gt_mha_meta(const int32_t* tok_graph, const int32_t* tok_off, int64_t n_rows, int64_t B,
                int32_t* row_bounds, int32_t* tile_bounds, void* stream);
int gt_mha_fwd(int dt, const void* qkv, const int32_t* tok_graph, const int32_t* tok_off,
               const int32_t* key_start, const int32_t* row_bounds, const int32_t* tile_bounds, int64_t n_rows, int64_t B, int32_t nhead, int32_t dh,
               float scale, void* out, float* lse, float drop_p, const uint64_t* rng_state,
               uint64_t salt, int impl, void* stream);
int gt_mha_bwd(int dt, const void* qkv, const void* out, const void* dout, const float* lse,
               const int32_t* tok_graph, const int32_t* tok_off, const int32_t* key_start,
               const int32_t* row_bounds, const int32_t* tile_bounds,
               int64_t n_rows, int64_t B, int32_t nhead, int32_t dh, float scale, void* dqkv,
               float* delta, float drop_p, const uint64_t* rng_state, uint64_t salt, int impl,
               void* stream);

/* Tile-local attention for batches of SMALL graphs (every graph <= 128 tokens incl. <CLS>; same math, dropout hash and
 * lse as gt_mha_fwd/bwd).  gt_mha_local_tiles packs consecutive graphs greedily into graph-aligned tiles of <= 128
 * packed rows: tiles int32 [max_tiles][2] = (first row, rows), unused slots (0, 0); *count = tiles used, or -1 when the
 * batch violates the bound (a graph longer than 128 rows, more tiles than max_tiles): gt_mha_local_fwd (count may be
 * NULL) then writes a NaN row so that the loss fails loudly instead of silently skipping graphs.  max_tiles is a
 * host-side upper bound, e.g. min(B, ceil(n_rows / (129 - max_tokens_per_graph))).  A CTA = (tile, head) keeps Q, K, V
 * (and dO) of the tile in shared memory: forward = 2 tcgen05 MMA groups, backward = 5 (dQ, dK, dV in ONE launch, no
 * atomics, delta computed in-kernel).  bf16 only, dh in {32, 64}. */
int gt_mha_local_tiles(const int32_t* tok_off, int64_t B, int64_t max_tiles, int32_t* tiles, int32_t* count, void* stream);
int gt_mha_local_fwd(int dt, const void* qkv, const int32_t* row_bounds, const int32_t* tiles, const int32_t* count,
                     int64_t max_tiles, int64_t n_rows, int32_t nhead, int32_t dh, float scale, void* out, float* lse, float drop_p,
                     const uint64_t* rng_state, uint64_t salt, void* stream);
int gt_mha_local_bwd(int dt, const void* qkv, const void* out, const void* dout, const float* lse,
                     const int32_t* row_bounds, const int32_t* tiles, int64_t max_tiles, int64_t n_rows, int32_t nhead,
                     int32_t dh, float scale, void* dqkv, float drop_p, const uint64_t* rng_state, uint64_t salt,
                     void* stream);

/* Pooled-query attention of the LAST encoder layer: the model reads only the pooled row of the encoder output
 * (reference models/gnn_transformer.py:114-115 `transformer_out[-1]`), so that layer needs ONE query per graph.
 * q [B, d] (unscaled query rows), kv [n_rows, 2d] (k | v of every packed token), q_rows int32 [B] = packed row of each
 * query (dropout row id: the same probabilities are dropped as in a full-layer run), out [B, d], lse fp32 [B, nhead].
 * Backward: dq [B, d], dkv [n_rows, 2d] fully overwritten (rows beyond tok_off[B] := 0).  dh % 4 == 0, dh <= 64. */
int gt_mha_cls_fwd(int dt, const void* q, const void* kv, const int32_t* tok_off, const int32_t* q_rows,
                   int64_t n_rows, int64_t B, int32_t nhead, int32_t dh, float scale, void* out, float* lse,
                   float drop_p, const uint64_t* rng_state, uint64_t salt, void* stream);
int gt_mha_cls_bwd(int dt, const void* q, const void* kv, const void* out, const void* dout, const float* lse,
                   const int32_t* tok_off, const int32_t* q_rows, int64_t n_rows, int64_t B, int32_t nhead,
                   int32_t dh, float scale, void* dq, void* dkv, float drop_p, const uint64_t* rng_state,
                   uint64_t salt, void* stream);

/* ---- fused AdamW (reference main.py:178 optim.AdamW; trainers/base_trainer.py:36 optimizer.step()) ----------
 * One launch for all parameters.  desc_dev int64 [n][4] = {parameter pointer (fp32), offset of its gradient / moments in
 * the flat arenas (elements), numel, first block}; a block updates 2048 elements; total_blocks = sum ceil(numel/2048).
 * hyper_dev fp32 [5] = {lr, beta1, beta2, eps, weight_decay} and step_dev int64 [1] are DEVICE memory (graph replay sees
 * host updates); the call increments *step_dev first.  p = p (1 - lr wd) - lr/(1-b1^t) m / (sqrt(v)/sqrt(1-b2^t) + eps). */
int gt_adamw_multi(const int64_t* desc_dev, int32_t n, int64_t total_blocks, const float* grad_flat, float* m_flat,
                   float* v_flat, const float* hyper_dev, int64_t* step_dev, const float* clip_dev, void* stream);
/* clip_dev (optional DEVICE fp32[2] = {max_norm, sum of squares of the gradient arena}): torch.nn.utils.clip_grad_norm_
 * (reference trainers/base_trainer.py:34-35) folded into the optimizer's gradient read - every gradient is scaled by
 * min(1, max_norm / (sqrt(sumsq) + 1e-6)); the arena itself is left unscaled.  gt_sumsq writes out[0] = sum x^2 of a
 * 16-byte aligned fp32 buffer DETERMINISTICALLY (per-block partials in scratch, added in index order by the last block):
 * every rank of a data-parallel job derives the same clip coefficient from the same averaged gradients.  scratch: fp32
 * [n_scratch >= 2], zero-initialised once by the caller (the last element is a ticket the kernel re-arms itself);
 * with several ranks run it after the allreduce. */
int gt_sumsq(const float* x, int64_t n, float* out, float* scratch, int32_t n_scratch, void* stream);

/* ---- graph-level read-outs of the reference's baseline models (PyG global_mean_pool / global_max_pool, reference
 *      models/gnn.py:64-69, models/pna.py:74-79, models/transformer.py:49-54; global_add_pool = gt_segment_sum_sorted) --
 * mode 1 = mean, 2 = max over the rows [node_off[g], node_off[g+1]) of x [N, ld]; out fp32 [B, ld]; an empty graph gives
 * 0.  arg int32 [B, ld] (max only): row index of the first maximum, used by the backward (gradient to one winner per
 * (graph, channel), torch_scatter semantics).  Backward: dx [N, ld] (activation dtype) fully overwritten; node_graph < 0
 * (shape-bucket slack nodes) -> 0. */
int gt_segment_pool_fwd(int dt, int mode, const void* x, const int32_t* node_off, int64_t B, int32_t ld,
                        float* out, int32_t* arg, void* stream);
int gt_segment_pool_bwd(int dt, int mode, const float* dout, const int32_t* node_off, const int32_t* node_graph,
                        const int32_t* arg, int64_t N, int32_t ld, void* dx, void* stream);
/* eval read-out: out[r] = index of the first maximum of x[r, :cols] (reference dataset/code.py:64 torch.argmax) */
int gt_argmax_rows(const float* x, int64_t rows, int32_t cols, int64_t ldx, int64_t* out, void* stream);

/* ---- PNA multi-aggregator reduce (reference modules/pna_layer.py:131-167 via
 *      modules/pna/pna_module.py:43-51; aggregators.py:11-34; scalers.py:10-31) ---------------
 * x [N, ld]: layer input viewed as `towers` towers of F channels (d = towers*F <= ld, F % 4 == 0);
 * pj, pi [N, ld]: per-node tower projections W_j x_j and W_i x_i + b.  For every
 * target i and channel over its in-edges: mean, max, min, std of m = pi[i] + pj[src]; empty
 * segment -> mean = max = min = 0, std = sqrt(1e-5).  delta = avg_deg['log'] (pna_layer.py:92-97).
 * out [N, ld_out >= towers*13F]: per tower [x_t | 1*(mean,max,min,std) | amp*(..) | att*(..)],
 * i.e. exactly the operand of post_nns[t]; argmax/argmin int32 [N, ld] saved for the backward. */
int gt_pna_reduce_fwd(int dt, const void* x, const void* pj, const void* pi, int64_t N, int32_t towers,
                      int32_t F, int32_t ld, const int32_t* rowptr_dst, const int32_t* src_by_dst,
                      float delta, void* out, int32_t ld_out, int32_t* argmax, int32_t* argmin,
                      void* stream);
/* from dout [N, ld_out]: dpj accumulated with atomics (fp32 [N, ld], pre-zeroed), dpi [N, ld] and
 * the pass-through dx [N, ld] (gradient of the x_t slots) overwritten */
int gt_pna_reduce_bwd(int dt, const void* pj, const void* out, const void* dout, int64_t N,
                      int32_t towers, int32_t F, int32_t ld, int32_t ld_out, const int32_t* rowptr_dst,
                      const int32_t* src_by_dst, float delta, const int32_t* argmax,
                      const int32_t* argmin, float* dpj, void* dpi, void* dx, void* stream);

/* ---- fused losses of the reference's dataset adapters (caller side of the path; device-resident, no host sync) ----
 * BCE-with-logits averaged over the labelled (non-NaN) entries (reference dataset/mol.py:24-31): x, y fp32
 * [rows, cols] with row pitches ldx / ldy; acc fp32 [3] zeroed by the caller (sum, count, block ticket); loss fp32 [1].
 * Backward: dx [rows, dcols >= cols] (pad columns := 0) = g[0] * (sigmoid(x) - y) / count on labelled entries. */
int gt_bce_masked_fwd(const float* x, const float* y, int64_t rows, int32_t cols, int64_t ldx, int64_t ldy,
                      float* acc, float* loss, void* stream);
int gt_bce_masked_bwd(const float* x, const float* y, int64_t rows, int32_t cols, int64_t ldx, int64_t ldy,
                      const float* acc, const float* g, float* dx, int64_t lddx, int32_t dcols, void* stream);
/* mean cross-entropy over rows (reference dataset/code.py:39-45 per head, dataset/tud.py:25-27): target int64 with
 * element stride tstride; lse fp32 [rows] saved for the backward; acc fp32 [3] zeroed by the caller. */
int gt_ce_fwd(const float* x, const int64_t* target, int64_t tstride, int64_t rows, int32_t cols, int64_t ldx,
              float* lse, float* acc, float* loss, void* stream);
int gt_ce_bwd(const float* x, const int64_t* target, int64_t tstride, int64_t rows, int32_t cols, int64_t ldx,
              const float* lse, const float* g, float* dx, int64_t lddx, int32_t dcols, void* stream);

#ifdef __cplusplus
}
#endif
#endif
