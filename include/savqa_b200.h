/* savqa_b200 -- C ABI of the B200-native graph-guided attention encoder path of SA-VQA.
 *
 * The reference (Peixixiong/Structured-Alignment-VQA) is pure Python on PyTorch and has NO plugin / FFI
 * boundary of its own; the seam this library plugs into is the nn.Module surface of models/modules.py and
 * the two branch models of models/AttModel_x3.py.  Each entry point below names the
 * reference code it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference
 * side (it is what structured-alignment-vqa_b200/savqa_b200/_lib.py does).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates; the library owns nothing
 *     except cached TMA descriptors); row-major; "ld*" are leading dimensions in ELEMENTS.
 *   - bf16 tensors are passed as void*; fp32 as float*.
 *   - `stream` is a cudaStream_t; kernels are only enqueued (no sync, CUDA-graph capturable).
 *   - return value: 0 on success, a SAVQA_ERR_* code otherwise; savqa_last_error() gives the message of the
 *     calling thread.  Never aborts, never falls back to a CPU path.
 *   - requires an sm_100a device (B200); savqa_device_check() reports anything else as an error.
 */
#ifndef SAVQA_B200_H_
#define SAVQA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAVQA_OK 0
#define SAVQA_ERR_BAD_ARGUMENT 1
#define SAVQA_ERR_CUDA 2
#define SAVQA_ERR_UNSUPPORTED 3

#define SAVQA_ABI_VERSION 5

typedef void* savqa_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------------- */
int savqa_abi_version(void);
const char* savqa_last_error(void);
/* 0 if the current device is sm_100 (B200); error otherwise.  *sm_count receives the SM count. */
int savqa_device_check(int* sm_count);

/* Cumulative number of launches by kernel family since the library was loaded, in this order: gemm_pair (CTA-pair tcgen05 GEMM),
 * gemm_single, attn_fwd_tc, attn_bwd_tc_shared (two-CTAs-per-SM shared-tile backward), attn_bwd_tc, attn_fwd_simt, attn_bwd_simt,
 * attn_row1_fwd, attn_row1_bwd, rowln_gemm, mil_nce.  Writes min(n, families) values and returns the number of families.  Test /
 * diagnostic aid: the parity tests assert which engine a shape actually took. */
int savqa_launch_counts(int64_t* out, int n);

/* ---- a4: scene-graph mask construction (AttModel_x3.py:103-122 vis, :229-247 syb) ------------------
 * first_mask [B,V,V], q_mask [B,Q,Q], q_graph [B,Q,Q], first_graph [B,V,V] or NULL (vis branch: top-left
 * block of `graph` is all ones).  Inputs are int32 when in_is_float == 0 (collate_fn output), fp32 otherwise.
 * Outputs fp32, T = V+Q: graph_diag [B,T,T] (only the Q x Q block = q_mask), graph [B,T,T]
 * (TL = 1 | first_graph, TR = BL = 1, BR = q_graph; the reference aliases graph_cross and graph),
 * dec_mask [B,1,T] (1 where the block-diagonal mask row has a non-zero sum; all zero if !dec_mask_on).
 * Bit exact with the reference. */
int savqa_build_masks(const void* first_mask, const void* q_mask, const void* q_graph, const void* first_graph,
                      int in_is_float, int B, int V, int Q, int dec_mask_on,
                      float* graph_diag, float* graph, float* dec_mask, savqa_stream_t stream);

/* ---- f2: the same masks from the loader's COMPACT hand-off ------------------------------------------------
 * collate_fn's masks are prefix blocks (ones on [:n,:n]; data_loader_itp_bbox_super_node_onlyobj.py:355-358, 369-372, 412-414),
 * so a length per sample carries them: first_len [B], q_len [B] (int32).  The 0/1 adjacency matrices travel bit-packed:
 * first_graph_bits [B,V,ceil(V/32)] (NULL: visual branch, top-left block all ones), q_graph_bits [B,Q,ceil(Q/32)], bit j of
 * word w of row r = edge r -> 32 w + j.  Outputs: graph_diag / graph / dec_mask exactly as savqa_build_masks writes them from
 * the dense planes (bit for bit), plus -- when non-NULL -- the bit-packed forms [B,T,ceil(T/32)] of graph_diag and graph that
 * the tcgen05 attention kernels read.  ~100 KB of input per step instead of ~30 MB of dense int32 planes. */
int savqa_build_masks_compact(const int32_t* first_len, const int32_t* q_len, const uint32_t* first_graph_bits,
                              const uint32_t* q_graph_bits, int B, int V, int Q, int dec_mask_on, float* graph_diag, float* graph,
                              float* dec_mask, uint32_t* diag_bits, uint32_t* graph_bits, savqa_stream_t stream);

/* bits[r, w] bit j = (graph[r, 32 w + j] != 0), zero past Tk, for r < rows, w < words_per_row (>= ceil(Tk / 32)).
 * Only meaningful for graphs whose entries are exactly 0 or 1 (the outputs of savqa_build_masks): the attention kernels
 * then read 4 bytes per 32 keys instead of 128. */
int savqa_pack_graph_bits(const float* graph, int64_t rows, int Tk, uint32_t* bits, int words_per_row, savqa_stream_t stream);

/* ---- a1/a2: embedding gathers (nn.Embedding at AttModel_x3.py:96,216; modules.py:32-46) ------------
 * out[r, 0:width] = table[idx[r], 0:width] * scale   (scale == 1.0f leaves the bits untouched).
 * out_f32 and/or out_bf16 may be NULL.  Columns [width, pad_to) of out_bf16 are zero-filled (TMA needs
 * 16-byte row pitches: 300 -> 304).  Index out of [0,table_rows) -> the row is zero-filled and the call
 * still succeeds (checked on the host side of the Python modules, like F.embedding). */
int savqa_gather_rows(const float* table, int64_t table_rows, int width, const int64_t* idx, int64_t n_idx, float scale,
                      float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16, int pad_to, savqa_stream_t stream);
/* dtable[idx[r], :] += dout[r, :] * scale, skipping rows whose index == skip_row (padding_idx semantics of
 * F.embedding: modules.py:34-41).  dense fp32 table gradient, atomics. */
int savqa_scatter_add_rows(float* dtable, int64_t table_rows, int width, const int64_t* idx, int64_t n_idx, const float* dout,
                           int64_t ld_dout, float scale, int64_t skip_row, savqa_stream_t stream);
/* The same accumulation into a signed Q15.48 fixed-point table (acc[idx[r], c] += round(dout[r, c] * scale * 2^48), 64-bit integer
 * atomics): integer addition is associative, so every data-parallel replica that accumulates the same gathered lists ends with the
 * same bits whatever order its atomics ran in -- what DDP's all-reduce of the dense embedding gradient guarantees in the reference
 * (main_itp_ddp_tar_super_node.py:404).  Resolution 3.6e-15, range +-32768 (saturating per addend); a NaN addend counts as 0.
 * savqa_adam_rows consumes the table with grad_q48 = 1. */
int savqa_scatter_add_rows_q48(int64_t* acc, int64_t table_rows, int width, const int64_t* idx, int64_t n_idx, const float* dout,
                               int64_t ld_dout, float scale, int64_t skip_row, savqa_stream_t stream);

/* ---- staging casts --------------------------------------------------------------------------------- */
/* dst_bf16[r, c] = src[r, c] for c < cols, 0 for cols <= c < pad_to. */
int savqa_cast_bf16(const float* src, int64_t ld_src, void* dst_bf16, int64_t ld_dst, int64_t rows, int cols, int pad_to,
                    savqa_stream_t stream);
/* dst_bf16[c, r] = src[r, c]  (weight transposes for dgrad), dst row pitch ld_dst, columns [rows, pad_to) zero. */
int savqa_cast_transpose_bf16(const float* src, int64_t ld_src, void* dst_bf16, int64_t ld_dst, int64_t rows, int cols, int pad_to,
                              savqa_stream_t stream);
/* Column regrouping: the matrix is `groups` groups of w_in columns per row; out gets `groups` groups of w_out columns, the first
 * min(w_in, w_out) columns of every group copied, the rest zero.  Pads the heads of a [rows, H*32] projection to [rows, H*64] (and
 * drops the padding again on the way back).  elem_bytes 2 (bf16) or 4 (fp32). */
int savqa_regroup_cols(const void* in, int64_t ld_in, void* out, int64_t ld_out, int64_t rows, int groups, int w_in, int w_out,
                       int elem_bytes, savqa_stream_t stream);
/* on[r] = (sum_c x[r, c] != 0) ? 1.0f : 0.0f  -- the activation-derived key / query masks
 * sign(abs(sum(x,-1))) of modules.py:257,289.  Optionally also writes a bf16 copy of x. */
int savqa_row_nonzero(const float* x, int64_t ld, int64_t rows, int cols, float* on, void* x_bf16, int64_t ld_bf16,
                      savqa_stream_t stream);
/* out_bf16[r,c] = (act_bf16[r,c] > 0) ? dy[src(r),c] : 0   -- ReLU backward staged as the bf16 GEMM operand.
 * dy is fp32 when dy_is_f32 != 0, bf16 otherwise.  src(r) = (r / group_rows) * group_stride + r % group_rows lets the rows of dy sit
 * in equally spaced groups inside a larger matrix (the node rows of every sample inside d x_in [B, T, 2048], AttModel_x3.py:218-219);
 * group_rows <= 0: src(r) = r. */
int savqa_relu_gate_bf16(const void* dy, int dy_is_f32, int64_t ld_dy, const void* act_bf16, int64_t ld_act, void* out_bf16,
                         int64_t ld_out, int64_t rows, int cols, int64_t group_rows, int64_t group_stride, savqa_stream_t stream);
/* Zero fill by at most max_blocks (<= 0: 8 per SM) persistent blocks: a background fill that does not hold the block scheduler. */
int savqa_fill_zero(void* p, int64_t bytes, int max_blocks, savqa_stream_t stream);
/* out[c] += sum_r x_bf16[r, c]  (bias gradients). */
int savqa_colsum_bf16(const void* x_bf16, int64_t ld, int64_t rows, int cols, float* out, savqa_stream_t stream);

/* ---- a6: residual + layer_normalization (modules.py:62-65; residuals at :304, :439) ------------------
 * pre = x (+ res);  y = gamma * (pre - mean) / (std_unbiased + eps) + beta.
 * Optional outputs: pre (saved for backward), y_bf16, on[r] = (sum_c y[r,c] != 0), stats[r] = {mean, sigma} (what a fused
 * dgrad + LayerNorm-backward epilogue reads, savqa_gemm_rowln mode 2). */
int savqa_residual_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, float eps,
                                 int64_t rows, int C, float* pre, float* y, void* y_bf16, float* on, float* stats,
                                 savqa_stream_t stream);
/* dx = LN'(pre)[dy] (+ dres_in);  dgamma/dbeta are ACCUMULATED (+=).  sigma == 0 rows follow autograd:
 * dx = (g - mean g) / eps.  dx_bf16 optional.  dxsum (optional, fp32 [C]) += sum_rows dx: the bias gradient of the
 * Linear whose output fed this LayerNorm (feedforward.conv2, modules.py:429, 439). */
int savqa_layernorm_bwd(const float* dy, const float* pre, const float* gamma, float eps, int64_t rows, int C,
                        const float* dres_in, float* dx, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum,
                        savqa_stream_t stream);

/* ---- a3/a5/a7: bf16 tensor-core GEMM with fused epilogue ----------------------------------------------
 * acc[m,n] = sum_k A[m,k] * B[n,k]   (nn.Linear: A = activations [M,K], B = weight [N,K])
 *   K-major operands (a_mn_major == 0): A is [M, lda] with k contiguous;   B is [N, ldb] with k contiguous.
 *   MN-major operands (== 1, used by wgrad): A is stored [K, lda] with m contiguous; B is [K, ldb] with n contiguous.
 * v = alpha*acc + bias[n] + res[m,n] + rowtab[(m % rowtab_period), n];  relu -> v = max(v,0);
 * gate -> v = (gate_bf16[m,n] > 0) ? v : 0  (ReLU backward);
 * out_f32 / out_bf16 receive v (either may be NULL).  accumulate: 0 store, 1 out_f32 += v, 2 atomic add
 * (required when split_k > 1).  Replaces nn.Linear (+ReLU) at modules.py:227-229, 428-429 and
 * AttModel_x3.py:42-44, 97-101. */
typedef struct savqa_gemm_epilogue {
  float alpha;
  int relu;
  int accumulate;
  int rowtab_period;
  const float* bias;
  const float* res;
  int64_t ld_res;
  const float* rowtab;
  int64_t ld_rowtab;
  const void* gate_bf16;
  int64_t ld_gate;
  float* out_f32;
  int64_t ld_out_f32;
  void* out_bf16;
  int64_t ld_out_bf16;
  float* colsum; /* optional fp32 [N]: colsum[n] += sum_m v[m,n] (bias gradient of the layer that produced A; atomics) */
} savqa_gemm_epilogue_t;

int savqa_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, int M, int N, int K,
                    const savqa_gemm_epilogue_t* epilogue, int split_k, savqa_stream_t stream);

/* `count` (1 or 2) independent GEMMs with the same N, operand majors and kind of output in ONE launch: the same layer of the
 * visual and the symbolic branch model (AttModel_x3.py:529-530) -- same shapes, different row counts (B*56 and B*128 tokens)
 * and different weights.  One launch halves the fixed launch / pipeline-fill cost and spreads the tiles of both problems over
 * the SMs together.  Falls back to one savqa_gemm_bf16 call per problem when the grouped kernel does not take the shapes. */
typedef struct savqa_gemm_problem {
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  int M, K;
  savqa_gemm_epilogue_t epilogue;
} savqa_gemm_problem_t;

int savqa_gemm_bf16_grouped(const savqa_gemm_problem_t* problems, int count, int a_mn_major, int b_mn_major, int N, int split_k,
                            savqa_stream_t stream);

/* ---- a9: cluster GEMM with a row-wise (LayerNorm) epilogue for the decoder's M = B-row chain (AttModel_x3.py:141-154) ----
 * acc[m,n] = sum_k A[m,k] B[n,k] (B K-major [N, ldb], or MN-major [K, ldb] when b_mn_major: a dgrad reads the weight that way).
 * A thread-block cluster computes a [128 x N] block, one 64-column slab per CTA; row reductions cross the CTAs through
 * distributed shared memory (csrc/rowln_tcgen05.cu).
 *   mode 0: v = gate?(relu?(acc + bias)) + res                            -> y (fp32) and / or y_bf16           (N % 64 == 0)
 *   mode 1: a = relu?(acc + bias) [-> act_bf16; the fp32 path continues from the ROUNDED value]; pre = a * rowscale[m] + res[m,n];
 *           y = gamma (pre - mean) / (sigma_unbiased + eps) + beta          -> pre, y, y_bf16, on[m] = (sum_n y != 0), stats[m] = {mean, sigma}
 *           replaces Linear(+ReLU) -> (* query mask) -> + residual -> layer_normalization (modules.py:62-65, 304-307, 439-447)
 *   mode 2: dy = acc + res; dx = d layer_normalization(pre)[dy] using stats    -> y = dx (fp32), y_bf16, dxg_bf16 = (gate_bf16 > 0) ? dx * rowscale : 0,
 *           dgamma[n] += sum_m dy c / s, dbeta[n] += sum_m dy, dxsum[n] += sum_m dx  (atomics)
 * modes 1 / 2 need N in {64, 128, 256, 512} (cluster of N / 64 CTAs).  Unused pointers are NULL. */
typedef struct savqa_rowln_args {
  const void* A; int64_t lda;
  const void* B; int64_t ldb; int b_mn_major;
  int M, N, K, mode, relu;
  const float* bias;
  const float* rowscale;
  const float* res; int64_t ld_res;
  const void* gate_bf16; int64_t ld_gate;
  const float* gamma; const float* beta; float eps;
  void* act_bf16; int64_t ld_act;
  float* pre; int64_t ld_pre;          /* mode 1: output; mode 2: input */
  float* y; int64_t ld_y;
  void* y_bf16; int64_t ld_yb;
  float* on;
  float* stats;                        /* [M, 2]; mode 1: output (optional); mode 2: input */
  void* dxg_bf16; int64_t ld_dxg;
  float* dgamma; float* dbeta; float* dxsum;
} savqa_rowln_args_t;

int savqa_gemm_rowln(const savqa_rowln_args_t* args, savqa_stream_t stream);

/* Scheduling hint for the calling thread's NEXT savqa_gemm_bf16* launches: the persistent CTA-pair kernel takes at most `sms`
 * SMs (0 = all).  The training step sets it around the GEMMs it puts on side streams (weight gradients, the decoder's K/V
 * projections of the encoder output, AttModel_x3.py:148-152): a persistent GEMM on every SM would make the decoder's chain of
 * small kernels on the main stream queue behind whole GEMMs.  Results do not depend on it.  Returns the previous value. */
int savqa_set_gemm_sm_limit(int sms);

/* ---- a5: graph-weighted attention core (modules.py:246-301 between the projections and the residual) ---
 * For each sample n and head h (channels [h*d,(h+1)*d) of q/k/v):
 *   S = Q K^T / sqrt(d);  S[:, j] = -4294967296 where !key_on[n,j];  optional causal tril mask;
 *   P = softmax(S);  renorm 0: W = P;  1: A = G*P, W = A / max(sum|A|, 1e-12);  2: W = A / (sum A + 1e-7);
 *   att[h*N+n] = W (optional, BEFORE the query mask);  W' = W * query_on[n,i];  O = W' V.
 * graph is fp32 [N,Tq,Tk]; graph_q_stride = Tk normally, 0 to broadcast one row over all queries ([N,1,Tk]).
 * engine: 0 = tcgen05/TMEM kernel (TMA-staged tiles), 1 = CUDA-core fp32 kernels: the verification kernel for Tq > 1
 * and the one-warp-per-(sample, head) row kernel for Tq == 1 (the decoder's single query, AttModel_x3.py:141-154). */
typedef struct savqa_attn_args {
  const void* q; int64_t ldq;      /* bf16 [N*Tq, ldq] */
  const void* k; int64_t ldk;      /* bf16 [N*Tk, ldk] */
  const void* v; int64_t ldv;      /* bf16 [N*Tk, ldv] */
  const float* graph; int64_t graph_n_stride; int64_t graph_q_stride;
  const float* key_on;             /* fp32 [N,Tk] */
  const float* query_on;           /* fp32 [N,Tq] */
  int N, H, Tq, Tk, d;
  int causal, renorm, engine;
  float* out; int64_t ldo;         /* fp32 [N*Tq, ldo] */
  float* att;                      /* fp32 [H*N,Tq,Tk] or NULL */
  /* backward only */
  const float* dout; int64_t ld_dout;   /* fp32 [N*Tq, ld_dout] */
  void* dq; int64_t ld_dq;         /* bf16, ReLU-gated: dq = dQ * (q > 0) */
  void* dk; int64_t ld_dk;
  void* dv; int64_t ld_dv;
  float* scratch;                  /* fp32 [2, H*N, Tq, Tk] workspace (dS and W'); engine 1 with Tq > 1 only */
  /* optional bias gradients of the Q/K/V projections: db*[c] += sum_rows gated d*[row, c]  (fp32 [H*d], atomics) */
  float* dbq; float* dbk; float* dbv;
  /* optional bit-packed form of a 0/1 `graph` (savqa_pack_graph_bits): word w of a row holds keys [32w, 32w+32); strides in
   * 32-bit words (bits_q_stride = 0 broadcasts one row).  The tcgen05 engine reads it instead of the fp32 matrix. */
  const uint32_t* graph_bits; int64_t bits_n_stride; int64_t bits_q_stride;
  /* softmax statistics of every (head, sample, query) row, fp32 [H*N*Tq, 4] = {row max m, 1/Z (negated when the 1e-12
   * clamp of the L1 renormalisation bit), scale (W = G e scale), beta}: written by the tcgen05 forward when non-NULL, read
   * -- together with the forward output `out` -- by the tcgen05 backward, which then needs ONE pass over the score tile
   * (sum_j W_j dW_j == <dO_row, out_row>) instead of three. */
  float* stats;
  /* head size the 1/sqrt(d) score scale is taken from when it differs from the tile width `d` (0: use d).  Heads of 32 channels run
   * on the tcgen05 engine as 64-wide tiles whose upper halves are zero (savqa_regroup_cols pads them): the contractions are
   * unchanged by the zero channels, only the scale must stay 1/sqrt(32). */
  int scale_d;
} savqa_attn_args_t;

int savqa_graph_attn_fwd(const savqa_attn_args_t* args, savqa_stream_t stream);
int savqa_graph_attn_bwd(const savqa_attn_args_t* args, savqa_stream_t stream);

/* ---- a10: label-smoothed three-head loss (main_itp_ddp_tar_super_node.py:335-345) ---------------------
 * loss = mean_b -sum_c t[b,c] * (lsm(lv)+lsm(ls)+lsm(lc))[b,c]/3, t = (1-eps)*onehot + eps/ncls.
 * Writes loss[0] and the three logit gradients (scaled by grad_scale; any may be NULL). */
int savqa_answer_loss(const float* logits_concat, const float* logits_vis, const float* logits_syb, const int64_t* answer, int B,
                      int ncls, float epsilon, float grad_scale, float* loss, float* d_concat, float* d_vis, float* d_syb,
                      savqa_stream_t stream);

/* ---- f1: MIL_NCE object-word alignment head (AttModel_x3.py:285-443, only_obj=True) between its Linear layers -----------
 * pn_h  bf16 [2*B*V*topN, ld_pn]: relu(syb_mlp(word rows)), the B*V*topN positive rows first, then the negative rows;
 * vis_h bf16 [B*V, ld_vis]: relu(vis_mlp(vis_fea));  mask int32 [B,V,topN];  loc int64 [B,V] (node position of object v, < 0: none).
 *   raw_pos[b,v,i] = <pos_h, vis_h>, raw_neg likewise                       (:365-366; written to raw[0 / 1][B*V*topN] for the backward)
 *   mil_nce_obj    = sum_{b,v} [lse_i(eps) - lse_i(max(mask*raw_neg, eps))] / (2 B V)   (:367 -- the positive halves of the two
 *                    concatenated logsumexp's are identical and cancel, in value and in gradient)          -> obj[0]
 *   nodes[b, loc[b,v], :] = sum_i softmax_i(raw_pos)[i] * pos_h[b,v,i,:]    (:372-379; nodes bf16 [B*M, ld_nodes] holds
 *                    relu(marco_mlp(.)) on entry, detached as at :354)
 * term fp32 [B*V] is workspace.  topN <= 8, h even. */
int savqa_mil_nce_fwd(const void* pn_h, int64_t ld_pn, const void* vis_h, int64_t ld_vis, const int32_t* mask, const int64_t* loc,
                      void* nodes, int64_t ld_nodes, int B, int V, int M, int topN, int h, float* raw, float* term, float* obj,
                      savqa_stream_t stream);
/* Backward: d_nodes fp32 [B*M, ld_dn] (gradient of the node rows, may be NULL), d_obj device scalar (gradient of mil_nce_obj, may
 * be NULL) -> ReLU-gated gradients of the pre-activations of syb_mlp (d_pn, layout of pn_h) and vis_mlp (d_vis), bf16. */
int savqa_mil_nce_bwd(const void* pn_h, int64_t ld_pn, const void* vis_h, int64_t ld_vis, const int32_t* mask, const int64_t* loc,
                      const float* raw, const float* d_nodes, int64_t ld_dn, const float* d_obj, int B, int V, int M, int topN, int h,
                      void* d_pn, int64_t ld_dpn, void* d_vis, int64_t ld_dvis, savqa_stream_t stream);

/* ---- f3: fused Adam over a flat fp32 parameter / gradient pair (torch.optim.Adam semantics,
 * main_itp_ddp_tar_super_node.py:206) and its row-sparse form for the word tables. ------------------------ */
/* dyn (device, may be NULL) = {lr / (1 - beta1^step), sqrt(1 - beta2^step), step}: read at run time instead of the host
 * scalars, so that a captured CUDA graph follows the step counter. */
/* param_bf16 (optional, same length): receives the bf16 copy of the updated parameters -- the MMA-operand mirror the
 * GEMMs read, so no per-step staging casts are needed. */
int savqa_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                    float beta2, float eps, int step, const float* dyn, void* param_bf16, savqa_stream_t stream);

/* dyn[2] += 1 (the step counter lives on the DEVICE), then dyn[0] = lr / (1 - beta1^step), dyn[1] = sqrt(1 - beta2^step).
 * One launch at the head of every step (captured in the step graph): a host that queues several steps ahead of the GPU, or a
 * replayed graph, can then never hand a step the scalars of another one.  Replaces the `state['step'] += 1` of torch.optim.Adam
 * (main_itp_ddp_tar_super_node.py:206, 366). */
int savqa_adam_advance(float* dyn, float lr, float beta1, float beta2, savqa_stream_t stream);

/* Row-sparse Adam for the 407000 x 300 word tables with the semantics of the reference's DENSE torch.optim.Adam ("deferred"
 * Adam).  row_stamp[row] = last step the row is current through (0: never touched).  Each row named in idx[0..n_idx) is handled
 * exactly once per call even if it occurs several times: the steps it missed since row_stamp are replayed with a zero gradient
 * (m *= beta1, v *= beta2, p -= lr_s m / (sqrt(v) / sqrt(1 - beta2^s) + eps): what dense Adam does to a row that is absent from a
 * batch), then
 *   apply != 0: the step-`step` update with grad (the dense table savqa_scatter_add_rows accumulated into as float, grad_q48 = 0,
 *               or savqa_scatter_add_rows_q48 as int64 fixed point, grad_q48 = 1; consumed rows are zeroed again, so the table
 *               never needs a memset);
 *   apply == 0: nothing more -- the catch-up through step - 1, run BEFORE the step's gathers read the rows; idx == NULL brings
 *               every row of the table up to date (before a checkpoint or an evaluation pass).
 * step is read from dyn[2] when dyn != NULL.  The replay of a row stops as soon as a zero-gradient step leaves every element unchanged
 * (the updates shrink by ~beta1 / sqrt(beta2) per step: from there on dense Adam's are lost in fp32 rounding too), at the latest
 * after 256 steps; the remaining decay of the moments is applied in closed form. */
int savqa_adam_rows(float* param, void* grad, int grad_q48, float* exp_avg, float* exp_avg_sq, int32_t* row_stamp, int64_t table_rows,
                    int width, const int64_t* idx, int64_t n_idx, float lr, float beta1, float beta2, float eps, int step,
                    const float* dyn, int apply, savqa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SAVQA_B200_H_ */
