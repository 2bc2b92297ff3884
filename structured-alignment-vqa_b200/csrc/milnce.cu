// f1: the object-word alignment head MIL_NCE (reference models/AttModel_x3.py:285-443, only_obj=True) between its small GEMMs,
// and f2: the scene-graph masks built from the loader's COMPACT hand-off (lengths + bit-packed adjacency) instead of dense
// int32 [B,T,T] planes.  Both are short HBM / latency bound kernels: one warp per object (MIL_NCE), one thread per output
// element (masks); no tensor-core shape here (per-object dot products of length h, topN <= 8 words).
#include "common.cuh"

namespace savqa {
namespace {

constexpr int kMaxTopN = 8;
constexpr float kMilEps = 1e-6f;  // `eps` of MIL_NCE.forward (AttModel_x3.py:345)

__device__ __forceinline__ float dot_bf16_row(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, int h, int lane) {
  float acc = 0.0f;
  for (int c = 2 * lane; c < h; c += 64) {
    const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(a + c));
    const float2 y = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(b + c));
    acc = fmaf(x.x, y.x, acc);
    acc = fmaf(x.y, y.y, acc);
  }
  return warp_sum(acc);
}

// softmax over n <= kMaxTopN values held by every lane (identical in all lanes)
__device__ __forceinline__ void softmax_small(const float* x, int n, float* p) {
  float m = x[0];
  for (int i = 1; i < n; ++i) m = fmaxf(m, x[i]);
  float z = 0.0f;
  for (int i = 0; i < n; ++i) {
    p[i] = expf(x[i] - m);
    z += p[i];
  }
  const float inv = 1.0f / z;
  for (int i = 0; i < n; ++i) p[i] *= inv;
}

// One warp per (sample b, object v).
//   raw_pos[i] = <pos_h[b,v,i,:], vis_h[b,v,:]>, raw_neg[i] likewise                                  (AttModel_x3.py:365-366)
//   term[b,v]  = logsumexp_i(max(0, eps)) - logsumexp_i(max(mask * raw_neg, eps))                       (:367; the positive rows of the
//                two concatenations are the same numbers and cancel exactly, value and gradient)
//   refined    = sum_i softmax_i(raw_pos) * pos_h[b,v,i,:]  ->  nodes[b, loc[b,v], :] when loc >= 0      (:372-379)
__global__ void __launch_bounds__(128) mil_nce_fwd_kernel(const __nv_bfloat16* __restrict__ pn_h, long ld_pn, const __nv_bfloat16* __restrict__ vis_h,
                                                          long ld_vis, const int* __restrict__ mask, const int64_t* __restrict__ loc,
                                                          __nv_bfloat16* __restrict__ nodes, long ld_nodes, int B, int V, int M, int topN, int h,
                                                          float* __restrict__ raw, float* __restrict__ term) {
  const int lane = threadIdx.x & 31;
  const long obj = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long n_obj = static_cast<long>(B) * V;
  if (obj >= n_obj) return;
  const long n_rows = n_obj * topN;
  const __nv_bfloat16* vis = vis_h + obj * ld_vis;
  float rp[kMaxTopN], rn[kMaxTopN], p[kMaxTopN], sn[kMaxTopN];
  for (int i = 0; i < topN; ++i) {
    const long r = obj * topN + i;
    rp[i] = dot_bf16_row(pn_h + r * ld_pn, vis, h, lane);
    rn[i] = dot_bf16_row(pn_h + (n_rows + r) * ld_pn, vis, h, lane);
    sn[i] = fmaxf(static_cast<float>(mask[r]) * rn[i], kMilEps);
  }
  softmax_small(rp, topN, p);
  if (lane == 0) {
    float m = sn[0];
    for (int i = 1; i < topN; ++i) m = fmaxf(m, sn[i]);
    float z = 0.0f;
    for (int i = 0; i < topN; ++i) z += expf(sn[i] - m);
    const float lse_neg = m + logf(z);
    const float lse_floor = kMilEps + logf(static_cast<float>(topN));
    term[obj] = lse_floor - lse_neg;
    for (int i = 0; i < topN; ++i) {
      raw[obj * topN + i] = rp[i];
      raw[n_rows + obj * topN + i] = rn[i];
    }
  }
  const long j = loc[obj];
  if (j < 0 || j >= M) return;
  __nv_bfloat16* dst = nodes + ((obj / V) * M + j) * ld_nodes;
  for (int c = 2 * lane; c < h; c += 64) {
    float a0 = 0.0f, a1 = 0.0f;
    for (int i = 0; i < topN; ++i) {
      const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(pn_h + (obj * topN + i) * ld_pn + c));
      a0 = fmaf(p[i], x.x, a0);
      a1 = fmaf(p[i], x.y, a1);
    }
    *reinterpret_cast<uint32_t*>(dst + c) = pack_bf16x2(a0, a1);
  }
}

// mil_nce_obj = mean over the [B, 2V, 1] tensor of :367 = sum_{b,v} term / (2 B V); one block, fixed summation order.
__global__ void __launch_bounds__(256) mil_nce_obj_kernel(const float* __restrict__ term, long n, float* __restrict__ obj) {
  __shared__ float sh[256];
  float acc = 0.0f;
  for (long i = threadIdx.x; i < n; i += 256) acc += term[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) obj[0] = sh[0] / (2.0f * static_cast<float>(n));
}

// Backward of the kernel above, one warp per (b, v).  d_nodes (fp32 [B*M, ld_dn], may be null) is the gradient of the node rows
// that feed ipt_mlp; d_obj (device scalar, may be null) the gradient of mil_nce_obj.  Outputs are the ReLU-gated gradients of the
// PRE-activation outputs of syb_mlp (pos rows, then neg rows) and vis_mlp, staged in bf16 as the next GEMMs' operands.
__global__ void __launch_bounds__(128) mil_nce_bwd_kernel(const __nv_bfloat16* __restrict__ pn_h, long ld_pn, const __nv_bfloat16* __restrict__ vis_h,
                                                          long ld_vis, const int* __restrict__ mask, const int64_t* __restrict__ loc,
                                                          const float* __restrict__ raw, const float* __restrict__ d_nodes, long ld_dn,
                                                          const float* __restrict__ d_obj, int B, int V, int M, int topN, int h,
                                                          __nv_bfloat16* __restrict__ d_pn, long ld_dpn, __nv_bfloat16* __restrict__ d_vis, long ld_dvis) {
  const int lane = threadIdx.x & 31;
  const long obj = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long n_obj = static_cast<long>(B) * V;
  if (obj >= n_obj) return;
  const long n_rows = n_obj * topN;
  const __nv_bfloat16* vis = vis_h + obj * ld_vis;
  const float g = d_obj ? d_obj[0] / (2.0f * static_cast<float>(n_obj)) : 0.0f;
  float rp[kMaxTopN], p[kMaxTopN], sn[kMaxTopN], q[kMaxTopN], dn[kMaxTopN], dp[kMaxTopN], draw[kMaxTopN];
  for (int i = 0; i < topN; ++i) {
    rp[i] = raw[obj * topN + i];
    const float mrn = static_cast<float>(mask[obj * topN + i]) * raw[n_rows + obj * topN + i];
    sn[i] = fmaxf(mrn, kMilEps);
    dn[i] = (mrn >= kMilEps) ? static_cast<float>(mask[obj * topN + i]) : 0.0f;  // clamp(min) passes the gradient where x >= min
  }
  softmax_small(rp, topN, p);
  softmax_small(sn, topN, q);
  for (int i = 0; i < topN; ++i) dn[i] *= -g * q[i];  // d(-logsumexp(s_neg)) / d raw_neg_i
  const long j = loc[obj];
  const bool hit = d_nodes != nullptr && j >= 0 && j < M;
  const float* dr = hit ? d_nodes + ((obj / V) * M + j) * ld_dn : nullptr;
  float s = 0.0f;
  for (int i = 0; i < topN; ++i) {
    float acc = 0.0f;
    if (hit) {
      const __nv_bfloat16* pos = pn_h + (obj * topN + i) * ld_pn;
      for (int c = 2 * lane; c < h; c += 64) {
        const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(pos + c));
        acc = fmaf(dr[c], x.x, acc);
        acc = fmaf(dr[c + 1], x.y, acc);
      }
    }
    dp[i] = warp_sum(acc);
    s = fmaf(p[i], dp[i], s);
  }
  for (int i = 0; i < topN; ++i) draw[i] = p[i] * (dp[i] - s);
  for (int c = 2 * lane; c < h; c += 64) {
    const float2 vv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vis + c));
    const float d0 = hit ? dr[c] : 0.0f, d1 = hit ? dr[c + 1] : 0.0f;
    float v0 = 0.0f, v1 = 0.0f;
    for (int i = 0; i < topN; ++i) {
      const long r = obj * topN + i;
      const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(pn_h + r * ld_pn + c));
      const float2 y = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(pn_h + (n_rows + r) * ld_pn + c));
      const float gp0 = x.x > 0.0f ? fmaf(p[i], d0, draw[i] * vv.x) : 0.0f;
      const float gp1 = x.y > 0.0f ? fmaf(p[i], d1, draw[i] * vv.y) : 0.0f;
      const float gn0 = y.x > 0.0f ? dn[i] * vv.x : 0.0f;
      const float gn1 = y.y > 0.0f ? dn[i] * vv.y : 0.0f;
      *reinterpret_cast<uint32_t*>(d_pn + r * ld_dpn + c) = pack_bf16x2(gp0, gp1);
      *reinterpret_cast<uint32_t*>(d_pn + (n_rows + r) * ld_dpn + c) = pack_bf16x2(gn0, gn1);
      v0 = fmaf(draw[i], x.x, fmaf(dn[i], y.x, v0));
      v1 = fmaf(draw[i], x.y, fmaf(dn[i], y.y, v1));
    }
    *reinterpret_cast<uint32_t*>(d_vis + obj * ld_dvis + c) = pack_bf16x2(vv.x > 0.0f ? v0 : 0.0f, vv.y > 0.0f ? v1 : 0.0f);
  }
}

// ------------------------------------------------------------------------------------------------------
// f2: masks from the compact hand-off.  collate_fn's masks are prefix blocks (ones on [:n, :n],
// data_loader_itp_bbox_super_node_onlyobj.py:355-358, 369-372, 412-414): a length per sample carries them; the 0/1
// adjacency matrices travel bit-packed (bit j of word w of row r = edge r -> 32 w + j).  Same outputs, bit for bit, as
// build_masks_kernel on the dense int32 planes -- plus the bit-packed forms of the two graphs the attention kernels read.
// One warp per (sample, row, 32-column word).
// ------------------------------------------------------------------------------------------------------
__global__ void build_masks_compact_kernel(const int* __restrict__ first_len, const int* __restrict__ q_len,
                                           const uint32_t* __restrict__ first_graph_bits, const uint32_t* __restrict__ q_graph_bits, int B, int V,
                                           int Q, int dec_mask_on, float* __restrict__ graph_diag, float* __restrict__ graph,
                                           float* __restrict__ dec_mask, uint32_t* __restrict__ diag_bits, uint32_t* __restrict__ graph_bits) {
  const int T = V + Q;
  const int wpr = (T + 31) / 32, wv = (V + 31) / 32, wq = (Q + 31) / 32;
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  const long total = static_cast<long>(B) * T * wpr;
  for (long i = warp0; i < total; i += nwarps) {
    const int w = static_cast<int>(i % wpr);
    const int r = static_cast<int>((i / wpr) % T);
    const long b = i / (static_cast<long>(wpr) * T);
    const int c = w * 32 + lane;
    const int vl = first_len[b], ql = q_len[b];
    float gd = 0.0f, g = 1.0f;
    if (c < T) {
      if (r >= V && c >= V) {
        const int rr = r - V, cc = c - V;
        gd = (rr < ql && cc < ql) ? 1.0f : 0.0f;
        g = ((q_graph_bits[(b * Q + rr) * wq + (cc >> 5)] >> (cc & 31)) & 1u) ? 1.0f : 0.0f;
      } else if (r < V && c < V) {
        g = first_graph_bits ? (((first_graph_bits[(b * V + r) * wv + (c >> 5)] >> (c & 31)) & 1u) ? 1.0f : 0.0f) : 1.0f;
      }
      graph_diag[(b * T + r) * T + c] = gd;
      graph[(b * T + r) * T + c] = g;
    }
    const uint32_t wd = __ballot_sync(0xffffffffu, c < T && gd != 0.0f);
    const uint32_t wg = __ballot_sync(0xffffffffu, c < T && g != 0.0f);
    if (lane == 0) {
      if (diag_bits) diag_bits[i] = wd;
      if (graph_bits) graph_bits[i] = wg;
    }
    // dec_mask[b,0,j] = (row j of block_diag(first_mask, q_mask) has a non-zero sum) = j inside its prefix block
    if (r == 0 && c < T) dec_mask[b * T + c] = dec_mask_on ? ((c < V ? c < vl : (c - V) < ql) ? 1.0f : 0.0f) : 0.0f;
  }
}

inline int blocks_for_warps(long warps, int warps_per_block) {
  long b = (warps + warps_per_block - 1) / warps_per_block;
  return static_cast<int>(b < 1 ? 1 : b);
}

}  // namespace
}  // namespace savqa

using namespace savqa;

extern "C" int savqa_mil_nce_fwd(const void* pn_h, int64_t ld_pn, const void* vis_h, int64_t ld_vis, const int32_t* mask, const int64_t* loc,
                                 void* nodes, int64_t ld_nodes, int B, int V, int M, int topN, int h, float* raw, float* term, float* obj,
                                 savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SAVQA_REQUIRE(pn_h && vis_h && mask && loc && nodes && raw && term && obj, "savqa_mil_nce_fwd: null tensor");
  SAVQA_REQUIRE(B > 0 && V > 0 && M > 0 && h > 0 && h % 2 == 0, "savqa_mil_nce_fwd: bad sizes B=%d V=%d M=%d h=%d", B, V, M, h);
  SAVQA_REQUIRE(topN >= 1 && topN <= kMaxTopN, "savqa_mil_nce_fwd: topN=%d outside [1, %d]", topN, kMaxTopN);
  SAVQA_REQUIRE(ld_pn % 2 == 0 && ld_vis % 2 == 0 && ld_nodes % 2 == 0, "savqa_mil_nce_fwd: odd leading dimension");
  const long n_obj = static_cast<long>(B) * V;
  count_launch(LK_MILNCE);
  mil_nce_fwd_kernel<<<blocks_for_warps(n_obj, 4), 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(pn_h), ld_pn, static_cast<const __nv_bfloat16*>(vis_h),
                                                                     ld_vis, mask, loc, static_cast<__nv_bfloat16*>(nodes), ld_nodes, B, V, M, topN, h, raw,
                                                                     term);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  mil_nce_obj_kernel<<<1, 256, 0, stream>>>(term, n_obj, obj);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_mil_nce_bwd(const void* pn_h, int64_t ld_pn, const void* vis_h, int64_t ld_vis, const int32_t* mask, const int64_t* loc,
                                 const float* raw, const float* d_nodes, int64_t ld_dn, const float* d_obj, int B, int V, int M, int topN, int h,
                                 void* d_pn, int64_t ld_dpn, void* d_vis, int64_t ld_dvis, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SAVQA_REQUIRE(pn_h && vis_h && mask && loc && raw && d_pn && d_vis, "savqa_mil_nce_bwd: null tensor");
  SAVQA_REQUIRE(B > 0 && V > 0 && M > 0 && h > 0 && h % 2 == 0 && topN >= 1 && topN <= kMaxTopN, "savqa_mil_nce_bwd: bad sizes");
  SAVQA_REQUIRE(ld_pn % 2 == 0 && ld_vis % 2 == 0 && ld_dpn % 2 == 0 && ld_dvis % 2 == 0, "savqa_mil_nce_bwd: odd leading dimension");
  const long n_obj = static_cast<long>(B) * V;
  count_launch(LK_MILNCE);
  mil_nce_bwd_kernel<<<blocks_for_warps(n_obj, 4), 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(pn_h), ld_pn, static_cast<const __nv_bfloat16*>(vis_h),
                                                                     ld_vis, mask, loc, raw, d_nodes, ld_dn, d_obj, B, V, M, topN, h,
                                                                     static_cast<__nv_bfloat16*>(d_pn), ld_dpn, static_cast<__nv_bfloat16*>(d_vis), ld_dvis);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_build_masks_compact(const int32_t* first_len, const int32_t* q_len, const uint32_t* first_graph_bits,
                                         const uint32_t* q_graph_bits, int B, int V, int Q, int dec_mask_on, float* graph_diag, float* graph,
                                         float* dec_mask, uint32_t* diag_bits, uint32_t* graph_bits, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SAVQA_REQUIRE(first_len && q_len && q_graph_bits && graph_diag && graph && dec_mask, "savqa_build_masks_compact: null tensor");
  SAVQA_REQUIRE(B > 0 && V > 0 && Q > 0, "savqa_build_masks_compact: bad sizes B=%d V=%d Q=%d", B, V, Q);
  const int T = V + Q;
  const long warps = static_cast<long>(B) * T * ((T + 31) / 32);
  long blocks = (warps + 7) / 8;
  const long cap = static_cast<long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  build_masks_compact_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(first_len, q_len, first_graph_bits, q_graph_bits, B, V, Q, dec_mask_on,
                                                                           graph_diag, graph, dec_mask, diag_bits, graph_bits);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}
