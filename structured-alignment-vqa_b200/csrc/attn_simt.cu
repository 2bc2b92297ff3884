// a5: graph-weighted attention core on CUDA cores in fp32 (engine 1).
//
// Role: (i) the backward of the attention core (a11), (ii) the on-device cross-check of the tcgen05 forward
// kernel (attn_tcgen05.cu), (iii) head sizes the tensor-core kernel does not take (d not in {32,64,128}).
// Semantics follow modules.py:246-301 exactly, including the fp32 key-mask constant, softmax over ALL keys,
// multiplicative graph after the softmax, L1 renormalisation with the 1e-12 clamp, and the query mask.
#include <stdlib.h>

#include "common.cuh"

namespace savqa {

int colsum_bf16(const void* x, int64_t ld, int64_t rows, int cols, float* out, cudaStream_t stream);  // elementwise.cu

namespace {

constexpr int kMaxJ = 16;  // scores of one query row live in registers: Tk <= 32 * 16 = 512
constexpr float kMaskFill = -4294967296.0f;  // fp32 value of -2**32 + 1 (modules.py:261)

struct RowSoftmax {
  float w[kMaxJ];   // W  (attention weight before the query mask)
  float p[kMaxJ];   // P  (plain softmax)
  float r;          // sum |G*P|
  float sumw;       // sum W
};

// Computes P and W for one query row held across the warp (key j = jj*32 + lane).
__device__ __forceinline__ void row_weights(const float (&s)[kMaxJ], const float* __restrict__ grow, int Tk, int renorm, int lane,
                                            RowSoftmax& o) {
  float m = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj)
    if (jj * 32 + lane < Tk) m = fmaxf(m, s[jj]);
  m = warp_max(m);
  float z = 0.0f;
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) {
    const int j = jj * 32 + lane;
    o.p[jj] = (j < Tk) ? expf(s[jj] - m) : 0.0f;
    z += o.p[jj];
  }
  z = warp_sum(z);
  const float iz = 1.0f / z;
  float r = 0.0f, sa = 0.0f;
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) {
    const int j = jj * 32 + lane;
    o.p[jj] *= iz;
    float a = o.p[jj];
    if (renorm != 0) a *= (j < Tk) ? grow[j] : 0.0f;
    o.w[jj] = a;
    r += fabsf(a);
    sa += a;
  }
  r = warp_sum(r);
  sa = warp_sum(sa);
  o.r = r;
  float scale = 1.0f;
  if (renorm == 1) scale = 1.0f / fmaxf(r, 1e-12f);
  else if (renorm == 2) scale = 1.0f / (sa + 1e-7f);
  float sw = 0.0f;
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) {
    o.w[jj] *= scale;
    sw += o.w[jj];
  }
  o.sumw = warp_sum(sw);
}

// Shared-memory staging of one head's K and V ([Tk][d+2] bf16: the +2 keeps row-per-lane reads conflict free).
__device__ __forceinline__ void stage_kv(const savqa_attn_args_t& a, int n, int h, __nv_bfloat16* sK, __nv_bfloat16* sV) {
  const int d = a.d, dp = d + 2;
  const __nv_bfloat16* K = static_cast<const __nv_bfloat16*>(a.k);
  const __nv_bfloat16* V = static_cast<const __nv_bfloat16*>(a.v);
  for (int i = threadIdx.x; i < a.Tk * (d / 2); i += blockDim.x) {
    const int j = i / (d / 2), c = (i % (d / 2)) * 2;
    const long row = static_cast<long>(n) * a.Tk + j;
    *reinterpret_cast<uint32_t*>(sK + j * dp + c) = *reinterpret_cast<const uint32_t*>(K + row * a.ldk + h * d + c);
    *reinterpret_cast<uint32_t*>(sV + j * dp + c) = *reinterpret_cast<const uint32_t*>(V + row * a.ldv + h * d + c);
  }
}

__device__ __forceinline__ void row_scores(const savqa_attn_args_t& a, int n, int i, const float* sq, const __nv_bfloat16* sK, int lane,
                                           float (&s)[kMaxJ]) {
  const int d = a.d, dp = d + 2;
  const float sqrt_d = sqrtf(static_cast<float>(d));
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) {
    const int j = jj * 32 + lane;
    float acc = kMaskFill;
    if (j < a.Tk) {
      acc = 0.0f;
      const __nv_bfloat16* kr = sK + j * dp;
      for (int c = 0; c < d; c += 2) {
        const float2 kv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(kr + c));
        acc = fmaf(sq[c], kv.x, acc);
        acc = fmaf(sq[c + 1], kv.y, acc);
      }
      acc = acc / sqrt_d;  // divide AFTER the contraction (modules.py:254)
      if (a.key_on && a.key_on[static_cast<long>(n) * a.Tk + j] == 0.0f) acc = kMaskFill;
      if (a.causal && j > i) acc = kMaskFill;
    }
    s[jj] = acc;
  }
}

constexpr int kRowsPerCta = 32;

__global__ void __launch_bounds__(256) attn_fwd_simt_kernel(const savqa_attn_args_t a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int d = a.d, dp = d + 2;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sV = sK + a.Tk * dp;
  float* sW = reinterpret_cast<float*>(sV + a.Tk * dp);  // [8 warps][Tk]
  float* sQ = sW + 8 * a.Tk;                             // [8 warps][d]
  const int hn = blockIdx.x;  // = h * N + n  (the reference's head-major batch index)
  const int h = hn / a.N, n = hn % a.N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_kv(a, n, h, sK, sV);
  __syncthreads();
  const __nv_bfloat16* Q = static_cast<const __nv_bfloat16*>(a.q);
  float* wbuf = sW + warp * a.Tk;
  float* sq = sQ + warp * d;
  const int i_end = min(a.Tq, (blockIdx.y + 1) * kRowsPerCta);
  for (int i = blockIdx.y * kRowsPerCta + warp; i < i_end; i += 8) {
    const long qrow = static_cast<long>(n) * a.Tq + i;
    for (int c = lane; c < d; c += 32) sq[c] = __bfloat162float(Q[qrow * a.ldq + h * d + c]);
    __syncwarp();
    float s[kMaxJ];
    row_scores(a, n, i, sq, sK, lane, s);
    const float* grow = a.graph ? a.graph + static_cast<long>(n) * a.graph_n_stride + static_cast<long>(i) * a.graph_q_stride : nullptr;
    RowSoftmax rs;
    row_weights(s, grow, a.Tk, a.graph ? a.renorm : 0, lane, rs);
    const float qon = a.query_on ? a.query_on[qrow] : 1.0f;
#pragma unroll
    for (int jj = 0; jj < kMaxJ; ++jj) {
      const int j = jj * 32 + lane;
      if (j < a.Tk) {
        if (a.att) a.att[(static_cast<long>(hn) * a.Tq + i) * a.Tk + j] = rs.w[jj];
        wbuf[j] = rs.w[jj] * qon;
      }
    }
    __syncwarp();
    for (int c = lane; c < d; c += 32) {
      float acc = 0.0f;
      for (int j = 0; j < a.Tk; ++j) acc = fmaf(wbuf[j], __bfloat162float(sV[j * dp + c]), acc);
      a.out[qrow * a.ldo + h * d + c] = acc;
    }
    __syncwarp();
  }
}

// Backward, pass 1 (query-row parallel): recompute P and W, dW = (dO V^T) * qm, dS, dQ; spill dS/sqrt(d) and W'.
__global__ void __launch_bounds__(256) attn_bwd_rows_kernel(const savqa_attn_args_t a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int d = a.d, dp = d + 2;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sV = sK + a.Tk * dp;
  float* sW = reinterpret_cast<float*>(sV + a.Tk * dp);  // [8][Tk]
  float* sQ = sW + 8 * a.Tk;                             // [8][d]
  float* sG = sQ + 8 * d;                                // [8][d]  dO row
  const int hn = blockIdx.x;
  const int h = hn / a.N, n = hn % a.N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_kv(a, n, h, sK, sV);
  __syncthreads();
  const __nv_bfloat16* Q = static_cast<const __nv_bfloat16*>(a.q);
  __nv_bfloat16* dQ = static_cast<__nv_bfloat16*>(a.dq);
  float* wbuf = sW + warp * a.Tk;
  float* sq = sQ + warp * d;
  float* sg = sG + warp * d;
  const float sqrt_d = sqrtf(static_cast<float>(d));
  const long plane = static_cast<long>(a.H) * a.N * a.Tq * a.Tk;
  const int renorm = a.graph ? a.renorm : 0;
  const int i_end = min(a.Tq, (blockIdx.y + 1) * kRowsPerCta);
  for (int i = blockIdx.y * kRowsPerCta + warp; i < i_end; i += 8) {
    const long qrow = static_cast<long>(n) * a.Tq + i;
    for (int c = lane; c < d; c += 32) {
      sq[c] = __bfloat162float(Q[qrow * a.ldq + h * d + c]);
      sg[c] = a.dout[qrow * a.ld_dout + h * d + c];
    }
    __syncwarp();
    float s[kMaxJ];
    row_scores(a, n, i, sq, sK, lane, s);
    const float* grow = a.graph ? a.graph + static_cast<long>(n) * a.graph_n_stride + static_cast<long>(i) * a.graph_q_stride : nullptr;
    RowSoftmax rs;
    row_weights(s, grow, a.Tk, renorm, lane, rs);
    const float qon = a.query_on ? a.query_on[qrow] : 1.0f;
    // dW_j = qm * sum_c dO[c] V[j][c];  t = sum_j W_j dW_j
    float dw[kMaxJ];
    float t = 0.0f;
#pragma unroll
    for (int jj = 0; jj < kMaxJ; ++jj) {
      const int j = jj * 32 + lane;
      float acc = 0.0f;
      if (j < a.Tk) {
        const __nv_bfloat16* vr = sV + j * dp;
        for (int c = 0; c < d; c += 2) {
          const float2 vv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vr + c));
          acc = fmaf(sg[c], vv.x, acc);
          acc = fmaf(sg[c + 1], vv.y, acc);
        }
        acc *= qon;
      }
      dw[jj] = acc;
      t += rs.w[jj] * acc;
    }
    t = warp_sum(t);
    const bool clamped = (renorm == 1) && (rs.r < 1e-12f);
    float* dS_out = a.scratch + (static_cast<long>(hn) * a.Tq + i) * a.Tk;
    float* Wq_out = dS_out + plane;
#pragma unroll
    for (int jj = 0; jj < kMaxJ; ++jj) {
      const int j = jj * 32 + lane;
      if (j < a.Tk) {
        float ds;
        if (renorm == 1 && clamped) ds = rs.w[jj] * dw[jj] - rs.p[jj] * t;
        else if (renorm == 2) ds = rs.w[jj] * (dw[jj] - t) - rs.p[jj] * t * (1.0f - rs.sumw);
        else ds = rs.w[jj] * (dw[jj] - t);
        // masked scores are constants: no gradient flows into them
        if (a.key_on && a.key_on[static_cast<long>(n) * a.Tk + j] == 0.0f) ds = 0.0f;
        if (a.causal && j > i) ds = 0.0f;
        ds = ds / sqrt_d;
        wbuf[j] = ds;
        dS_out[j] = ds;
        Wq_out[j] = rs.w[jj] * qon;
      }
    }
    __syncwarp();
    for (int c = lane; c < d; c += 32) {
      float acc = 0.0f;
      for (int j = 0; j < a.Tk; ++j) acc = fmaf(wbuf[j], __bfloat162float(sK[j * dp + c]), acc);
      if (!(sq[c] > 0.0f)) acc = 0.0f;  // ReLU of the Q projection (modules.py:227)
      dQ[qrow * a.ld_dq + h * d + c] = __float2bfloat16_rn(acc);
    }
    __syncwarp();
  }
}

// Backward, pass 2 (key parallel): dK[j,:] = sum_i dS[i,j] Q[i,:],  dV[j,:] = sum_i W'[i,j] dO[i,:], ReLU-gated.
// Thread (jj, cg): key j0+jj, columns [cg*CPT, (cg+1)*CPT) of the head.
template <int CPT>
__global__ void __launch_bounds__(256) attn_bwd_keys_kernel(const savqa_attn_args_t a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int d = a.d;  // == 8 * CPT
  float* sQ = reinterpret_cast<float*>(smem);  // [Tq][d]
  float* sG = sQ + a.Tq * d;                   // [Tq][d] dO
  const int hn = blockIdx.x;
  const int h = hn / a.N, n = hn % a.N;
  const __nv_bfloat16* Q = static_cast<const __nv_bfloat16*>(a.q);
  for (int idx = threadIdx.x; idx < a.Tq * d; idx += blockDim.x) {
    const int i = idx / d, c = idx % d;
    const long qrow = static_cast<long>(n) * a.Tq + i;
    sQ[idx] = __bfloat162float(Q[qrow * a.ldq + h * d + c]);
    sG[idx] = a.dout[qrow * a.ld_dout + h * d + c];
  }
  __syncthreads();
  const int jj = threadIdx.x & 31, cg = threadIdx.x >> 5;
  const int j = blockIdx.y * 32 + jj;
  const long plane = static_cast<long>(a.H) * a.N * a.Tq * a.Tk;
  const float* dS = a.scratch + static_cast<long>(hn) * a.Tq * a.Tk;
  const float* Wq = dS + plane;
  float accK[CPT], accV[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) accK[c] = accV[c] = 0.0f;
  if (j < a.Tk) {
    for (int i = 0; i < a.Tq; ++i) {
      const float ds = dS[static_cast<long>(i) * a.Tk + j];
      const float wq = Wq[static_cast<long>(i) * a.Tk + j];
      const float* qi = sQ + i * d + cg * CPT;
      const float* gi = sG + i * d + cg * CPT;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        accK[c] = fmaf(ds, qi[c], accK[c]);
        accV[c] = fmaf(wq, gi[c], accV[c]);
      }
    }
    const long krow = static_cast<long>(n) * a.Tk + j;
    const __nv_bfloat16* K = static_cast<const __nv_bfloat16*>(a.k);
    const __nv_bfloat16* V = static_cast<const __nv_bfloat16*>(a.v);
    __nv_bfloat16* dK = static_cast<__nv_bfloat16*>(a.dk);
    __nv_bfloat16* dV = static_cast<__nv_bfloat16*>(a.dv);
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = h * d + cg * CPT + c;
      const float kk = __bfloat162float(K[krow * a.ldk + col]);
      const float vv = __bfloat162float(V[krow * a.ldv + col]);
      dK[krow * a.ld_dk + col] = __float2bfloat16_rn(kk > 0.0f ? accK[c] : 0.0f);
      dV[krow * a.ld_dv + col] = __float2bfloat16_rn(vv > 0.0f ? accV[c] : 0.0f);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// Tq == 1 (the decoder: one class-token query per sample, AttModel_x3.py:141-154): one warp per (sample, head).
// K and V rows are read straight from HBM/L2 (each is one 128-byte line at d = 64): once key-per-lane for the
// scores / dW, once channel-per-lane for the weighted sums.  No scratch, no second kernel.
// ------------------------------------------------------------------------------------------------------
constexpr int kRowWarps = 8;

__device__ __forceinline__ float dot_row_bf16(const float* __restrict__ x, const __nv_bfloat16* __restrict__ row, int d, bool vec) {
  float acc = 0.0f;
  if (vec) {
    for (int c = 0; c < d; c += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(row + c));
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), e = unpack_bf16x2(u.z), f = unpack_bf16x2(u.w);
      acc = fmaf(x[c], a.x, acc); acc = fmaf(x[c + 1], a.y, acc);
      acc = fmaf(x[c + 2], b.x, acc); acc = fmaf(x[c + 3], b.y, acc);
      acc = fmaf(x[c + 4], e.x, acc); acc = fmaf(x[c + 5], e.y, acc);
      acc = fmaf(x[c + 6], f.x, acc); acc = fmaf(x[c + 7], f.y, acc);
    }
  } else {
    for (int c = 0; c < d; c += 2) {
      const float2 kv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row + c));
      acc = fmaf(x[c], kv.x, acc);
      acc = fmaf(x[c + 1], kv.y, acc);
    }
  }
  return acc;
}

// scores of the single query against every key (key j = jj*32 + lane), same arithmetic as row_scores()
__device__ __forceinline__ void row1_scores(const savqa_attn_args_t& a, int n, int h, const float* sq, int lane, bool vec, float (&s)[kMaxJ]) {
  const int d = a.d;
  const float sqrt_d = sqrtf(static_cast<float>(d));
  const __nv_bfloat16* K = static_cast<const __nv_bfloat16*>(a.k);
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) {
    const int j = jj * 32 + lane;
    float acc = kMaskFill;
    if (j < a.Tk) {
      acc = dot_row_bf16(sq, K + (static_cast<long>(n) * a.Tk + j) * a.ldk + h * d, d, vec) / sqrt_d;
      if (a.key_on && a.key_on[static_cast<long>(n) * a.Tk + j] == 0.0f) acc = kMaskFill;
      if (a.causal && j > 0) acc = kMaskFill;
    }
    s[jj] = acc;
  }
}

__global__ void __launch_bounds__(kRowWarps * 32) attn_row1_fwd_kernel(const savqa_attn_args_t a, int vec) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int d = a.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wbuf = reinterpret_cast<float*>(smem) + warp * (a.Tk + d);  // [Tk] W', then [d] q
  float* sq = wbuf + a.Tk;
  const long wi = static_cast<long>(blockIdx.x) * kRowWarps + warp;  // work item: the H heads of one sample are neighbours, so that the CTAs / warps that
                                                       // run together read the same K / V rows (one DRAM page per key row)
  if (wi >= static_cast<long>(a.N) * a.H) return;
  const int n = static_cast<int>(wi / a.H), h = static_cast<int>(wi % a.H);
  const long hn = static_cast<long>(h) * a.N + n;  // the reference's head-major batch index (layout of `att`)
  const __nv_bfloat16* Q = static_cast<const __nv_bfloat16*>(a.q) + static_cast<long>(n) * a.ldq + h * d;
  for (int c = lane; c < d; c += 32) sq[c] = __bfloat162float(Q[c]);
  __syncwarp();
  float s[kMaxJ];
  row1_scores(a, n, h, sq, lane, vec != 0, s);
  const float* grow = a.graph ? a.graph + static_cast<long>(n) * a.graph_n_stride : nullptr;
  RowSoftmax rs;
  row_weights(s, grow, a.Tk, a.graph ? a.renorm : 0, lane, rs);
  const float qon = a.query_on ? a.query_on[n] : 1.0f;
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) {
    const int j = jj * 32 + lane;
    if (j < a.Tk) {
      if (a.att) a.att[hn * a.Tk + j] = rs.w[jj];
      wbuf[j] = rs.w[jj] * qon;
    }
  }
  __syncwarp();
  const __nv_bfloat16* V = static_cast<const __nv_bfloat16*>(a.v) + static_cast<long>(n) * a.Tk * a.ldv + h * d;
  for (int c = 2 * lane; c < d; c += 64) {
    float a0 = 0.0f, a1 = 0.0f;
    for (int j = 0; j < a.Tk; ++j) {
      const float2 vv = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(V + static_cast<long>(j) * a.ldv + c)));
      a0 = fmaf(wbuf[j], vv.x, a0);
      a1 = fmaf(wbuf[j], vv.y, a1);
    }
    float* o = a.out + static_cast<long>(n) * a.ldo + h * d + c;
    o[0] = a0;
    o[1] = a1;
  }
}

__global__ void __launch_bounds__(kRowWarps * 32) attn_row1_bwd_kernel(const savqa_attn_args_t a, int vec) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int d = a.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dsbuf = reinterpret_cast<float*>(smem) + warp * (2 * a.Tk + 2 * d);  // [Tk] dS/sqrt(d), [Tk] W', [d] q, [d] dO
  float* wbuf = dsbuf + a.Tk;
  float* sq = wbuf + a.Tk;
  float* sg = sq + d;
  const long wi = static_cast<long>(blockIdx.x) * kRowWarps + warp;  // work item: the H heads of one sample are neighbours, so that the CTAs / warps that
                                                       // run together read the same K / V rows (one DRAM page per key row)
  if (wi >= static_cast<long>(a.N) * a.H) return;
  const int n = static_cast<int>(wi / a.H), h = static_cast<int>(wi % a.H);
  const long hn = static_cast<long>(h) * a.N + n;  // the reference's head-major batch index (layout of `att`)
  const __nv_bfloat16* Q = static_cast<const __nv_bfloat16*>(a.q) + static_cast<long>(n) * a.ldq + h * d;
  const float* dO = a.dout + static_cast<long>(n) * a.ld_dout + h * d;
  for (int c = lane; c < d; c += 32) {
    sq[c] = __bfloat162float(Q[c]);
    sg[c] = dO[c];
  }
  __syncwarp();
  float s[kMaxJ];
  row1_scores(a, n, h, sq, lane, vec != 0, s);
  const int renorm = a.graph ? a.renorm : 0;
  const float* grow = a.graph ? a.graph + static_cast<long>(n) * a.graph_n_stride : nullptr;
  RowSoftmax rs;
  row_weights(s, grow, a.Tk, renorm, lane, rs);
  const float qon = a.query_on ? a.query_on[n] : 1.0f;
  const __nv_bfloat16* Kb = static_cast<const __nv_bfloat16*>(a.k) + static_cast<long>(n) * a.Tk * a.ldk + h * d;
  const __nv_bfloat16* Vb = static_cast<const __nv_bfloat16*>(a.v) + static_cast<long>(n) * a.Tk * a.ldv + h * d;
  float dw[kMaxJ];
  float t = 0.0f;
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) {
    const int j = jj * 32 + lane;
    float acc = 0.0f;
    if (j < a.Tk) acc = dot_row_bf16(sg, Vb + static_cast<long>(j) * a.ldv, d, vec != 0) * qon;
    dw[jj] = acc;
    t += rs.w[jj] * acc;
  }
  t = warp_sum(t);
  const bool clamped = (renorm == 1) && (rs.r < 1e-12f);
  const float sqrt_d = sqrtf(static_cast<float>(d));
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) {
    const int j = jj * 32 + lane;
    if (j < a.Tk) {
      float ds;
      if (renorm == 1 && clamped) ds = rs.w[jj] * dw[jj] - rs.p[jj] * t;
      else if (renorm == 2) ds = rs.w[jj] * (dw[jj] - t) - rs.p[jj] * t * (1.0f - rs.sumw);
      else ds = rs.w[jj] * (dw[jj] - t);
      if (a.key_on && a.key_on[static_cast<long>(n) * a.Tk + j] == 0.0f) ds = 0.0f;
      if (a.causal && j > 0) ds = 0.0f;
      dsbuf[j] = ds / sqrt_d;
      wbuf[j] = rs.w[jj] * qon;
    }
  }
  __syncwarp();
  // channel-parallel: lane owns channels c, c+1 of the head
  __nv_bfloat16* dQ = static_cast<__nv_bfloat16*>(a.dq) + static_cast<long>(n) * a.ld_dq + h * d;
  __nv_bfloat16* dK = static_cast<__nv_bfloat16*>(a.dk) + static_cast<long>(n) * a.Tk * a.ld_dk + h * d;
  __nv_bfloat16* dV = static_cast<__nv_bfloat16*>(a.dv) + static_cast<long>(n) * a.Tk * a.ld_dv + h * d;
  for (int c = 2 * lane; c < d; c += 64) {
    const float q0 = sq[c], q1 = sq[c + 1], g0 = sg[c], g1 = sg[c + 1];
    float dq0 = 0.0f, dq1 = 0.0f, bk0 = 0.0f, bk1 = 0.0f, bv0 = 0.0f, bv1 = 0.0f;
    for (int j = 0; j < a.Tk; ++j) {
      const float2 kk = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(Kb + static_cast<long>(j) * a.ldk + c)));
      const float2 vv = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(Vb + static_cast<long>(j) * a.ldv + c)));
      const float ds = dsbuf[j], w = wbuf[j];
      dq0 = fmaf(ds, kk.x, dq0);
      dq1 = fmaf(ds, kk.y, dq1);
      const float k0 = kk.x > 0.0f ? ds * q0 : 0.0f, k1 = kk.y > 0.0f ? ds * q1 : 0.0f;  // ReLU of the K projection
      const float v0 = vv.x > 0.0f ? w * g0 : 0.0f, v1 = vv.y > 0.0f ? w * g1 : 0.0f;    // ReLU of the V projection
      *reinterpret_cast<uint32_t*>(dK + static_cast<long>(j) * a.ld_dk + c) = pack_bf16x2(k0, k1);
      *reinterpret_cast<uint32_t*>(dV + static_cast<long>(j) * a.ld_dv + c) = pack_bf16x2(v0, v1);
      bk0 += k0; bk1 += k1; bv0 += v0; bv1 += v1;
    }
    if (!(q0 > 0.0f)) dq0 = 0.0f;  // ReLU of the Q projection
    if (!(q1 > 0.0f)) dq1 = 0.0f;
    *reinterpret_cast<uint32_t*>(dQ + c) = pack_bf16x2(dq0, dq1);
    if (a.dbq) { atomicAdd(a.dbq + h * d + c, dq0); atomicAdd(a.dbq + h * d + c + 1, dq1); }
    if (a.dbk) { atomicAdd(a.dbk + h * d + c, bk0); atomicAdd(a.dbk + h * d + c + 1, bk1); }
    if (a.dbv) { atomicAdd(a.dbv + h * d + c, bv0); atomicAdd(a.dbv + h * d + c + 1, bv1); }
  }
}

// ---- piece-parallel row kernels (head size a power of two, 16-byte aligned rows): lane = (key slot ks, 16-byte piece p) ----
// A warp sweeps the keys 32 / P at a time (P = d / 8 pieces per row), every lane moving 16 bytes per load; four sweeps are in
// flight at once.  Scores, dW and the weights travel through a per-warp shared-memory row.
constexpr int kPieceWarps = 4;
constexpr int kSweepUnroll = 4;

__device__ __forceinline__ float dot8(const float* __restrict__ x, const uint4& u) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
  float acc = x[0] * a.x;
  acc = fmaf(x[1], a.y, acc); acc = fmaf(x[2], b.x, acc); acc = fmaf(x[3], b.y, acc);
  acc = fmaf(x[4], c.x, acc); acc = fmaf(x[5], c.y, acc); acc = fmaf(x[6], e.x, acc); acc = fmaf(x[7], e.y, acc);
  return acc;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = e.x; f[7] = e.y;
}

// dst[j] = scale * <x, M[j, head slice]> for every key j (M = K or V rows of this sample/head); result written by the p == 0 lanes
__device__ __forceinline__ void sweep_dots(const __nv_bfloat16* __restrict__ base, long ld, int Tk, int P, int ks, int pc, const float* xp,
                                           float scale, float* dst) {
  const int kpi = 32 / P;
  for (int j0 = 0; j0 < Tk; j0 += kpi * kSweepUnroll) {
    uint4 u[kSweepUnroll];
#pragma unroll
    for (int i = 0; i < kSweepUnroll; ++i) {
      const int j = j0 + i * kpi + ks;
      u[i] = (j < Tk) ? __ldg(reinterpret_cast<const uint4*>(base + static_cast<long>(j) * ld) + pc) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < kSweepUnroll; ++i) {
      float part = dot8(xp, u[i]);
      for (int o = 1; o < P; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      const int j = j0 + i * kpi + ks;
      if (pc == 0 && j < Tk) dst[j] = part * scale;
    }
  }
}

__global__ void __launch_bounds__(kPieceWarps * 32) attn_row1_fwd_piece_kernel(const savqa_attn_args_t a) {
  extern __shared__ __align__(16) uint8_t smem[];
  pdl_trigger();
  pdl_wait();
  const int d = a.d, P = d >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sS = reinterpret_cast<float*>(smem) + warp * (a.Tk + d);  // [Tk] scores, then W'; [d] q
  float* sq = sS + a.Tk;
  const long wi = static_cast<long>(blockIdx.x) * kPieceWarps + warp;  // work item: the H heads of one sample are neighbours, so that the CTAs / warps that
                                                       // run together read the same K / V rows (one DRAM page per key row)
  if (wi >= static_cast<long>(a.N) * a.H) return;
  const int n = static_cast<int>(wi / a.H), h = static_cast<int>(wi % a.H);
  const long hn = static_cast<long>(h) * a.N + n;  // the reference's head-major batch index (layout of `att`)
  const int pc = lane % P, ks = lane / P;
  const __nv_bfloat16* Q = static_cast<const __nv_bfloat16*>(a.q) + static_cast<long>(n) * a.ldq + h * d;
  for (int c = lane; c < d; c += 32) sq[c] = __bfloat162float(Q[c]);
  __syncwarp();
  const __nv_bfloat16* Kb = static_cast<const __nv_bfloat16*>(a.k) + static_cast<long>(n) * a.Tk * a.ldk + h * d;
  const __nv_bfloat16* Vb = static_cast<const __nv_bfloat16*>(a.v) + static_cast<long>(n) * a.Tk * a.ldv + h * d;
  sweep_dots(Kb, a.ldk, a.Tk, P, ks, pc, sq + pc * 8, 1.0f, sS);
  __syncwarp();
  // softmax / graph / renormalisation: key j = jj*32 + lane (same arithmetic as the generic kernels)
  const float sqrt_d = sqrtf(static_cast<float>(d));
  float s[kMaxJ];
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) {
    const int j = jj * 32 + lane;
    float v = kMaskFill;
    if (j < a.Tk) {
      v = sS[j] / sqrt_d;
      if (a.key_on && a.key_on[static_cast<long>(n) * a.Tk + j] == 0.0f) v = kMaskFill;
      if (a.causal && j > 0) v = kMaskFill;
    }
    s[jj] = v;
  }
  const float* grow = a.graph ? a.graph + static_cast<long>(n) * a.graph_n_stride : nullptr;
  RowSoftmax rs;
  row_weights(s, grow, a.Tk, a.graph ? a.renorm : 0, lane, rs);
  const float qon = a.query_on ? a.query_on[n] : 1.0f;
  __syncwarp();
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) {
    const int j = jj * 32 + lane;
    if (j < a.Tk) {
      if (a.att) a.att[hn * a.Tk + j] = rs.w[jj];
      sS[j] = rs.w[jj] * qon;
    }
  }
  __syncwarp();
  // out[piece] = sum_j W'_j V[j][piece]
  const int kpi = 32 / P;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.0f;
  for (int j0 = 0; j0 < a.Tk; j0 += kpi * kSweepUnroll) {
    uint4 u[kSweepUnroll];
#pragma unroll
    for (int i = 0; i < kSweepUnroll; ++i) {
      const int j = j0 + i * kpi + ks;
      u[i] = (j < a.Tk) ? __ldg(reinterpret_cast<const uint4*>(Vb + static_cast<long>(j) * a.ldv) + pc) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < kSweepUnroll; ++i) {
      const int j = j0 + i * kpi + ks;
      const float w = (j < a.Tk) ? sS[j] : 0.0f;
      float f[8];
      unpack8(u[i], f);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = fmaf(w, f[c], acc[c]);
    }
  }
  for (int o = P; o < 32; o <<= 1) {
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
  }
  if (ks == 0) {
    float* o = a.out + static_cast<long>(n) * a.ldo + h * d + pc * 8;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = acc[c];
  }
}

__global__ void __launch_bounds__(kPieceWarps * 32) attn_row1_bwd_piece_kernel(const savqa_attn_args_t a) {
  extern __shared__ __align__(16) uint8_t smem[];
  pdl_trigger();
  pdl_wait();
  const int d = a.d, P = d >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sS = reinterpret_cast<float*>(smem) + warp * (3 * a.Tk + 2 * d);  // [Tk] scores -> dS/sqrt(d); [Tk] dW -> W'; [d] q; [d] dO
  float* sD = sS + a.Tk;
  float* sq = sD + a.Tk;
  float* sg = sq + d;
  const long wi = static_cast<long>(blockIdx.x) * kPieceWarps + warp;  // work item: the H heads of one sample are neighbours, so that the CTAs / warps that
                                                       // run together read the same K / V rows (one DRAM page per key row)
  if (wi >= static_cast<long>(a.N) * a.H) return;
  const int n = static_cast<int>(wi / a.H), h = static_cast<int>(wi % a.H);
  const long hn = static_cast<long>(h) * a.N + n;  // the reference's head-major batch index (layout of `att`)
  const int pc = lane % P, ks = lane / P;
  const __nv_bfloat16* Q = static_cast<const __nv_bfloat16*>(a.q) + static_cast<long>(n) * a.ldq + h * d;
  const float* dO = a.dout + static_cast<long>(n) * a.ld_dout + h * d;
  for (int c = lane; c < d; c += 32) {
    sq[c] = __bfloat162float(Q[c]);
    sg[c] = dO[c];
  }
  __syncwarp();
  const __nv_bfloat16* Kb = static_cast<const __nv_bfloat16*>(a.k) + static_cast<long>(n) * a.Tk * a.ldk + h * d;
  const __nv_bfloat16* Vb = static_cast<const __nv_bfloat16*>(a.v) + static_cast<long>(n) * a.Tk * a.ldv + h * d;
  const float qon = a.query_on ? a.query_on[n] : 1.0f;
  sweep_dots(Kb, a.ldk, a.Tk, P, ks, pc, sq + pc * 8, 1.0f, sS);
  sweep_dots(Vb, a.ldv, a.Tk, P, ks, pc, sg + pc * 8, qon, sD);  // dW_j = qm * <dO, V_j>
  __syncwarp();
  const float sqrt_d = sqrtf(static_cast<float>(d));
  const int renorm = a.graph ? a.renorm : 0;
  float s[kMaxJ], dw[kMaxJ];
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) {
    const int j = jj * 32 + lane;
    float v = kMaskFill, w = 0.0f;
    if (j < a.Tk) {
      v = sS[j] / sqrt_d;
      w = sD[j];
      if (a.key_on && a.key_on[static_cast<long>(n) * a.Tk + j] == 0.0f) v = kMaskFill;
      if (a.causal && j > 0) v = kMaskFill;
    }
    s[jj] = v;
    dw[jj] = w;
  }
  const float* grow = a.graph ? a.graph + static_cast<long>(n) * a.graph_n_stride : nullptr;
  RowSoftmax rs;
  row_weights(s, grow, a.Tk, renorm, lane, rs);
  float t = 0.0f;
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) t += rs.w[jj] * dw[jj];
  t = warp_sum(t);
  const bool clamped = (renorm == 1) && (rs.r < 1e-12f);
  __syncwarp();
#pragma unroll
  for (int jj = 0; jj < kMaxJ; ++jj) {
    const int j = jj * 32 + lane;
    if (j < a.Tk) {
      float ds;
      if (renorm == 1 && clamped) ds = rs.w[jj] * dw[jj] - rs.p[jj] * t;
      else if (renorm == 2) ds = rs.w[jj] * (dw[jj] - t) - rs.p[jj] * t * (1.0f - rs.sumw);
      else ds = rs.w[jj] * (dw[jj] - t);
      if (s[jj] == kMaskFill) ds = 0.0f;  // masked scores are constants
      sS[j] = ds / sqrt_d;
      sD[j] = rs.w[jj] * qon;
    }
  }
  __syncwarp();
  // piece pass: dQ = sum_j dS_j K_j;  dK_j = dS_j q;  dV_j = W'_j dO   (ReLU gates of the projections; bias-gradient sums)
  const int kpi = 32 / P;
  float qf[8], gf[8], dq[8], bk[8], bv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    qf[c] = sq[pc * 8 + c];
    gf[c] = sg[pc * 8 + c];
    dq[c] = bk[c] = bv[c] = 0.0f;
  }
  __nv_bfloat16* dK = static_cast<__nv_bfloat16*>(a.dk) + static_cast<long>(n) * a.Tk * a.ld_dk + h * d;
  __nv_bfloat16* dV = static_cast<__nv_bfloat16*>(a.dv) + static_cast<long>(n) * a.Tk * a.ld_dv + h * d;
  for (int j0 = 0; j0 < a.Tk; j0 += kpi * kSweepUnroll) {
    uint4 uk[kSweepUnroll], uv[kSweepUnroll];
#pragma unroll
    for (int i = 0; i < kSweepUnroll; ++i) {
      const int j = j0 + i * kpi + ks;
      const bool ok = j < a.Tk;
      uk[i] = ok ? __ldg(reinterpret_cast<const uint4*>(Kb + static_cast<long>(j) * a.ldk) + pc) : make_uint4(0u, 0u, 0u, 0u);
      uv[i] = ok ? __ldg(reinterpret_cast<const uint4*>(Vb + static_cast<long>(j) * a.ldv) + pc) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < kSweepUnroll; ++i) {
      const int j = j0 + i * kpi + ks;
      if (j < a.Tk) {
        const float ds = sS[j], w = sD[j];
        float kf[8], vf[8], ok_[8], ov_[8];
        unpack8(uk[i], kf);
        unpack8(uv[i], vf);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          dq[c] = fmaf(ds, kf[c], dq[c]);
          ok_[c] = kf[c] > 0.0f ? ds * qf[c] : 0.0f;
          ov_[c] = vf[c] > 0.0f ? w * gf[c] : 0.0f;
          bk[c] += ok_[c];
          bv[c] += ov_[c];
        }
        *(reinterpret_cast<uint4*>(dK + static_cast<long>(j) * a.ld_dk) + pc) =
            make_uint4(pack_bf16x2(ok_[0], ok_[1]), pack_bf16x2(ok_[2], ok_[3]), pack_bf16x2(ok_[4], ok_[5]), pack_bf16x2(ok_[6], ok_[7]));
        *(reinterpret_cast<uint4*>(dV + static_cast<long>(j) * a.ld_dv) + pc) =
            make_uint4(pack_bf16x2(ov_[0], ov_[1]), pack_bf16x2(ov_[2], ov_[3]), pack_bf16x2(ov_[4], ov_[5]), pack_bf16x2(ov_[6], ov_[7]));
      }
    }
  }
  for (int o = P; o < 32; o <<= 1) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      dq[c] += __shfl_xor_sync(0xffffffffu, dq[c], o);
      bk[c] += __shfl_xor_sync(0xffffffffu, bk[c], o);
      bv[c] += __shfl_xor_sync(0xffffffffu, bv[c], o);
    }
  }
  if (ks == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (!(qf[c] > 0.0f)) dq[c] = 0.0f;  // ReLU of the Q projection
    __nv_bfloat16* dQ = static_cast<__nv_bfloat16*>(a.dq) + static_cast<long>(n) * a.ld_dq + h * d;
    *(reinterpret_cast<uint4*>(dQ) + pc) =
        make_uint4(pack_bf16x2(dq[0], dq[1]), pack_bf16x2(dq[2], dq[3]), pack_bf16x2(dq[4], dq[5]), pack_bf16x2(dq[6], dq[7]));
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int col = h * d + pc * 8 + c;
      if (a.dbq) atomicAdd(a.dbq + col, dq[c]);
      if (a.dbk) atomicAdd(a.dbk + col, bk[c]);
      if (a.dbv) atomicAdd(a.dbv + col, bv[c]);
    }
  }
}

// ---- key-split row kernels: one CTA of kSplitWarps warps per (sample, head); every warp sweeps its own slice of the keys with
// all of its 16-byte loads in flight at once (the one-warp kernels above walk the keys in dependent rounds of 16: the decoder's
// cross-attention is a chain link, and its latency -- not its bandwidth -- is what the step pays).  The softmax of the whole row
// is recomputed by every warp from the shared score row (identical arithmetic, so the redundant writes agree bit for bit). ----
constexpr int kSplitWarps = 4;
constexpr int kSplitLoads = 8;

template <int MJ>
struct RowSoftmaxT {
  float w[MJ], p[MJ], r, sumw;
};

template <int MJ>
__device__ __forceinline__ void row_weights_t(const float (&s)[MJ], const float* __restrict__ grow, int Tk, int renorm, int lane,
                                              RowSoftmaxT<MJ>& o) {  // row_weights() for MJ * 32 keys
  float m = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < MJ; ++jj)
    if (jj * 32 + lane < Tk) m = fmaxf(m, s[jj]);
  m = warp_max(m);
  float z = 0.0f;
#pragma unroll
  for (int jj = 0; jj < MJ; ++jj) {
    const int j = jj * 32 + lane;
    o.p[jj] = (j < Tk) ? expf(s[jj] - m) : 0.0f;
    z += o.p[jj];
  }
  z = warp_sum(z);
  const float iz = 1.0f / z;
  float r = 0.0f, sa = 0.0f;
#pragma unroll
  for (int jj = 0; jj < MJ; ++jj) {
    const int j = jj * 32 + lane;
    o.p[jj] *= iz;
    float a = o.p[jj];
    if (renorm != 0) a *= (j < Tk) ? grow[j] : 0.0f;
    o.w[jj] = a;
    r += fabsf(a);
    sa += a;
  }
  r = warp_sum(r);
  sa = warp_sum(sa);
  o.r = r;
  float scale = 1.0f;
  if (renorm == 1) scale = 1.0f / fmaxf(r, 1e-12f);
  else if (renorm == 2) scale = 1.0f / (sa + 1e-7f);
  float sw = 0.0f;
#pragma unroll
  for (int jj = 0; jj < MJ; ++jj) {
    o.w[jj] *= scale;
    sw += o.w[jj];
  }
  o.sumw = warp_sum(sw);
}

template <int MJ>
__global__ void __launch_bounds__(kSplitWarps * 32) attn_row1_fwd_split_kernel(const savqa_attn_args_t a) {
  extern __shared__ __align__(16) uint8_t smem[];
  pdl_trigger();
  pdl_wait();
  const int d = a.d, P = d >> 3, kpi = 32 / P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sS = reinterpret_cast<float*>(smem);  // [Tk] raw scores
  float* sW = sS + a.Tk;                       // [Tk] W' = W * query mask
  float* sq = sW + a.Tk;                       // [d]  q
  float* sO = sq + d;                          // [kSplitWarps][d] partial outputs
  const long wi = blockIdx.x;  // work item: the H heads of one sample are neighbours, so that the CTAs / warps that
                                                       // run together read the same K / V rows (one DRAM page per key row)
  const int n = static_cast<int>(wi / a.H), h = static_cast<int>(wi % a.H);
  const long hn = static_cast<long>(h) * a.N + n;  // the reference's head-major batch index (layout of `att`)
  const int pc = lane % P, ks = lane / P;
  const __nv_bfloat16* Q = static_cast<const __nv_bfloat16*>(a.q) + static_cast<long>(n) * a.ldq + h * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) sq[c] = __bfloat162float(Q[c]);
  __syncthreads();
  const __nv_bfloat16* Kb = static_cast<const __nv_bfloat16*>(a.k) + static_cast<long>(n) * a.Tk * a.ldk + h * d;
  const __nv_bfloat16* Vb = static_cast<const __nv_bfloat16*>(a.v) + static_cast<long>(n) * a.Tk * a.ldv + h * d;
  const int per = ((a.Tk + kSplitWarps - 1) / kSplitWarps + kpi - 1) / kpi * kpi;  // keys per warp, a multiple of the keys per load
  const int j_lo = warp * per, j_hi = min(a.Tk, j_lo + per);
  for (int j0 = j_lo; j0 < j_hi; j0 += kpi * kSplitLoads) {
    uint4 u[kSplitLoads];
#pragma unroll
    for (int i = 0; i < kSplitLoads; ++i) {
      const int j = j0 + i * kpi + ks;
      u[i] = (j < j_hi) ? __ldg(reinterpret_cast<const uint4*>(Kb + static_cast<long>(j) * a.ldk) + pc) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < kSplitLoads; ++i) {
      float part = dot8(sq + pc * 8, u[i]);
      for (int o = 1; o < P; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      const int j = j0 + i * kpi + ks;
      if (pc == 0 && j < j_hi) sS[j] = part;
    }
  }
  __syncthreads();
  const float sqrt_d = sqrtf(static_cast<float>(d));
  float s[MJ];
#pragma unroll
  for (int jj = 0; jj < MJ; ++jj) {
    const int j = jj * 32 + lane;
    float v = kMaskFill;
    if (j < a.Tk) {
      v = sS[j] / sqrt_d;
      if (a.key_on && a.key_on[static_cast<long>(n) * a.Tk + j] == 0.0f) v = kMaskFill;
      if (a.causal && j > 0) v = kMaskFill;
    }
    s[jj] = v;
  }
  const float* grow = a.graph ? a.graph + static_cast<long>(n) * a.graph_n_stride : nullptr;
  RowSoftmaxT<MJ> rs;
  row_weights_t<MJ>(s, grow, a.Tk, a.graph ? a.renorm : 0, lane, rs);
  const float qon = a.query_on ? a.query_on[n] : 1.0f;
#pragma unroll
  for (int jj = 0; jj < MJ; ++jj) {
    const int j = jj * 32 + lane;
    if (j < a.Tk) {
      if (a.att && warp == 0) a.att[hn * a.Tk + j] = rs.w[jj];
      sW[j] = rs.w[jj] * qon;  // every warp writes the same value; it reads back only what it wrote itself
    }
  }
  __syncwarp();
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.0f;
  for (int j0 = j_lo; j0 < j_hi; j0 += kpi * kSplitLoads) {
    uint4 u[kSplitLoads];
#pragma unroll
    for (int i = 0; i < kSplitLoads; ++i) {
      const int j = j0 + i * kpi + ks;
      u[i] = (j < j_hi) ? __ldg(reinterpret_cast<const uint4*>(Vb + static_cast<long>(j) * a.ldv) + pc) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < kSplitLoads; ++i) {
      const int j = j0 + i * kpi + ks;
      const float w = (j < j_hi) ? sW[j] : 0.0f;
      float f[8];
      unpack8(u[i], f);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = fmaf(w, f[c], acc[c]);
    }
  }
  for (int o = P; o < 32; o <<= 1) {
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
  }
  if (ks == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) sO[warp * d + pc * 8 + c] = acc[c];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float o = 0.0f;
#pragma unroll
    for (int w = 0; w < kSplitWarps; ++w) o += sO[w * d + c];
    a.out[static_cast<long>(n) * a.ldo + h * d + c] = o;
  }
}

template <int MJ>
__global__ void __launch_bounds__(kSplitWarps * 32, 4) attn_row1_bwd_split_kernel(const savqa_attn_args_t a) {
  extern __shared__ __align__(16) uint8_t smem[];
  pdl_trigger();
  pdl_wait();
  const int d = a.d, P = d >> 3, kpi = 32 / P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sS = reinterpret_cast<float*>(smem);  // [Tk] raw scores
  float* sD = sS + a.Tk;                       // [Tk] dW = qm <dO, V_j>
  float* sG = sD + a.Tk;                       // [Tk] dS / sqrt(d)
  float* sW = sG + a.Tk;                       // [Tk] W'
  float* sq = sW + a.Tk;                       // [d]  q
  float* sg = sq + d;                          // [d]  dO
  float* sP = sg + d;                          // [kSplitWarps][3][d] partial dQ, bias-gradient sums of dK, dV
  const long wi = blockIdx.x;  // work item: the H heads of one sample are neighbours, so that the CTAs / warps that
                                                       // run together read the same K / V rows (one DRAM page per key row)
  const int n = static_cast<int>(wi / a.H), h = static_cast<int>(wi % a.H);
  const long hn = static_cast<long>(h) * a.N + n;  // the reference's head-major batch index (layout of `att`)
  const int pc = lane % P, ks = lane / P;
  const __nv_bfloat16* Q = static_cast<const __nv_bfloat16*>(a.q) + static_cast<long>(n) * a.ldq + h * d;
  const float* dO = a.dout + static_cast<long>(n) * a.ld_dout + h * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    sq[c] = __bfloat162float(Q[c]);
    sg[c] = dO[c];
  }
  __syncthreads();
  const __nv_bfloat16* Kb = static_cast<const __nv_bfloat16*>(a.k) + static_cast<long>(n) * a.Tk * a.ldk + h * d;
  const __nv_bfloat16* Vb = static_cast<const __nv_bfloat16*>(a.v) + static_cast<long>(n) * a.Tk * a.ldv + h * d;
  const float qon = a.query_on ? a.query_on[n] : 1.0f;
  const int per = ((a.Tk + kSplitWarps - 1) / kSplitWarps + kpi - 1) / kpi * kpi;
  const int j_lo = warp * per, j_hi = min(a.Tk, j_lo + per);
  for (int j0 = j_lo; j0 < j_hi; j0 += kpi * kSplitLoads) {
    uint4 uk[kSplitLoads], uv[kSplitLoads];
#pragma unroll
    for (int i = 0; i < kSplitLoads; ++i) {
      const int j = j0 + i * kpi + ks;
      const bool ok = j < j_hi;
      uk[i] = ok ? __ldg(reinterpret_cast<const uint4*>(Kb + static_cast<long>(j) * a.ldk) + pc) : make_uint4(0u, 0u, 0u, 0u);
      uv[i] = ok ? __ldg(reinterpret_cast<const uint4*>(Vb + static_cast<long>(j) * a.ldv) + pc) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < kSplitLoads; ++i) {
      float ps = dot8(sq + pc * 8, uk[i]), pd = dot8(sg + pc * 8, uv[i]);
      for (int o = 1; o < P; o <<= 1) {
        ps += __shfl_xor_sync(0xffffffffu, ps, o);
        pd += __shfl_xor_sync(0xffffffffu, pd, o);
      }
      const int j = j0 + i * kpi + ks;
      if (pc == 0 && j < j_hi) {
        sS[j] = ps;
        sD[j] = pd * qon;
      }
    }
  }
  __syncthreads();
  const float sqrt_d = sqrtf(static_cast<float>(d));
  const int renorm = a.graph ? a.renorm : 0;
  float s[MJ], dw[MJ];
#pragma unroll
  for (int jj = 0; jj < MJ; ++jj) {
    const int j = jj * 32 + lane;
    float v = kMaskFill, w = 0.0f;
    if (j < a.Tk) {
      v = sS[j] / sqrt_d;
      w = sD[j];
      if (a.key_on && a.key_on[static_cast<long>(n) * a.Tk + j] == 0.0f) v = kMaskFill;
      if (a.causal && j > 0) v = kMaskFill;
    }
    s[jj] = v;
    dw[jj] = w;
  }
  const float* grow = a.graph ? a.graph + static_cast<long>(n) * a.graph_n_stride : nullptr;
  RowSoftmaxT<MJ> rs;
  row_weights_t<MJ>(s, grow, a.Tk, renorm, lane, rs);
  float t = 0.0f;
#pragma unroll
  for (int jj = 0; jj < MJ; ++jj) t += rs.w[jj] * dw[jj];
  t = warp_sum(t);
  const bool clamped = (renorm == 1) && (rs.r < 1e-12f);
#pragma unroll
  for (int jj = 0; jj < MJ; ++jj) {
    const int j = jj * 32 + lane;
    if (j < a.Tk) {
      float ds;
      if (renorm == 1 && clamped) ds = rs.w[jj] * dw[jj] - rs.p[jj] * t;
      else if (renorm == 2) ds = rs.w[jj] * (dw[jj] - t) - rs.p[jj] * t * (1.0f - rs.sumw);
      else ds = rs.w[jj] * (dw[jj] - t);
      if (s[jj] == kMaskFill) ds = 0.0f;  // masked scores are constants
      sG[j] = ds / sqrt_d;                // every warp writes the same values; it reads back only its own
      sW[j] = rs.w[jj] * qon;
    }
  }
  __syncwarp();
  // piece pass over this warp's keys: dQ += dS_j K_j;  dK_j = dS_j q;  dV_j = W'_j dO  (ReLU gates of the projections, bias sums)
  float qf[8], gf[8], dq[8], bk[8], bv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    qf[c] = sq[pc * 8 + c];
    gf[c] = sg[pc * 8 + c];
    dq[c] = bk[c] = bv[c] = 0.0f;
  }
  __nv_bfloat16* dK = static_cast<__nv_bfloat16*>(a.dk) + static_cast<long>(n) * a.Tk * a.ld_dk + h * d;
  __nv_bfloat16* dV = static_cast<__nv_bfloat16*>(a.dv) + static_cast<long>(n) * a.Tk * a.ld_dv + h * d;
  for (int j0 = j_lo; j0 < j_hi; j0 += kpi * kSplitLoads) {
    uint4 uk[kSplitLoads], uv[kSplitLoads];
#pragma unroll
    for (int i = 0; i < kSplitLoads; ++i) {  // second read of the slice: L1 hits
      const int j = j0 + i * kpi + ks;
      const bool ok = j < j_hi;
      uk[i] = ok ? __ldg(reinterpret_cast<const uint4*>(Kb + static_cast<long>(j) * a.ldk) + pc) : make_uint4(0u, 0u, 0u, 0u);
      uv[i] = ok ? __ldg(reinterpret_cast<const uint4*>(Vb + static_cast<long>(j) * a.ldv) + pc) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < kSplitLoads; ++i) {
      const int j = j0 + i * kpi + ks;
      if (j < j_hi) {
        const float ds = sG[j], w = sW[j];
        float kf[8], vf[8], ok_[8], ov_[8];
        unpack8(uk[i], kf);
        unpack8(uv[i], vf);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          dq[c] = fmaf(ds, kf[c], dq[c]);
          ok_[c] = kf[c] > 0.0f ? ds * qf[c] : 0.0f;
          ov_[c] = vf[c] > 0.0f ? w * gf[c] : 0.0f;
          bk[c] += ok_[c];
          bv[c] += ov_[c];
        }
        *(reinterpret_cast<uint4*>(dK + static_cast<long>(j) * a.ld_dk) + pc) =
            make_uint4(pack_bf16x2(ok_[0], ok_[1]), pack_bf16x2(ok_[2], ok_[3]), pack_bf16x2(ok_[4], ok_[5]), pack_bf16x2(ok_[6], ok_[7]));
        *(reinterpret_cast<uint4*>(dV + static_cast<long>(j) * a.ld_dv) + pc) =
            make_uint4(pack_bf16x2(ov_[0], ov_[1]), pack_bf16x2(ov_[2], ov_[3]), pack_bf16x2(ov_[4], ov_[5]), pack_bf16x2(ov_[6], ov_[7]));
      }
    }
  }
  for (int o = P; o < 32; o <<= 1) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      dq[c] += __shfl_xor_sync(0xffffffffu, dq[c], o);
      bk[c] += __shfl_xor_sync(0xffffffffu, bk[c], o);
      bv[c] += __shfl_xor_sync(0xffffffffu, bv[c], o);
    }
  }
  if (ks == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      sP[(warp * 3 + 0) * d + pc * 8 + c] = dq[c];
      sP[(warp * 3 + 1) * d + pc * 8 + c] = bk[c];
      sP[(warp * 3 + 2) * d + pc * 8 + c] = bv[c];
    }
  }
  __syncthreads();
  if (warp == 0 && lane < P) {  // lane = piece: the four warps' partial sums in a fixed order, ReLU gate of the Q projection
    float tq[8], tk[8], tv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      tq[c] = tk[c] = tv[c] = 0.0f;
#pragma unroll
      for (int w = 0; w < kSplitWarps; ++w) {
        tq[c] += sP[(w * 3 + 0) * d + lane * 8 + c];
        tk[c] += sP[(w * 3 + 1) * d + lane * 8 + c];
        tv[c] += sP[(w * 3 + 2) * d + lane * 8 + c];
      }
      if (!(sq[lane * 8 + c] > 0.0f)) tq[c] = 0.0f;
    }
    __nv_bfloat16* dQ = static_cast<__nv_bfloat16*>(a.dq) + static_cast<long>(n) * a.ld_dq + h * d;
    *(reinterpret_cast<uint4*>(dQ) + lane) =
        make_uint4(pack_bf16x2(tq[0], tq[1]), pack_bf16x2(tq[2], tq[3]), pack_bf16x2(tq[4], tq[5]), pack_bf16x2(tq[6], tq[7]));
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int col = h * d + lane * 8 + c;
      if (a.dbq) atomicAdd(a.dbq + col, tq[c]);
      if (a.dbk) atomicAdd(a.dbk + col, tk[c]);
      if (a.dbv) atomicAdd(a.dbv + col, tv[c]);
    }
  }
}

bool row1_piece_ok(const savqa_attn_args_t* a, bool bwd) {
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int d = a->d;
  if (!(d == 8 || d == 16 || d == 32 || d == 64 || d == 128 || d == 256)) return false;
  if (a->ldk % 8 || a->ldv % 8 || !a16(a->k) || !a16(a->v)) return false;
  if (bwd && (a->ld_dq % 8 || a->ld_dk % 8 || a->ld_dv % 8 || !a16(a->dq) || !a16(a->dk) || !a16(a->dv))) return false;
  return true;
}

bool row1_vec_ok(const savqa_attn_args_t* a) {
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return a->d % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a16(a->k) && a16(a->v);
}

int check_common(const savqa_attn_args_t* a, const char* who) {
  SAVQA_REQUIRE(a, "%s: null args", who);
  SAVQA_REQUIRE(a->q && a->k && a->v, "%s: null q/k/v", who);
  SAVQA_REQUIRE(a->N > 0 && a->H > 0 && a->Tq > 0 && a->Tk > 0 && a->d > 0, "%s: empty problem", who);
  SAVQA_REQUIRE(a->d % 2 == 0, "%s: head size %d must be even", who, a->d);
  SAVQA_REQUIRE(a->Tk <= 32 * kMaxJ, "%s: Tk=%d exceeds the single-pass limit %d", who, a->Tk, 32 * kMaxJ);
  SAVQA_REQUIRE(a->renorm >= 0 && a->renorm <= 2, "%s: renorm mode %d", who, a->renorm);
  SAVQA_REQUIRE(a->ldq % 2 == 0 && a->ldk % 2 == 0 && a->ldv % 2 == 0, "%s: odd leading dimension", who);
  SAVQA_REQUIRE(static_cast<long>(a->N) * a->H <= 2147483647L, "%s: too many (sample, head) pairs", who);
  return SAVQA_OK;
}

}  // namespace

int attn_fwd_simt(const savqa_attn_args_t* a, cudaStream_t stream) {
  if (int rc = check_common(a, "savqa_graph_attn_fwd")) return rc;
  SAVQA_REQUIRE(a->out, "savqa_graph_attn_fwd: null out");
  count_launch(a->Tq == 1 ? LK_ATTN_ROW1_FWD : LK_ATTN_FWD_SIMT);
  if (a->Tq == 1 && row1_piece_ok(a, false) && a->Tk >= 32 && getenv("SAVQA_ROW1_SPLIT_OFF") == nullptr) {
    const size_t smem1 = (static_cast<size_t>(2) * a->Tk + static_cast<size_t>(1 + kSplitWarps) * a->d) * 4;
    const dim3 grid(static_cast<unsigned>(static_cast<long>(a->N) * a->H)), block(kSplitWarps * 32);
#define SAVQA_ROW1_FWD(MJ)                                                                                                            \
  do {                                                                                                                                \
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_row1_fwd_split_kernel<MJ>), smem1, "savqa_graph_attn_fwd (row kernel)")) \
      return rc;                                                                                                                      \
    SAVQA_CHECK_CUDA(launch_kernel(true, attn_row1_fwd_split_kernel<MJ>, grid, block, smem1, stream, *a));                            \
  } while (0)
    if (a->Tk <= 64) SAVQA_ROW1_FWD(2);
    else if (a->Tk <= 128) SAVQA_ROW1_FWD(4);
    else if (a->Tk <= 256) SAVQA_ROW1_FWD(8);
    else SAVQA_ROW1_FWD(16);
#undef SAVQA_ROW1_FWD
    return SAVQA_OK;
  }
  if (a->Tq == 1 && row1_piece_ok(a, false)) {
    const size_t smem1 = static_cast<size_t>(kPieceWarps) * (a->Tk + a->d) * 4;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_row1_fwd_piece_kernel), smem1, "savqa_graph_attn_fwd (row kernel)")) return rc;
    const long warps = static_cast<long>(a->N) * a->H;
    SAVQA_CHECK_CUDA(launch_kernel(true, attn_row1_fwd_piece_kernel, dim3(static_cast<unsigned>((warps + kPieceWarps - 1) / kPieceWarps)),
                                   dim3(kPieceWarps * 32), smem1, stream, *a));
    return SAVQA_OK;
  }
  if (a->Tq == 1) {
    const size_t smem1 = static_cast<size_t>(kRowWarps) * (a->Tk + a->d) * 4;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_row1_fwd_kernel), smem1, "savqa_graph_attn_fwd (row kernel)")) return rc;
    const long warps = static_cast<long>(a->N) * a->H;
    attn_row1_fwd_kernel<<<static_cast<unsigned>((warps + kRowWarps - 1) / kRowWarps), kRowWarps * 32, smem1, stream>>>(*a, row1_vec_ok(a) ? 1 : 0);
    SAVQA_CHECK_CUDA(cudaGetLastError());
    return SAVQA_OK;
  }
  const int dp = a->d + 2;
  const size_t smem = static_cast<size_t>(2) * a->Tk * dp * 2 + static_cast<size_t>(8) * a->Tk * 4 + static_cast<size_t>(8) * a->d * 4;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_fwd_simt_kernel), smem, "savqa_graph_attn_fwd (engine 1)")) return rc;
  dim3 grid(a->N * a->H, (a->Tq + kRowsPerCta - 1) / kRowsPerCta);
  attn_fwd_simt_kernel<<<grid, 256, smem, stream>>>(*a);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

int attn_bwd_simt(const savqa_attn_args_t* a, cudaStream_t stream) {
  if (int rc = check_common(a, "savqa_graph_attn_bwd")) return rc;
  SAVQA_REQUIRE(a->dout && a->dq && a->dk && a->dv, "savqa_graph_attn_bwd: null gradient buffer");
  SAVQA_REQUIRE(a->ld_dq % 2 == 0 && a->ld_dk % 2 == 0 && a->ld_dv % 2 == 0, "savqa_graph_attn_bwd: odd leading dimension");
  count_launch(a->Tq == 1 ? LK_ATTN_ROW1_BWD : LK_ATTN_BWD_SIMT);
  // (measured, N = 128: Tk = 128 52 -> 45 us, but Tk = 56 35 -> 38 us: short rows stay on the one-warp kernel)
  if (a->Tq == 1 && row1_piece_ok(a, true) && a->Tk >= 64 && getenv("SAVQA_ROW1_SPLIT_OFF") == nullptr) {
    const size_t smem1 = (static_cast<size_t>(4) * a->Tk + static_cast<size_t>(2 + 3 * kSplitWarps) * a->d) * 4;
    const dim3 grid(static_cast<unsigned>(static_cast<long>(a->N) * a->H)), block(kSplitWarps * 32);
#define SAVQA_ROW1_BWD(MJ)                                                                                                            \
  do {                                                                                                                                \
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_row1_bwd_split_kernel<MJ>), smem1, "savqa_graph_attn_bwd (row kernel)")) \
      return rc;                                                                                                                      \
    SAVQA_CHECK_CUDA(launch_kernel(true, attn_row1_bwd_split_kernel<MJ>, grid, block, smem1, stream, *a));                            \
  } while (0)
    if (a->Tk <= 64) SAVQA_ROW1_BWD(2);
    else if (a->Tk <= 128) SAVQA_ROW1_BWD(4);
    else if (a->Tk <= 256) SAVQA_ROW1_BWD(8);
    else SAVQA_ROW1_BWD(16);
#undef SAVQA_ROW1_BWD
    return SAVQA_OK;
  }
  if (a->Tq == 1 && row1_piece_ok(a, true)) {
    const size_t smem1 = static_cast<size_t>(kPieceWarps) * (3 * a->Tk + 2 * a->d) * 4;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_row1_bwd_piece_kernel), smem1, "savqa_graph_attn_bwd (row kernel)")) return rc;
    const long warps = static_cast<long>(a->N) * a->H;
    SAVQA_CHECK_CUDA(launch_kernel(true, attn_row1_bwd_piece_kernel, dim3(static_cast<unsigned>((warps + kPieceWarps - 1) / kPieceWarps)),
                                   dim3(kPieceWarps * 32), smem1, stream, *a));
    return SAVQA_OK;
  }
  if (a->Tq == 1) {
    const size_t smem1 = static_cast<size_t>(kRowWarps) * (2 * a->Tk + 2 * a->d) * 4;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_row1_bwd_kernel), smem1, "savqa_graph_attn_bwd (row kernel)")) return rc;
    const long warps = static_cast<long>(a->N) * a->H;
    attn_row1_bwd_kernel<<<static_cast<unsigned>((warps + kRowWarps - 1) / kRowWarps), kRowWarps * 32, smem1, stream>>>(*a, row1_vec_ok(a) ? 1 : 0);
    SAVQA_CHECK_CUDA(cudaGetLastError());
    return SAVQA_OK;
  }
  SAVQA_REQUIRE(a->scratch, "savqa_graph_attn_bwd: engine 1 with Tq > 1 needs the scratch buffer");
  SAVQA_REQUIRE(a->d == 16 || a->d == 32 || a->d == 64 || a->d == 128, "savqa_graph_attn_bwd: head size %d not in {16,32,64,128}", a->d);
  const int dp = a->d + 2;
  const size_t smem1 = static_cast<size_t>(2) * a->Tk * dp * 2 + static_cast<size_t>(8) * a->Tk * 4 + static_cast<size_t>(16) * a->d * 4;
  const size_t smem2 = static_cast<size_t>(2) * a->Tq * a->d * 4;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_bwd_rows_kernel), smem1, "savqa_graph_attn_bwd")) return rc;
  const void* keys_kernel = a->d == 16   ? reinterpret_cast<const void*>(attn_bwd_keys_kernel<2>)
                            : a->d == 32 ? reinterpret_cast<const void*>(attn_bwd_keys_kernel<4>)
                            : a->d == 64 ? reinterpret_cast<const void*>(attn_bwd_keys_kernel<8>)
                                         : reinterpret_cast<const void*>(attn_bwd_keys_kernel<16>);
  if (int rc = ensure_dynamic_smem(keys_kernel, smem2, "savqa_graph_attn_bwd")) return rc;
  dim3 grid1(a->N * a->H, (a->Tq + kRowsPerCta - 1) / kRowsPerCta);
  attn_bwd_rows_kernel<<<grid1, 256, smem1, stream>>>(*a);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  dim3 grid2(a->N * a->H, (a->Tk + 31) / 32);
  switch (a->d) {
    case 16: attn_bwd_keys_kernel<2><<<grid2, 256, smem2, stream>>>(*a); break;
    case 32: attn_bwd_keys_kernel<4><<<grid2, 256, smem2, stream>>>(*a); break;
    case 64: attn_bwd_keys_kernel<8><<<grid2, 256, smem2, stream>>>(*a); break;
    default: attn_bwd_keys_kernel<16><<<grid2, 256, smem2, stream>>>(*a); break;
  }
  SAVQA_CHECK_CUDA(cudaGetLastError());
  const int Ch = a->H * a->d;
  if (a->dbq) if (int rc = colsum_bf16(a->dq, a->ld_dq, static_cast<long>(a->N) * a->Tq, Ch, a->dbq, stream)) return rc;
  if (a->dbk) if (int rc = colsum_bf16(a->dk, a->ld_dk, static_cast<long>(a->N) * a->Tk, Ch, a->dbk, stream)) return rc;
  if (a->dbv) if (int rc = colsum_bf16(a->dv, a->ld_dv, static_cast<long>(a->N) * a->Tk, Ch, a->dbv, stream)) return rc;
  return SAVQA_OK;
}

}  // namespace savqa
