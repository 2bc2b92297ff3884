// a11: backward of the graph-weighted attention core on the 5th-gen tensor cores (engine 0).
//
// One CTA (128 threads) per (sample n, head h); Tq <= 128 query rows, Tk <= 256 keys.
//   stage     Q, K, V by TMA (3-D maps, zero fill past the sample's rows); dO (fp32 in HBM) is converted to a bf16,
//             128B-swizzled tile by the CTA's threads
//   S  = Q K^T      tcgen05.mma -> TMEM [0, Tk)           } one commit
//   dWr = dO V^T    tcgen05.mma -> TMEM [dw_off, dw_off+Tk) }
//   rows      thread t owns query row t: recompute the forward's softmax statistics and W = G*e*scale (same arithmetic
//             as attn_tcgen05.cu), t = sum_j W dW, dS; bf16 dS and W' = W*qmask go to two swizzled smem tiles
//   dQ = dS K       A = dS  (K-major),            B = K  (MN-major)
//   dK = dS^T Q     A = dS  (MN-major, same tile), B = Q  (MN-major)     accumulators overlay the consumed S / dW columns
//   dV = W'^T dO    A = W'  (MN-major),           B = dO (MN-major)
//   epilogue  ReLU gates of the Q/K/V projections (modules.py:227-229) applied from q/k/v, bf16 stores into the
//             fused [.., 3C] gradient layout.
// Formulas: SURVEY.md Appendix A (verified against autograd); attn_simt.cu is the fp32 restatement of the same maths.
#include <stdlib.h>

#include "common.cuh"

namespace savqa {

int attn_bwd_tc(const savqa_attn_args_t* a, cudaStream_t stream);
bool attn_bwd_tc_fits(const savqa_attn_args_t* a);

namespace {

constexpr float kMaskFill = -4294967296.0f;

struct BwdParams {
  savqa_attn_args_t a;
  int tk_pad16;  // Tk rounded up to 16
  int kt;        // 128-key output tiles (1 or 2)
  int kc;        // 64-key chunks of the dS / W' tiles
  int kv_rows;   // rows of the K / V smem tiles (= TMA box rows)
  int dw_off;    // TMEM column of the raw dW accumulator
  int tmem_cols;
  int gvec;      // graph rows readable as float4
  int dvec;      // dout rows readable as float4
  int ovec;      // forward-output rows readable as float4 (statistics path)
};

__device__ __forceinline__ void tmem_alloc_rt(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_rt(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// bf16 row-gradient store with the ReLU gate of the projection output `act` (same layout as the gradient); rows that do
// not exist (!ok) store nothing.  When `dbias` is given the warp also adds the column sums of its 32 gated rows to it
// (the bias gradient of the projection, modules.py:227-229) -- warp-uniform call, 31 shuffles + one atomic per lane.
__device__ __forceinline__ void store_gated_row32(const uint32_t (&r)[32], const __nv_bfloat16* act, __nv_bfloat16* dst, bool ok,
                                                  float* dbias, int lane) {
  float gv[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) gv[j] = 0.0f;
  if (ok) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint4 av = __ldg(reinterpret_cast<const uint4*>(act) + u);
      const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack_bf16x2(aw[q]);
        const float lo = f.x > 0.0f ? __uint_as_float(r[8 * u + 2 * q]) : 0.0f;
        const float hi = f.y > 0.0f ? __uint_as_float(r[8 * u + 2 * q + 1]) : 0.0f;
        gv[8 * u + 2 * q] = lo;
        gv[8 * u + 2 * q + 1] = hi;
        o[q] = pack_bf16x2(lo, hi);
      }
      reinterpret_cast<uint4*>(dst)[u] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  if (dbias) {
    const float cs = warp_colsum32(gv, lane);
    atomicAdd(dbias + lane, cs);
  }
}

// g[j] = graph weight of column c0 + j of this thread's query row (1 when there is no graph, 0 past Tk); `bits_row`
// (shared memory, one word per 32 keys) replaces the fp32 row when the graph came bit-packed
__device__ __forceinline__ void load_graph32(const float* grow, const uint32_t* bits_row, bool use, int c0, int Tk, int gvec, float (&g)[32]) {
  if (use && bits_row) {
    const uint32_t word = bits_row[c0 >> 5];
#pragma unroll
    for (int j = 0; j < 32; ++j) g[j] = ((word >> j) & 1u) ? 1.0f : 0.0f;
  } else if (use && grow) {
    if (gvec && c0 + 32 <= Tk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(grow + c0) + j);
        g[4 * j] = v.x; g[4 * j + 1] = v.y; g[4 * j + 2] = v.z; g[4 * j + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) g[j] = (c0 + j < Tk) ? __ldg(grow + c0 + j) : 0.0f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) g[j] = 1.0f;
  }
}

// RS = thread-halves per query row: with RS == 2 (256 threads, chosen when Tk > 64 leaves room for only one CTA per SM) warps
// w and w + 4 share TMEM lane quadrant w & 3 and split the key columns of every row pass and the head columns of the
// epilogue; the row statistics are combined through shared memory.  One warp per scheduler cannot hide the exp / TMEM /
// shared-memory latencies of these dependent passes; two can.
template <int D, int RS>
__global__ void __launch_bounds__(128 * RS) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                              const __grid_constant__ CUtensorMap tmV, const BwdParams p) {
  constexpr int DCH = D / 64;
  constexpr int NT = 128 * RS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_tma, bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ float sStat[RS == 2 ? 2 * 128 * 4 : 1];  // per-half partial row statistics (recompute path only)
  __shared__ float sT[128];                            // t_i = <dO_i, O_i> = sum_j W_ij dW_ij  (statistics path)
  __shared__ uint32_t sKeyBits[8];                     // bit j of word w: key 32 w + j exists and is switched on (Tk <= 256)
  pdl_trigger();
  const savqa_attn_args_t& a = p.a;
  const int tid = threadIdx.x;
  const int t = tid & 127;       // query row / TMEM lane of this thread
  const int half = tid >> 7;     // which half of the columns it works on
  const int warp = t >> 5;       // TMEM lane quadrant
  // CTA order: the H heads of one sample are neighbours, so that the CTAs in flight together read the SAME rows of the fused
  // [q | k | v] projection (3 KB per token, 128 B of it per head and operand): one DRAM page serves all of them
  const int n = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int hn = h * a.N + n;  // the reference's head-major batch index (layout of `att` and of the row statistics)
  const int kc2 = p.kc;  // 64-key chunks of the dS / W' tiles that are written; the MMAs read whole 128-key tiles, so with an odd
                         // count the last tile's second chunk is whatever follows in shared memory (finite bf16 data of the next
                         // region): it only feeds dK / dV rows >= Tk, which are never stored

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sW = smem;                            // kc2 x [128][128 B]   W'
  uint8_t* sP = sW + kc2 * 16384;                // kc2 x [128][128 B]   dS
  uint8_t* sQ = sP + kc2 * 16384;                // DCH x [128][128 B]
  uint8_t* sdO = sQ + DCH * 16384;               // DCH x [128][128 B]
  uint8_t* sK = sdO + DCH * 16384;               // DCH x [kv_rows][128 B]
  uint8_t* sV = sK + DCH * p.kv_rows * 128;      // DCH x [kv_rows][128 B]
  float* sKeyOn = reinterpret_cast<float*>(sV + DCH * p.kv_rows * 128);  // [Tk]
  uint32_t* sBits = reinterpret_cast<uint32_t*>(sKeyOn + ((a.Tk + 3) & ~3));  // [128][wpr] bit-packed graph rows
  const int wpr = (a.Tk + 31) >> 5;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_o, 1);
    fence_barrier_init();
    pdl_wait();
    // Q / K / V travel while the CTA converts dO
    const uint32_t bytes = static_cast<uint32_t>(DCH) * (16384u + 2u * static_cast<uint32_t>(p.kv_rows) * 128u);
    mbar_arrive_expect_tx(&bar_tma, bytes);
    for (int c = 0; c < DCH; ++c) {
      tma_load_3d(sQ + c * 16384, &tmQ, &bar_tma, h * D + c * 64, 0, n);
      tma_load_3d(sK + c * p.kv_rows * 128, &tmK, &bar_tma, h * D + c * 64, 0, n);
      tma_load_3d(sV + c * p.kv_rows * 128, &tmV, &bar_tma, h * D + c * 64, 0, n);
    }
  }
  __syncwarp();
  if (tid < 32) tmem_alloc_rt(&tmem_slot, static_cast<uint32_t>(p.tmem_cols));
  pdl_wait();
  for (int j = tid; j < a.Tk; j += NT) sKeyOn[j] = a.key_on ? a.key_on[static_cast<long>(n) * a.Tk + j] : 1.0f;
  for (int w = tid >> 5; w < wpr; w += NT / 32) {
    const int col = w * 32 + (tid & 31);
    const bool on = col < a.Tk && (a.key_on == nullptr || a.key_on[static_cast<long>(n) * a.Tk + col] != 0.0f);
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if ((tid & 31) == 0) sKeyBits[w] = word;
  }
  if (a.graph_bits) {
    for (int idx = tid; idx < 128 * wpr; idx += NT) {
      const int row = idx / wpr, w = idx % wpr;
      sBits[idx] = (row < a.Tq) ? __ldg(a.graph_bits + static_cast<long>(n) * a.bits_n_stride + static_cast<long>(row) * a.bits_q_stride + w) : 0u;
    }
  }

  // ---- dO: fp32 [Tq, D] head slice -> bf16 K-major swizzled tile (rows >= Tq are zero) ----
  // all of a thread's loads are issued before the first conversion (the loop is latency bound otherwise: one HBM round trip
  // per iteration)
  {
    constexpr int V4 = D / 4;          // float4 per row
    constexpr int kIters = V4 / RS;    // 128 * V4 float4 over the CTA's threads
    float4 dv[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int idx = tid + it * NT;
      const int row = idx / V4, col = (idx % V4) * 4;
      dv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < a.Tq) {
        const float* src = a.dout + (static_cast<long>(n) * a.Tq + row) * a.ld_dout + h * D + col;
        if (p.dvec) dv[it] = __ldg(reinterpret_cast<const float4*>(src));
        else dv[it] = make_float4(src[0], src[1], src[2], src[3]);
      }
    }
    if (a.stats) {
      // t_row = <dO_row, O_row> over this head's columns: the V4 (16 or 32) consecutive threads that hold a row reduce it
      float part[kIters];
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int idx = tid + it * NT;
        const int row = idx / V4, col = (idx % V4) * 4;
        part[it] = 0.0f;
        if (row < a.Tq) {
          const float* src = a.out + (static_cast<long>(n) * a.Tq + row) * a.ldo + h * D + col;
          const float4 o = p.ovec ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(src[0], src[1], src[2], src[3]);
          // the bf16-rounded dO that the dW = dO V^T MMA sees: t then equals sum_j W~_j dW_j with the forward's own (bf16) W~
          const float2 d01 = unpack_bf16x2(pack_bf16x2(dv[it].x, dv[it].y)), d23 = unpack_bf16x2(pack_bf16x2(dv[it].z, dv[it].w));
          part[it] = (d01.x * o.x + d01.y * o.y) + (d23.x * o.z + d23.y * o.w);
        }
      }
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        float v = part[it];
#pragma unroll
        for (int o = (V4 < 32 ? V4 : 32) / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int idx = tid + it * NT;
        if ((idx % V4) == 0) sT[idx / V4] = v;
      }
    }
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int idx = tid + it * NT;
      const int row = idx / V4, col = (idx % V4) * 4;
      const int cc = col & 63;
      uint8_t* dst = sdO + (col >> 6) * 16384 + row * 128 + ((((cc >> 3) ^ (row & 7))) << 4) + (cc & 7) * 2;
      *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(dv[it].x, dv[it].y), pack_bf16x2(dv[it].z, dv[it].w));
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    mbar_wait(&bar_tma, 0);
    tc_fence_after();
    // S = Q K^T and raw dW = dO V^T (N up to 256 per instruction)
    for (int n0 = 0; n0 < p.tk_pad16; n0 += 256) {
      const int nn = min(256, p.tk_pad16 - n0);
      const uint32_t idesc = umma_idesc_bf16(128, nn, false, false);
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const int c = k / 4, kk = k % 4;
        const uint64_t adesc = umma_smem_desc(smem_u32(sQ + c * 16384) + kk * 32, 16, 1024);
        const uint64_t bdesc = umma_smem_desc(smem_u32(sK + (c * p.kv_rows + n0) * 128) + kk * 32, 16, 1024);
        umma_bf16_ss(tmem + n0, adesc, bdesc, idesc, k > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const int c = k / 4, kk = k % 4;
        const uint64_t adesc = umma_smem_desc(smem_u32(sdO + c * 16384) + kk * 32, 16, 1024);
        const uint64_t bdesc = umma_smem_desc(smem_u32(sV + (c * p.kv_rows + n0) * 128) + kk * 32, 16, 1024);
        umma_bf16_ss(tmem + p.dw_off + n0, adesc, bdesc, idesc, k > 0 ? 1u : 0u);
      }
    }
    umma_commit(&bar_s);
  }
  __syncwarp();

  // ---- row pass: thread t <-> query row t <-> TMEM lane t ----
  mbar_wait(&bar_s, 0);
  tc_fence_after();
  const int i = t;
  const bool row_ok = i < a.Tq;
  const long qrow = static_cast<long>(n) * a.Tq + (row_ok ? i : 0);
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const float inv_sqrt_d = 1.0f / sqrtf(static_cast<float>(a.scale_d > 0 ? a.scale_d : D));
  const int renorm = (a.graph || a.graph_bits) ? a.renorm : 0;
  const uint32_t* bits_row = a.graph_bits ? sBits + t * wpr : nullptr;
  const float* grow = (a.graph && row_ok) ? a.graph + static_cast<long>(n) * a.graph_n_stride + static_cast<long>(i) * a.graph_q_stride : nullptr;
  const float qon = (a.query_on && row_ok) ? a.query_on[qrow] : 1.0f;

  float m = -INFINITY, tsum, scale, alpha, beta, inv_z;
  if (a.stats) {
    // ---- statistics path: the forward kernel left {m, +-1/Z, scale, beta} per row, and t_i = <dO_i, O_i> ----
    const float4 st = __ldg(reinterpret_cast<const float4*>(a.stats) + static_cast<long>(hn) * a.Tq + (row_ok ? i : 0));
    m = st.x;
    inv_z = fabsf(st.y);
    scale = st.z;
    beta = st.w;
    alpha = st.y < 0.0f ? 0.0f : 1.0f;
    tsum = sT[t];
  } else {
  // ---- recompute path (no forward statistics): row max, then Z / R / SA / U, then the clamp logic of the forward ----
  // this half's share of the 32-column chunks of the score row
  const int nch = (a.Tk + 31) >> 5;
  const int ch_lo = (RS == 2 && half == 1) ? (nch + 1) / 2 : 0;
  const int ch_hi = (RS == 2 && half == 0) ? (nch + 1) / 2 : nch;
  for (int c0 = ch_lo * 32; c0 < ch_hi * 32; c0 += 32) {
    uint32_t r[32];
    __syncwarp();
    tmem_ld_32x32(t_lane + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = c0 + j;
      if (col < a.Tk) {
        float s = __uint_as_float(r[j]) * inv_sqrt_d;
        if (sKeyOn[col] == 0.0f) s = kMaskFill;
        if (a.causal && col > i) s = kMaskFill;
        m = fmaxf(m, s);
      }
    }
  }
  if constexpr (RS == 2) {  // row maximum over both halves
    sStat[(half * 128 + t) * 4] = m;
    __syncthreads();
    m = fmaxf(m, sStat[((half ^ 1) * 128 + t) * 4]);
    __syncthreads();
  }
  // statistics: Z = sum e, R = sum |g e|, SA = sum g e, U = sum g e * raw dW
  float Z = 0.0f, R = 0.0f, SA = 0.0f, U = 0.0f;
  for (int c0 = ch_lo * 32; c0 < ch_hi * 32; c0 += 32) {
    uint32_t r[32], w[32];
    __syncwarp();
    tmem_ld_32x32(t_lane + c0, r);
    tmem_ld_32x32(t_lane + p.dw_off + c0, w);
    tmem_ld_wait();
    float g[32];
    load_graph32(grow, bits_row, renorm != 0, c0, a.Tk, p.gvec, g);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = c0 + j;
      if (col < a.Tk) {
        float s = __uint_as_float(r[j]) * inv_sqrt_d;
        if (sKeyOn[col] == 0.0f) s = kMaskFill;
        if (a.causal && col > i) s = kMaskFill;
        const float e = ex2_approx(s == kMaskFill ? (kMaskFill - m) * kLog2e : fmaf(__uint_as_float(r[j]), inv_sqrt_d * kLog2e, -m * kLog2e));
        const float ge = g[j] * e;
        Z += e;
        R += fabsf(ge);
        SA += ge;
        U = fmaf(ge, __uint_as_float(w[j]), U);
      }
    }
  }
  if constexpr (RS == 2) {  // sums over both halves (added in the same order by both threads of a row: identical results)
    float* mine = sStat + (half * 128 + t) * 4;
    mine[0] = Z; mine[1] = R; mine[2] = SA; mine[3] = U;
    __syncthreads();
    const float* lo = sStat + t * 4;
    const float* hi = sStat + (128 + t) * 4;
    Z = lo[0] + hi[0]; R = lo[1] + hi[1]; SA = lo[2] + hi[2]; U = lo[3] + hi[3];
  }
  bool clamped = false;
  if (renorm == 1) {
    clamped = !(R / Z >= 1e-12f);
    scale = clamped ? 1.0f / (Z * 1e-12f) : 1.0f / R;
  } else if (renorm == 2) {
    scale = 1.0f / (SA + 1e-7f * Z);
  } else {
    scale = 1.0f / Z;
  }
  tsum = scale * qon * U;  // sum_j W_j dW_j
  // dS_j = W_j (dW_j - alpha t) - beta P_j t
  alpha = clamped ? 0.0f : 1.0f;
  beta = clamped ? 1.0f : (renorm == 2 ? 1.0f - scale * SA : 0.0f);
  inv_z = 1.0f / Z;
  }  // recompute path

  const int wch_lo = (RS == 2 && half == 1) ? kc2 : 0;            // 2 * kc2 chunks of 32 columns, zero fill past Tk included
  const int wch_hi = (RS == 2 && half == 0) ? kc2 : 2 * kc2;
  // Fast row pass (the training step): forward statistics, no causal mask, 0/1 graph bit-packed (or none).  Same e = 2^(raw c2 - m2)
  // as the forward kernel; the per-row factors are folded so that a score costs ~14 instructions (was ~40).
  const bool fast = a.stats != nullptr && !a.causal && (a.graph_bits != nullptr || renorm == 0);
  const float okf = row_ok ? 1.0f : 0.0f;
  const float c2 = inv_sqrt_d * kLog2e, m2 = m * kLog2e, marg = (kMaskFill - m) * kLog2e;
  const float sq = scale * qon * okf;                          // W' = ge sq
  const float ca = sq * inv_sqrt_d;                            // dS = ge (dW_raw ca - cb) - cc e, zero on masked keys
  const float cb = scale * alpha * tsum * inv_sqrt_d * okf;
  const float cc = beta * inv_z * tsum * inv_sqrt_d * okf;
  for (int c0 = wch_lo * 32; c0 < wch_hi * 32; c0 += 32) {
    float ds[32], wq[32];
    if (fast && c0 < a.Tk) {
      uint32_t r[32], w[32];
      __syncwarp();
      tmem_ld_32x32(t_lane + c0, r);
      tmem_ld_32x32(t_lane + p.dw_off + c0, w);
      const uint32_t vw = (c0 + 32 <= a.Tk) ? 0xffffffffu : ((1u << (a.Tk - c0)) - 1u);
      const uint32_t kw = sKeyBits[c0 >> 5];
      const uint32_t gw = renorm != 0 ? bits_row[c0 >> 5] : 0xffffffffu;
      tmem_ld_wait();
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float arg = fmaf(__uint_as_float(r[j]), c2, -m2);
        e[j] = ex2_approx(((kw >> j) & 1u) ? arg : marg);
      }
      if (vw != 0xffffffffu) {  // last, partial chunk (warp-uniform)
#pragma unroll
        for (int j = 0; j < 32; ++j) e[j] = ((vw >> j) & 1u) ? e[j] : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        // the weight exactly as the forward's P V MMA used it (bf16-rounded before the scale)
        const float ge = __bfloat162float(__float2bfloat16_rn(((gw >> j) & 1u) ? e[j] : 0.0f));
        const float dsv = fmaf(ge, fmaf(__uint_as_float(w[j]), ca, -cb), -cc * e[j]);
        ds[j] = ((kw >> j) & 1u) ? dsv : 0.0f;  // masked scores are constants
        wq[j] = ge * sq;
      }
    } else if (c0 < a.Tk) {  // warp-uniform: the TMEM loads are .sync.aligned
      uint32_t r[32], w[32];
      __syncwarp();
      tmem_ld_32x32(t_lane + c0, r);
      tmem_ld_32x32(t_lane + p.dw_off + c0, w);
      tmem_ld_wait();
      float g[32];
      load_graph32(grow, bits_row, renorm != 0, c0, a.Tk, p.gvec, g);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = c0 + j;
        float dsv = 0.0f, wv = 0.0f;
        if (col < a.Tk && row_ok) {
          float s = __uint_as_float(r[j]) * inv_sqrt_d;
          bool masked = false;
          if (sKeyOn[col] == 0.0f) { s = kMaskFill; masked = true; }
          if (a.causal && col > i) { s = kMaskFill; masked = true; }
          const float e = ex2_approx(s == kMaskFill ? (kMaskFill - m) * kLog2e : fmaf(__uint_as_float(r[j]), inv_sqrt_d * kLog2e, -m * kLog2e));
          // statistics path: the weight exactly as the forward's P V MMA used it (bf16-rounded before the scale), so that
          // sum_j dS_ij vanishes to fp32 accuracy against t = <dO~_i, O_i>
          const float ge = a.stats ? __bfloat162float(__float2bfloat16_rn(g[j] * e)) : g[j] * e;
          const float W = ge * scale;
          const float dW = __uint_as_float(w[j]) * qon;
          dsv = W * (dW - alpha * tsum) - beta * (e * inv_z) * tsum;
          if (masked) dsv = 0.0f;  // masked scores are constants
          dsv *= inv_sqrt_d;
          wv = W * qon;
        }
        ds[j] = dsv;
        wq[j] = wv;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) ds[j] = wq[j] = 0.0f;
    }
    const int off = (c0 >> 6) * 16384 + t * 128;
    const int u0 = (c0 & 63) >> 3;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int sw = ((u0 + u) ^ (t & 7)) << 4;
      *reinterpret_cast<uint4*>(sP + off + sw) = make_uint4(pack_bf16x2(ds[8 * u], ds[8 * u + 1]), pack_bf16x2(ds[8 * u + 2], ds[8 * u + 3]),
                                                           pack_bf16x2(ds[8 * u + 4], ds[8 * u + 5]), pack_bf16x2(ds[8 * u + 6], ds[8 * u + 7]));
      *reinterpret_cast<uint4*>(sW + off + sw) = make_uint4(pack_bf16x2(wq[8 * u], wq[8 * u + 1]), pack_bf16x2(wq[8 * u + 2], wq[8 * u + 3]),
                                                           pack_bf16x2(wq[8 * u + 4], wq[8 * u + 5]), pack_bf16x2(wq[8 * u + 6], wq[8 * u + 7]));
    }
  }

  // ---- dQ, dK, dV ----
  const int dq_off = 0, dk_off = D, dv_off = D + p.kt * D;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc_q = umma_idesc_bf16(128, 64, false, true);
    const uint32_t idesc_k = umma_idesc_bf16(128, 64, true, true);
    const int ksteps_keys = p.tk_pad16 / 16;
    const int ksteps_rows = (a.Tq + 15) / 16;
    for (int c = 0; c < DCH; ++c) {
      for (int k = 0; k < ksteps_keys; ++k) {
        const uint64_t adesc = umma_smem_desc(smem_u32(sP + (k >> 2) * 16384) + (k & 3) * 32, 16, 1024);
        const uint64_t bdesc = umma_smem_desc(smem_u32(sK + c * p.kv_rows * 128) + k * 2048, 8192, 1024);
        umma_bf16_ss(tmem + dq_off + c * 64, adesc, bdesc, idesc_q, k > 0 ? 1u : 0u);
      }
    }
    for (int kt = 0; kt < p.kt; ++kt) {
      for (int c = 0; c < DCH; ++c) {
        for (int k = 0; k < ksteps_rows; ++k) {
          const uint64_t a_ds = umma_smem_desc(smem_u32(sP + 2 * kt * 16384) + k * 2048, 16384, 1024);
          const uint64_t b_q = umma_smem_desc(smem_u32(sQ + c * 16384) + k * 2048, 8192, 1024);
          umma_bf16_ss(tmem + dk_off + kt * D + c * 64, a_ds, b_q, idesc_k, k > 0 ? 1u : 0u);
        }
        for (int k = 0; k < ksteps_rows; ++k) {
          const uint64_t a_w = umma_smem_desc(smem_u32(sW + 2 * kt * 16384) + k * 2048, 16384, 1024);
          const uint64_t b_do = umma_smem_desc(smem_u32(sdO + c * 16384) + k * 2048, 8192, 1024);
          umma_bf16_ss(tmem + dv_off + kt * D + c * 64, a_w, b_do, idesc_k, k > 0 ? 1u : 0u);
        }
      }
    }
    umma_commit(&bar_o);
  }
  __syncwarp();
  mbar_wait(&bar_o, 0);
  tc_fence_after();

  // ---- epilogue: ReLU gates + bf16 stores ----
  const __nv_bfloat16* Qg = static_cast<const __nv_bfloat16*>(a.q) + qrow * a.ldq + h * D;
  __nv_bfloat16* dQg = static_cast<__nv_bfloat16*>(a.dq) + qrow * a.ld_dq + h * D;
#pragma unroll 1
  for (int c0 = half * 32; c0 < D; c0 += 32 * RS) {
    uint32_t r[32];
    __syncwarp();
    tmem_ld_32x32(t_lane + dq_off + c0, r);
    tmem_ld_wait();
    store_gated_row32(r, Qg + c0, dQg + c0, row_ok, a.dbq ? a.dbq + h * D + c0 : nullptr, t & 31);
  }
#pragma unroll 1
  for (int kt = 0; kt < p.kt; ++kt) {
    const int j = kt * 128 + t;
    const bool key_ok = j < a.Tk;
    const long krow = static_cast<long>(n) * a.Tk + (key_ok ? j : 0);
    const __nv_bfloat16* Kg = static_cast<const __nv_bfloat16*>(a.k) + krow * a.ldk + h * D;
    const __nv_bfloat16* Vg = static_cast<const __nv_bfloat16*>(a.v) + krow * a.ldv + h * D;
    __nv_bfloat16* dKg = static_cast<__nv_bfloat16*>(a.dk) + krow * a.ld_dk + h * D;
    __nv_bfloat16* dVg = static_cast<__nv_bfloat16*>(a.dv) + krow * a.ld_dv + h * D;
#pragma unroll 1
    for (int c0 = half * 32; c0 < D; c0 += 32 * RS) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32(t_lane + dk_off + kt * D + c0, r);
      tmem_ld_wait();
      store_gated_row32(r, Kg + c0, dKg + c0, key_ok, a.dbk ? a.dbk + h * D + c0 : nullptr, t & 31);
      __syncwarp();
      tmem_ld_32x32(t_lane + dv_off + kt * D + c0, r);
      tmem_ld_wait();
      store_gated_row32(r, Vg + c0, dVg + c0, key_ok, a.dbv ? a.dbv + h * D + c0 : nullptr, t & 31);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc_rt(tmem, static_cast<uint32_t>(p.tmem_cols));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two-CTAs-per-SM variant for the training step (32 < Tk <= 128, Tq <= 128, d = 64, forward statistics,
// no causal mask, 0/1 graph bit-packed or absent).  The W' and dS tiles share ONE shared-memory tile: the row pass writes W'
// and keeps its 64 dS values packed in 32 registers; dV = W'^T dO is issued alone; once it has read the tile the dS values
// overwrite it and dQ / dK follow, while the dV accumulator is already being stored.  96 KB of tiles instead of 128 KB and
// 128 registers x 256 threads let two CTAs share an SM, so that one CTA's TMA / MMA / epilogue latencies hide under the
// other's row pass (the one-CTA kernel ran at IPC 0.24 with 42 % of its stalls on memory: profiles/r1_06_attn_bwd_lines.txt).
// ---------------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256, 2) attn_bwd_tc1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                             const __grid_constant__ CUtensorMap tmV, const BwdParams p) {
  constexpr int DCH = D / 64;
  constexpr int NT = 256;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_tma, bar_s, bar_v, bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ float sT[128];           // t_i = <dO_i, O_i>
  __shared__ uint32_t sKeyBits[4];    // bit j of word w: key 32 w + j exists and is switched on
  pdl_trigger();
  const savqa_attn_args_t& a = p.a;
  const int tid = threadIdx.x;
  const int t = tid & 127;       // query row / key row / TMEM lane of this thread
  const int half = tid >> 7;     // which 64 key columns (row pass) / which 32 head columns (epilogue) it works on
  const int warp = t >> 5;       // TMEM lane quadrant
  // CTA order: the H heads of one sample are neighbours, so that the CTAs in flight together read the SAME rows of the fused
  // [q | k | v] projection (3 KB per token, 128 B of it per head and operand): one DRAM page serves all of them
  const int n = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int hn = h * a.N + n;  // the reference's head-major batch index (layout of `att` and of the row statistics)

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int kc = p.kc;                           // 64-key chunks of the tile (1 or 2); each thread owns kc 32-column chunks of its row
  uint8_t* sX = smem;                            // kc x [128][128 B]   W', then dS  (kc == 1: the MN-major MMAs read a second chunk,
                                                 // the Q tile behind it -- finite values that only reach key rows >= Tk, never stored)
  uint8_t* sQ = sX + kc * 16384;                 // DCH x [128][128 B]
  uint8_t* sdO = sQ + DCH * 16384;               // DCH x [128][128 B]
  uint8_t* sK = sdO + DCH * 16384;               // DCH x [kv_rows][128 B]
  uint8_t* sV = sK + DCH * p.kv_rows * 128;      // DCH x [kv_rows][128 B]
  const int wpr = (a.Tk + 31) >> 5;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_v, 1);
    mbar_init(&bar_o, 1);
    fence_barrier_init();
    pdl_wait();
    const uint32_t bytes = static_cast<uint32_t>(DCH) * (16384u + 2u * static_cast<uint32_t>(p.kv_rows) * 128u);
    mbar_arrive_expect_tx(&bar_tma, bytes);
    for (int c = 0; c < DCH; ++c) {
      tma_load_3d(sQ + c * 16384, &tmQ, &bar_tma, h * D + c * 64, 0, n);
      tma_load_3d(sK + c * p.kv_rows * 128, &tmK, &bar_tma, h * D + c * 64, 0, n);
      tma_load_3d(sV + c * p.kv_rows * 128, &tmV, &bar_tma, h * D + c * 64, 0, n);
    }
  }
  __syncwarp();
  if (tid < 32) tmem_alloc_rt(&tmem_slot, static_cast<uint32_t>(p.tmem_cols));
  pdl_wait();
  for (int w = tid >> 5; w < wpr; w += NT / 32) {
    const int col = w * 32 + (tid & 31);
    const bool on = col < a.Tk && (a.key_on == nullptr || a.key_on[static_cast<long>(n) * a.Tk + col] != 0.0f);
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if ((tid & 31) == 0) sKeyBits[w] = word;
  }
  const int i = t;
  const bool row_ok = i < a.Tq;
  // this thread's two graph words (its 64 key columns of query row i), straight from global memory
  uint32_t gwd[2] = {0xffffffffu, 0xffffffffu};
  if (a.graph_bits && a.renorm != 0) {
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int w = half * kc + cc;
      gwd[cc] = (row_ok && cc < kc && w < wpr) ? __ldg(a.graph_bits + static_cast<long>(n) * a.bits_n_stride + static_cast<long>(i) * a.bits_q_stride + w) : 0u;
    }
  }

  // ---- dO: fp32 [Tq, D] head slice -> bf16 K-major swizzled tile (rows >= Tq are zero), and t = <dO~, O> ----
  {
    constexpr int V4 = D / 4;
    constexpr int kIters = V4 / 2;
    float4 dv[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int idx = tid + it * NT;
      const int row = idx / V4, col = (idx % V4) * 4;
      dv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < a.Tq) {
        const float* src = a.dout + (static_cast<long>(n) * a.Tq + row) * a.ld_dout + h * D + col;
        if (p.dvec) dv[it] = __ldg(reinterpret_cast<const float4*>(src));
        else dv[it] = make_float4(src[0], src[1], src[2], src[3]);
      }
    }
    float part[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int idx = tid + it * NT;
      const int row = idx / V4, col = (idx % V4) * 4;
      part[it] = 0.0f;
      if (row < a.Tq) {
        const float* src = a.out + (static_cast<long>(n) * a.Tq + row) * a.ldo + h * D + col;
        const float4 o = p.ovec ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(src[0], src[1], src[2], src[3]);
        const float2 d01 = unpack_bf16x2(pack_bf16x2(dv[it].x, dv[it].y)), d23 = unpack_bf16x2(pack_bf16x2(dv[it].z, dv[it].w));
        part[it] = (d01.x * o.x + d01.y * o.y) + (d23.x * o.z + d23.y * o.w);
      }
    }
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      float v = part[it];
#pragma unroll
      for (int o = (V4 < 32 ? V4 : 32) / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const int idx = tid + it * NT;
      if ((idx % V4) == 0) sT[idx / V4] = v;
    }
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int idx = tid + it * NT;
      const int row = idx / V4, col = (idx % V4) * 4;
      const int cc = col & 63;
      uint8_t* dst = sdO + (col >> 6) * 16384 + row * 128 + ((((cc >> 3) ^ (row & 7))) << 4) + (cc & 7) * 2;
      *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(dv[it].x, dv[it].y), pack_bf16x2(dv[it].z, dv[it].w));
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    mbar_wait(&bar_tma, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, p.tk_pad16, false, false);
#pragma unroll
    for (int k = 0; k < D / 16; ++k) {
      const int c = k / 4, kk = k % 4;
      umma_bf16_ss(tmem, umma_smem_desc(smem_u32(sQ + c * 16384) + kk * 32, 16, 1024),
                   umma_smem_desc(smem_u32(sK + c * p.kv_rows * 128) + kk * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
    }
#pragma unroll
    for (int k = 0; k < D / 16; ++k) {
      const int c = k / 4, kk = k % 4;
      umma_bf16_ss(tmem + p.dw_off, umma_smem_desc(smem_u32(sdO + c * 16384) + kk * 32, 16, 1024),
                   umma_smem_desc(smem_u32(sV + c * p.kv_rows * 128) + kk * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&bar_s);
  }
  __syncwarp();

  // ---- row pass: thread (t, half) <-> query row t, key columns [64 half, 64 half + 64) ----
  mbar_wait(&bar_s, 0);
  tc_fence_after();
  const long qrow = static_cast<long>(n) * a.Tq + (row_ok ? i : 0);
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const float inv_sqrt_d = 1.0f / sqrtf(static_cast<float>(a.scale_d > 0 ? a.scale_d : D));
  const float qon = (a.query_on && row_ok) ? a.query_on[qrow] : 1.0f;
  const float4 st = __ldg(reinterpret_cast<const float4*>(a.stats) + static_cast<long>(hn) * a.Tq + (row_ok ? i : 0));
  const float m = st.x, inv_z = fabsf(st.y), scale = st.z, beta = st.w, alpha = st.y < 0.0f ? 0.0f : 1.0f;
  const float tsum = sT[t];
  const float okf = row_ok ? 1.0f : 0.0f;
  const float c2 = inv_sqrt_d * kLog2e, m2 = m * kLog2e, marg = (kMaskFill - m) * kLog2e;
  const float sq = scale * qon * okf;                          // W' = ge sq
  const float ca = sq * inv_sqrt_d;                            // dS = ge (dW_raw ca - cb) - cc e, zero on masked keys
  const float cb = scale * alpha * tsum * inv_sqrt_d * okf;
  const float cc_ = beta * inv_z * tsum * inv_sqrt_d * okf;
  uint32_t dsp[2][16];  // this thread's 64 dS values, bf16 pairs: written to the shared tile once dV = W'^T dO has read W' from it
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    if (cc >= kc) break;
    const int c0 = (half * kc + cc) * 32;
    float wq[32];
    if (c0 < a.Tk) {  // warp-uniform
      uint32_t r[32], w[32];
      __syncwarp();
      tmem_ld_32x32(t_lane + c0, r);
      tmem_ld_32x32(t_lane + p.dw_off + c0, w);
      const uint32_t vw = (c0 + 32 <= a.Tk) ? 0xffffffffu : ((1u << (a.Tk - c0)) - 1u);
      const uint32_t kw = sKeyBits[c0 >> 5];
      const uint32_t gw = gwd[cc];
      tmem_ld_wait();
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float arg = fmaf(__uint_as_float(r[j]), c2, -m2);
        e[j] = ex2_approx(((kw >> j) & 1u) ? arg : marg);
      }
      if (vw != 0xffffffffu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) e[j] = ((vw >> j) & 1u) ? e[j] : 0.0f;
      }
      float ds[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float ge = __bfloat162float(__float2bfloat16_rn(((gw >> j) & 1u) ? e[j] : 0.0f));
        const float dsv = fmaf(ge, fmaf(__uint_as_float(w[j]), ca, -cb), -cc_ * e[j]);
        ds[j] = ((kw >> j) & 1u) ? dsv : 0.0f;
        wq[j] = ge * sq;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) dsp[cc][u] = pack_bf16x2(ds[2 * u], ds[2 * u + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) wq[j] = 0.0f;
#pragma unroll
      for (int u = 0; u < 16; ++u) dsp[cc][u] = 0u;
    }
    const int off = (c0 >> 6) * 16384 + t * 128;
    const int u0 = (c0 & 63) >> 3;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      *reinterpret_cast<uint4*>(sX + off + (((u0 + u) ^ (t & 7)) << 4)) =
          make_uint4(pack_bf16x2(wq[8 * u], wq[8 * u + 1]), pack_bf16x2(wq[8 * u + 2], wq[8 * u + 3]),
                     pack_bf16x2(wq[8 * u + 4], wq[8 * u + 5]), pack_bf16x2(wq[8 * u + 6], wq[8 * u + 7]));
  }

  // ---- dV = W'^T dO ----
  const int dq_off = 0, dk_off = D, dv_off = 2 * D;
  const int ksteps_rows = (a.Tq + 15) / 16;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc_k = umma_idesc_bf16(128, 64, true, true);
    for (int c = 0; c < DCH; ++c)
      for (int k = 0; k < ksteps_rows; ++k)
        umma_bf16_ss(tmem + dv_off + c * 64, umma_smem_desc(smem_u32(sX) + k * 2048, 16384, 1024),
                     umma_smem_desc(smem_u32(sdO + c * 16384) + k * 2048, 8192, 1024), idesc_k, k > 0 ? 1u : 0u);
    umma_commit(&bar_v);
  }
  __syncwarp();
  mbar_wait(&bar_v, 0);  // the MMAs have read W': the tile is free for dS
  tc_fence_after();
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    if (cc >= kc) break;
    const int c0 = (half * kc + cc) * 32;
    const int off = (c0 >> 6) * 16384 + t * 128;
    const int u0 = (c0 & 63) >> 3;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      *reinterpret_cast<uint4*>(sX + off + (((u0 + u) ^ (t & 7)) << 4)) =
          make_uint4(dsp[cc][4 * u], dsp[cc][4 * u + 1], dsp[cc][4 * u + 2], dsp[cc][4 * u + 3]);
  }
  // ---- dQ = dS K, dK = dS^T Q ----
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc_q = umma_idesc_bf16(128, 64, false, true);
    const uint32_t idesc_k = umma_idesc_bf16(128, 64, true, true);
    const int ksteps_keys = p.tk_pad16 / 16;
    for (int c = 0; c < DCH; ++c)
      for (int k = 0; k < ksteps_keys; ++k)
        umma_bf16_ss(tmem + dq_off + c * 64, umma_smem_desc(smem_u32(sX + (k >> 2) * 16384) + (k & 3) * 32, 16, 1024),
                     umma_smem_desc(smem_u32(sK + c * p.kv_rows * 128) + k * 2048, 8192, 1024), idesc_q, k > 0 ? 1u : 0u);
    for (int c = 0; c < DCH; ++c)
      for (int k = 0; k < ksteps_rows; ++k)
        umma_bf16_ss(tmem + dk_off + c * 64, umma_smem_desc(smem_u32(sX) + k * 2048, 16384, 1024),
                     umma_smem_desc(smem_u32(sQ + c * 16384) + k * 2048, 8192, 1024), idesc_k, k > 0 ? 1u : 0u);
    umma_commit(&bar_o);
  }
  __syncwarp();

  // ---- epilogue: dV while dQ / dK are still being computed, then dQ and dK (ReLU gates + bf16 stores) ----
  const int j = t;
  const bool key_ok = j < a.Tk;
  const long krow = static_cast<long>(n) * a.Tk + (key_ok ? j : 0);
  {
    const __nv_bfloat16* Vg = static_cast<const __nv_bfloat16*>(a.v) + krow * a.ldv + h * D;
    __nv_bfloat16* dVg = static_cast<__nv_bfloat16*>(a.dv) + krow * a.ld_dv + h * D;
#pragma unroll 1
    for (int c0 = half * 32; c0 < D; c0 += 64) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32(t_lane + dv_off + c0, r);
      tmem_ld_wait();
      store_gated_row32(r, Vg + c0, dVg + c0, key_ok, a.dbv ? a.dbv + h * D + c0 : nullptr, t & 31);
    }
  }
  mbar_wait(&bar_o, 0);
  tc_fence_after();
  {
    const __nv_bfloat16* Qg = static_cast<const __nv_bfloat16*>(a.q) + qrow * a.ldq + h * D;
    __nv_bfloat16* dQg = static_cast<__nv_bfloat16*>(a.dq) + qrow * a.ld_dq + h * D;
    const __nv_bfloat16* Kg = static_cast<const __nv_bfloat16*>(a.k) + krow * a.ldk + h * D;
    __nv_bfloat16* dKg = static_cast<__nv_bfloat16*>(a.dk) + krow * a.ld_dk + h * D;
#pragma unroll 1
    for (int c0 = half * 32; c0 < D; c0 += 64) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32(t_lane + dq_off + c0, r);
      tmem_ld_wait();
      store_gated_row32(r, Qg + c0, dQg + c0, row_ok, a.dbq ? a.dbq + h * D + c0 : nullptr, t & 31);
      __syncwarp();
      tmem_ld_32x32(t_lane + dk_off + c0, r);
      tmem_ld_wait();
      store_gated_row32(r, Kg + c0, dKg + c0, key_ok, a.dbk ? a.dbk + h * D + c0 : nullptr, t & 31);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc_rt(tmem, static_cast<uint32_t>(p.tmem_cols));
  }
}

int make_map3(CUtensorMap* m, const void* base, int64_t ld, int T, int N, int box_rows) {
  const uint64_t dims[3] = {static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(N)};
  const uint64_t str[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * static_cast<uint64_t>(T)};
  const uint32_t box[3] = {64, static_cast<uint32_t>(box_rows), 1};
  return make_tensor_map_bf16(m, base, 3, dims, str, box, true);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

void fill_params(const savqa_attn_args_t* a, BwdParams& p, size_t& smem) {
  p.a = *a;
  p.tk_pad16 = (a->Tk + 15) / 16 * 16;
  p.kt = (a->Tk + 127) / 128;
  p.kc = (p.tk_pad16 + 63) / 64;
  p.kv_rows = p.tk_pad16;
  p.dw_off = (p.tk_pad16 + 31) / 32 * 32;
  const int need = max(p.dw_off + p.tk_pad16, (1 + 2 * p.kt) * a->d);
  int cols = 32;
  while (cols < need) cols *= 2;
  p.tmem_cols = cols;
  p.gvec = (a->graph && a->Tk % 4 == 0 && a->graph_n_stride % 4 == 0 && a->graph_q_stride % 4 == 0 && aligned16(a->graph)) ? 1 : 0;
  p.dvec = (a->ld_dout % 4 == 0 && aligned16(a->dout) && a->d % 4 == 0) ? 1 : 0;
  p.ovec = (a->out && a->ldo % 4 == 0 && aligned16(a->out) && a->d % 4 == 0) ? 1 : 0;
  const int dch = a->d / 64;
  smem = 1024 + static_cast<size_t>(2) * dch * 16384 + static_cast<size_t>(2) * dch * p.kv_rows * 128 +
         static_cast<size_t>(2) * p.kc * 16384 + static_cast<size_t>(a->Tk) * 4 + 16 +
         (a->graph_bits ? static_cast<size_t>(128) * ((a->Tk + 31) / 32) * 4 : 0);
}

}  // namespace

bool attn_bwd_tc_fits(const savqa_attn_args_t* a) {
  if (!(a->d == 64 || a->d == 128)) return false;
  if (a->Tq > 128 || a->Tk > 256 || a->Tq < 1 || a->Tk < 1) return false;
  if (a->ldq % 8 || a->ldk % 8 || a->ldv % 8 || a->ld_dq % 8 || a->ld_dk % 8 || a->ld_dv % 8) return false;
  if (!aligned16(a->q) || !aligned16(a->k) || !aligned16(a->v) || !aligned16(a->dq) || !aligned16(a->dk) || !aligned16(a->dv)) return false;
  BwdParams p;
  size_t smem;
  fill_params(a, p, smem);
  if (p.tmem_cols > 512) return false;
  return smem + 64 <= 227 * 1024;
}

int attn_bwd_tc(const savqa_attn_args_t* a, cudaStream_t stream) {
  SAVQA_REQUIRE(a->q && a->k && a->v && a->dout && a->dq && a->dk && a->dv, "savqa_graph_attn_bwd: null tensor");
  SAVQA_REQUIRE(a->N > 0 && a->H > 0, "savqa_graph_attn_bwd: empty problem");
  SAVQA_REQUIRE(a->renorm >= 0 && a->renorm <= 2, "savqa_graph_attn_bwd: renorm mode %d", a->renorm);
  SAVQA_REQUIRE(attn_bwd_tc_fits(a),
                "savqa_graph_attn_bwd: the tcgen05 engine takes d in {64,128}, Tq <= 128, Tk <= 256 and 16-byte aligned rows "
                "(got d=%d Tq=%d Tk=%d); use engine 1",
                a->d, a->Tq, a->Tk);
  BwdParams p;
  size_t smem;
  fill_params(a, p, smem);
  alignas(64) CUtensorMap tmQ, tmK, tmV;
  if (int rc = make_map3(&tmQ, a->q, a->ldq, a->Tq, a->N, 128)) return rc;
  if (int rc = make_map3(&tmK, a->k, a->ldk, a->Tk, a->N, p.kv_rows)) return rc;
  if (int rc = make_map3(&tmV, a->v, a->ldv, a->Tk, a->N, p.kv_rows)) return rc;
  dim3 grid(a->N * a->H);
  // the training step's symbolic branch: two CTAs per SM with the shared W' / dS tile (attn_bwd_tc1_kernel)
  const bool one_tile = a->d == 64 && p.kc <= 2 && p.kt == 1 && a->Tk > 32 && a->stats && a->out && !a->causal &&
                        (a->graph_bits || !a->graph || a->renorm == 0) && p.tmem_cols <= 256 && getenv("SAVQA_ATTN_BWD_ONE_TILE_OFF") == nullptr;
  count_launch(one_tile ? LK_ATTN_BWD_TC_SHARED : LK_ATTN_BWD_TC);
  if (one_tile) {
    const size_t smem1 = 1024 + static_cast<size_t>(p.kc) * 16384 + static_cast<size_t>(2) * 16384 + static_cast<size_t>(2) * p.kv_rows * 128;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_bwd_tc1_kernel<64>), smem1, "savqa_graph_attn_bwd (tcgen05 engine, shared tile)"))
      return rc;
    SAVQA_CHECK_CUDA(launch_kernel(true, attn_bwd_tc1_kernel<64>, grid, dim3(256), smem1, stream, tmQ, tmK, tmV, p));
    SAVQA_CHECK_CUDA(cudaGetLastError());
    return SAVQA_OK;
  }
  // two CTAs fit an SM when the key tile is a single 64-key chunk (d = 64); otherwise one CTA with two threads per row
  const bool split = smem > 112 * 1024;
#define SAVQA_LAUNCH_BWD(DD, RR)                                                                                                           \
  do {                                                                                                                                     \
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_bwd_tc_kernel<DD, RR>), smem, "savqa_graph_attn_bwd (tcgen05 engine)")) \
      return rc;                                                                                                                           \
    SAVQA_CHECK_CUDA(launch_kernel(true, attn_bwd_tc_kernel<DD, RR>, grid, dim3(128 * RR), smem, stream, tmQ, tmK, tmV, p));               \
  } while (0)
  if (a->d == 64) {
    if (split) SAVQA_LAUNCH_BWD(64, 2);
    else SAVQA_LAUNCH_BWD(64, 1);
  } else {
    if (split) SAVQA_LAUNCH_BWD(128, 2);
    else SAVQA_LAUNCH_BWD(128, 1);
  }
#undef SAVQA_LAUNCH_BWD
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

}  // namespace savqa
