// a5: fused graph-weighted attention forward on the 5th-gen tensor cores (engine 0).
//
// One CTA per (sample n, head h, 128-query tile).  TMA stages the head's Q / K / V tiles (3-D tensor maps over
// the [N, T, ld] bf16 projections, zero fill past the sample's last row) into 128B-swizzled shared memory;
//   S = Q K^T          tcgen05.mma, A = Q (K-major), B = K (K-major), fp32 accumulator in TMEM columns [0, Tk)
//   softmax pass       thread t owns query row t (TMEM lane t): scale, key mask, causal mask, row max, exp, the
//                      graph weight G, the three row sums; un-normalised G*e goes to smem as the bf16 A operand
//   O = (G*e) V        tcgen05.mma, A = P (K-major, from smem), B = V (MN-major: V is [Tk, d] with d contiguous),
//                      accumulator re-uses TMEM columns [0, d) once every thread has consumed S
//   epilogue           O * (renormalisation scale * query mask) -> fp32 [N*Tq, C] head slice
// The probabilities never touch HBM unless the caller asks for `att` (return_att=True).
// Semantics: modules.py:246-301 (see attn_simt.cu for the scalar restatement used to cross-check this kernel).
#include "common.cuh"

namespace savqa {

int attn_fwd_simt(const savqa_attn_args_t* a, cudaStream_t stream);
int attn_bwd_simt(const savqa_attn_args_t* a, cudaStream_t stream);
int attn_bwd_tc(const savqa_attn_args_t* a, cudaStream_t stream);

namespace {

constexpr float kMaskFill = -4294967296.0f;

struct AttnTcParams {
  savqa_attn_args_t a;
  int tk_pad16;   // Tk rounded up to 16 (MMA N / K granularity)
  int tk_chunks;  // number of 64-column chunks of the P tile
  int kv_rows;    // rows of the K / V smem tiles (tk_pad16 rounded up to the TMA boxes)
  int kv_box;     // rows per K / V TMA box (<= 256)
  int tmem_cols;  // power of two >= max(tk_pad16, d), >= 32
  int qk_bytes;   // max(Q + K tiles, P tile): the region the P tile shares with Q and K
  int gvec;       // graph rows can be read with float4
  int ovec;       // out rows can be written with float4
};

__device__ __forceinline__ void tmem_alloc_rt(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_rt(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D = head size (64 or 128)
template <int D>
__global__ void __launch_bounds__(128) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                         const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  constexpr int DCH = D / 64;  // 64-column chunks of the head dimension
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_tma, bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t sKeyBits[16];  // bit j of word w: key 32 w + j exists and is switched on (Tk <= 512)
  pdl_trigger();
  const savqa_attn_args_t& a = p.a;
  const int t = threadIdx.x, warp = t >> 5;
  // CTA order: the H heads of one sample are neighbours, so that the CTAs in flight together read the SAME rows of the fused
  // [q | k | v] projection (3 KB per token, 128 B of it per head and operand): one DRAM page serves all of them
  const int n = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int hn = h * a.N + n;  // the reference's head-major batch index (layout of `att` and of the row statistics)
  const int q0 = blockIdx.y * 128;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                   // DCH chunks of [128][128 B]
  uint8_t* sK = sQ + DCH * 16384;                       // DCH chunks of [kv_rows][128 B]
  uint8_t* sP = smem;                                   // tk_chunks chunks of [128][128 B]: OVER Q and K, which are dead once
                                                        // S = Q K^T has completed (every thread waits for that before writing P);
                                                        // the smaller footprint is what lets 4 CTAs share an SM and hide each
                                                        // other's TMA / TMEM / exp latencies
  uint8_t* sV = smem + p.qk_bytes;                      // DCH chunks of [kv_rows][128 B]
  float* sKeyOn = reinterpret_cast<float*>(sV + DCH * p.kv_rows * 128);  // [Tk]
  uint32_t* sBits = reinterpret_cast<uint32_t*>(sKeyOn + ((a.Tk + 3) & ~3));  // [128][wpr] bit-packed graph rows of this query tile
  const int wpr = (a.Tk + 31) >> 5;

  if (t == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_rt(&tmem_slot, static_cast<uint32_t>(p.tmem_cols));
  pdl_wait();
  for (int j = t; j < a.Tk; j += 128) sKeyOn[j] = a.key_on ? a.key_on[static_cast<long>(n) * a.Tk + j] : 1.0f;
  for (int w = warp; w < wpr; w += 4) {
    const int col = w * 32 + (t & 31);
    const bool on = col < a.Tk && (a.key_on == nullptr || a.key_on[static_cast<long>(n) * a.Tk + col] != 0.0f);
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if ((t & 31) == 0) sKeyBits[w] = word;
  }
  if (a.graph_bits) {
    for (int idx = t; idx < 128 * wpr; idx += 128) {
      const int row = idx / wpr, w = idx % wpr;
      sBits[idx] = (q0 + row < a.Tq) ? __ldg(a.graph_bits + static_cast<long>(n) * a.bits_n_stride + static_cast<long>(q0 + row) * a.bits_q_stride + w) : 0u;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (t == 0) {
    const int boxes = p.kv_rows / p.kv_box;
    const uint32_t bytes = static_cast<uint32_t>(DCH) * (16384u + 2u * static_cast<uint32_t>(p.kv_rows) * 128u);
    mbar_arrive_expect_tx(&bar_tma, bytes);
    for (int c = 0; c < DCH; ++c) {
      tma_load_3d(sQ + c * 16384, &tmQ, &bar_tma, h * D + c * 64, q0, n);
      for (int b = 0; b < boxes; ++b) {
        tma_load_3d(sK + (c * p.kv_rows + b * p.kv_box) * 128, &tmK, &bar_tma, h * D + c * 64, b * p.kv_box, n);
        tma_load_3d(sV + (c * p.kv_rows + b * p.kv_box) * 128, &tmV, &bar_tma, h * D + c * 64, b * p.kv_box, n);
      }
    }
    // ---- S = Q K^T ----
    mbar_wait(&bar_tma, 0);
    tc_fence_after();
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
    }
    umma_commit(&bar_s);
  }
  __syncwarp();

  // ---- softmax: thread t <-> query row q0 + t <-> TMEM lane t ----
  mbar_wait(&bar_s, 0);
  tc_fence_after();
  const int i = q0 + t;
  const bool row_ok = i < a.Tq;
  const long qrow = static_cast<long>(n) * a.Tq + (row_ok ? i : 0);
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const float inv_sqrt_d = 1.0f / sqrtf(static_cast<float>(a.scale_d > 0 ? a.scale_d : D));  // exact for D = 64
  const int renorm = (a.graph || a.graph_bits) ? a.renorm : 0;
  const float* grow = (a.graph && row_ok) ? a.graph + static_cast<long>(n) * a.graph_n_stride + static_cast<long>(i) * a.graph_q_stride : nullptr;

  float m = -INFINITY;
  float Z = 0.0f, R = 0.0f, SA = 0.0f;
  // Fast row pass (what the training step runs): no causal mask and a 0/1 graph that came bit-packed (or no graph).  The key
  // mask, the validity of the last chunk and the graph are three 32-bit words per 32-column chunk; e = 2^(raw c2 - m2), the
  // same expression in attn_bwd_tcgen05.cu.  ~12 instructions per score instead of ~50 (profiles/r1_06_attn_fwd_lines.txt).
  const bool fast = !a.causal && (a.graph_bits != nullptr || renorm == 0);
  if (fast) {
    float mraw = -INFINITY;
    uint32_t any_off = 0;
    for (int c0 = 0; c0 < a.Tk; c0 += 32) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32(t_lane + c0, r);
      const uint32_t vw = (c0 + 32 <= a.Tk) ? 0xffffffffu : ((1u << (a.Tk - c0)) - 1u);
      const uint32_t kw = sKeyBits[c0 >> 5];
      any_off |= vw & ~kw;
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mraw = fmaxf(mraw, ((kw >> j) & 1u) ? __uint_as_float(r[j]) : -INFINITY);
    }
    m = mraw * inv_sqrt_d;                      // max and the positive scale commute exactly
    if (any_off) m = fmaxf(m, kMaskFill);       // a masked key takes part in the row maximum with the fill value (modules.py:269-274)
    const float c2 = inv_sqrt_d * kLog2e, m2 = m * kLog2e;
    const float marg = (kMaskFill - m) * kLog2e;  // exponent of a masked key: 0 when every key is masked (uniform row), else e = 0
    for (int c0 = 0; c0 < p.tk_chunks * 64; c0 += 32) {
      float ge[32];
      if (c0 < a.Tk) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld_32x32(t_lane + c0, r);
        const uint32_t vw = (c0 + 32 <= a.Tk) ? 0xffffffffu : ((1u << (a.Tk - c0)) - 1u);
        const uint32_t kw = sKeyBits[c0 >> 5];
        const uint32_t gw = renorm != 0 ? sBits[t * wpr + (c0 >> 5)] : 0xffffffffu;
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
          const float w = ((gw >> j) & 1u) ? e[j] : 0.0f;
          Z += e[j];
          SA += w;
          ge[j] = w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) ge[j] = 0.0f;
      }
      uint8_t* prow = sP + (c0 >> 6) * 16384 + t * 128;
      const int u0 = (c0 & 63) >> 3;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint4 v = make_uint4(pack_bf16x2(ge[8 * u], ge[8 * u + 1]), pack_bf16x2(ge[8 * u + 2], ge[8 * u + 3]),
                                   pack_bf16x2(ge[8 * u + 4], ge[8 * u + 5]), pack_bf16x2(ge[8 * u + 6], ge[8 * u + 7]));
        *reinterpret_cast<uint4*>(prow + (((u0 + u) ^ (t & 7)) << 4)) = v;
      }
    }
    R = SA;  // 0/1 graph: |G e| = G e
  } else {
  for (int c0 = 0; c0 < a.Tk; c0 += 32) {
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
  for (int c0 = 0; c0 < p.tk_chunks * 64; c0 += 32) {
    uint32_t r[32];
    float ge[32];
    if (c0 < a.Tk) {
      __syncwarp();
      tmem_ld_32x32(t_lane + c0, r);
      tmem_ld_wait();
      float g[32];
      if (a.graph_bits && renorm != 0) {
        const uint32_t word = sBits[t * wpr + (c0 >> 5)];
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] = ((word >> j) & 1u) ? 1.0f : 0.0f;
      } else if (grow && renorm != 0) {
        if (p.gvec && c0 + 32 <= a.Tk) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(grow + c0) + j);
            g[4 * j] = v.x; g[4 * j + 1] = v.y; g[4 * j + 2] = v.z; g[4 * j + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) g[j] = (c0 + j < a.Tk) ? __ldg(grow + c0 + j) : 0.0f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] = 1.0f;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = c0 + j;
        float e = 0.0f, w = 0.0f;
        if (col < a.Tk) {
          float s = __uint_as_float(r[j]) * inv_sqrt_d;
          if (sKeyOn[col] == 0.0f) s = kMaskFill;
          if (a.causal && col > i) s = kMaskFill;
          e = ex2_approx(s == kMaskFill ? (kMaskFill - m) * kLog2e : fmaf(__uint_as_float(r[j]), inv_sqrt_d * kLog2e, -m * kLog2e));
          w = g[j] * e;
        }
        Z += e;
        R += fabsf(w);
        SA += w;
        ge[j] = w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) ge[j] = 0.0f;
    }
    // bf16 A operand, K-major, 128B swizzle: (row t, col) -> chunk col/64, 16-byte unit ((col%64)/8) ^ (t%8)
    uint8_t* prow = sP + (c0 >> 6) * 16384 + t * 128;
    const int u0 = (c0 & 63) >> 3;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint4 v = make_uint4(pack_bf16x2(ge[8 * u], ge[8 * u + 1]), pack_bf16x2(ge[8 * u + 2], ge[8 * u + 3]),
                                 pack_bf16x2(ge[8 * u + 4], ge[8 * u + 5]), pack_bf16x2(ge[8 * u + 6], ge[8 * u + 7]));
      *reinterpret_cast<uint4*>(prow + (((u0 + u) ^ (t & 7)) << 4)) = v;
    }
  }
  }  // generic row pass
  float scale;
  bool clamped = false;
  if (renorm == 1) {
    clamped = !(R / Z >= 1e-12f);
    scale = clamped ? 1.0f / (Z * 1e-12f) : 1.0f / R;
  } else if (renorm == 2) {
    scale = 1.0f / (SA + 1e-7f * Z);
  } else {
    scale = 1.0f / Z;
  }
  if (a.stats && row_ok) {  // for the backward kernel: dS_j = W_j (dW_j - alpha t) - beta (e_j / Z) t
    const float beta = clamped ? 1.0f : (renorm == 2 ? 1.0f - scale * SA : 0.0f);
    const float inv_z = 1.0f / Z;
    reinterpret_cast<float4*>(a.stats)[static_cast<long>(hn) * a.Tq + i] = make_float4(m, clamped ? -inv_z : inv_z, scale, beta);
  }

  // return_att: recompute W = G*e*scale in fp32 and stream it out (pre query mask); whole-warp TMEM loads
  if (a.att) {
    float* arow = a.att + (static_cast<long>(hn) * a.Tq + (row_ok ? i : 0)) * a.Tk;
    for (int c0 = 0; c0 < a.Tk; c0 += 32) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32(t_lane + c0, r);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c0 + j;
          if (col < a.Tk) {
            float s = __uint_as_float(r[j]) * inv_sqrt_d;
            if (sKeyOn[col] == 0.0f) s = kMaskFill;
            if (a.causal && col > i) s = kMaskFill;
            const float g = (renorm == 0) ? 1.0f : (a.graph_bits ? (((sBits[t * wpr + (col >> 5)] >> (col & 31)) & 1u) ? 1.0f : 0.0f) : (grow ? __ldg(grow + col) : 1.0f));
            arow[col] = g * ex2_approx(s == kMaskFill ? (kMaskFill - m) * kLog2e : fmaf(__uint_as_float(r[j]), inv_sqrt_d * kLog2e, -m * kLog2e)) * scale;
          }
        }
      }
    }
  }

  // ---- O = P V ----
  fence_proxy_async_smem();  // generic-proxy smem writes of P -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  if (t == 0) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 64, false, true);
    for (int c = 0; c < DCH; ++c) {
      for (int k = 0; k < p.tk_pad16 / 16; ++k) {
        const uint64_t adesc = umma_smem_desc(smem_u32(sP + (k >> 2) * 16384) + (k & 3) * 32, 16, 1024);
        const uint64_t bdesc = umma_smem_desc(smem_u32(sV + c * p.kv_rows * 128) + k * 2048, 8192, 1024);
        umma_bf16_ss(tmem + c * 64, adesc, bdesc, idesc, k > 0 ? 1u : 0u);
      }
    }
    umma_commit(&bar_o);
  }
  __syncwarp();
  mbar_wait(&bar_o, 0);
  tc_fence_after();
  const float qon = (a.query_on && row_ok) ? a.query_on[qrow] : 1.0f;
  const float oscale = scale * qon;
  float* orow = a.out + qrow * a.ldo + h * D;
#pragma unroll 1
  for (int c0 = 0; c0 < D; c0 += 32) {
    uint32_t r[32];
    __syncwarp();
    tmem_ld_32x32(t_lane + c0, r);
    tmem_ld_wait();
    if (row_ok) {
      if (p.ovec) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          reinterpret_cast<float4*>(orow + c0)[j] = make_float4(__uint_as_float(r[4 * j]) * oscale, __uint_as_float(r[4 * j + 1]) * oscale,
                                                                __uint_as_float(r[4 * j + 2]) * oscale, __uint_as_float(r[4 * j + 3]) * oscale);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) orow[c0 + j] = __uint_as_float(r[j]) * oscale;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
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

int attn_fwd_tc(const savqa_attn_args_t* a, cudaStream_t stream) {
  AttnTcParams p;
  p.a = *a;
  p.tk_pad16 = (a->Tk + 15) / 16 * 16;
  p.tk_chunks = (a->Tk + 63) / 64;
  p.kv_box = p.tk_pad16 <= 256 ? p.tk_pad16 : 256;
  p.kv_rows = (p.tk_pad16 + p.kv_box - 1) / p.kv_box * p.kv_box;
  int cols = 32;
  while (cols < p.tk_pad16 || cols < a->d) cols *= 2;
  p.tmem_cols = cols;
  p.gvec = (a->graph && a->Tk % 4 == 0 && a->graph_n_stride % 4 == 0 && a->graph_q_stride % 4 == 0 &&
            (reinterpret_cast<uintptr_t>(a->graph) & 15) == 0)
               ? 1
               : 0;
  p.ovec = (a->ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0) ? 1 : 0;
  const int dch = a->d / 64;
  const int qk = dch * 16384 + dch * p.kv_rows * 128;
  p.qk_bytes = qk > p.tk_chunks * 16384 ? qk : p.tk_chunks * 16384;
  const size_t smem = 1024 + static_cast<size_t>(p.qk_bytes) + static_cast<size_t>(dch) * p.kv_rows * 128 +
                      static_cast<size_t>(a->Tk) * 4 + 16 + (a->graph_bits ? static_cast<size_t>(128) * ((a->Tk + 31) / 32) * 4 : 0);
  alignas(64) CUtensorMap tmQ, tmK, tmV;
  if (int rc = make_map3(&tmQ, a->q, a->ldq, a->Tq, a->N, 128)) return rc;
  if (int rc = make_map3(&tmK, a->k, a->ldk, a->Tk, a->N, p.kv_box)) return rc;
  if (int rc = make_map3(&tmV, a->v, a->ldv, a->Tk, a->N, p.kv_box)) return rc;
  dim3 grid(a->N * a->H, (a->Tq + 127) / 128);
  count_launch(LK_ATTN_FWD_TC);
  if (a->d == 64) {
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_fwd_tc_kernel<64>), smem, "savqa_graph_attn_fwd (tcgen05 engine)")) return rc;
    SAVQA_CHECK_CUDA(launch_kernel(true, attn_fwd_tc_kernel<64>, grid, dim3(128), smem, stream, tmQ, tmK, tmV, p));
  } else {
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_fwd_tc_kernel<128>), smem, "savqa_graph_attn_fwd (tcgen05 engine)")) return rc;
    SAVQA_CHECK_CUDA(launch_kernel(true, attn_fwd_tc_kernel<128>, grid, dim3(128), smem, stream, tmQ, tmK, tmV, p));
  }
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

}  // namespace
}  // namespace savqa

using namespace savqa;

extern "C" int savqa_graph_attn_fwd(const savqa_attn_args_t* a, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SAVQA_REQUIRE(a, "savqa_graph_attn_fwd: null args");
  if (a->engine == 1) return attn_fwd_simt(a, stream);
  SAVQA_REQUIRE(a->engine == 0, "savqa_graph_attn_fwd: unknown engine %d", a->engine);
  SAVQA_REQUIRE(a->q && a->k && a->v && a->out, "savqa_graph_attn_fwd: null tensor");
  SAVQA_REQUIRE(a->N > 0 && a->H > 0 && a->Tq > 0 && a->Tk > 0, "savqa_graph_attn_fwd: empty problem");
  SAVQA_REQUIRE(a->d == 64 || a->d == 128, "savqa_graph_attn_fwd: the tcgen05 engine takes head size 64 or 128 (got %d); use engine 1",
                a->d);
  SAVQA_REQUIRE(a->Tk <= 512, "savqa_graph_attn_fwd: Tk=%d exceeds the single-pass TMEM limit 512", a->Tk);
  SAVQA_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0, "savqa_graph_attn_fwd: leading dimensions must be multiples of 8");
  SAVQA_REQUIRE(a->renorm >= 0 && a->renorm <= 2, "savqa_graph_attn_fwd: renorm mode %d", a->renorm);
  return attn_fwd_tc(a, stream);
}

extern "C" int savqa_graph_attn_bwd(const savqa_attn_args_t* a, savqa_stream_t stream_) {
  SAVQA_REQUIRE(a, "savqa_graph_attn_bwd: null args");
  if (a->engine == 1) return attn_bwd_simt(a, static_cast<cudaStream_t>(stream_));
  SAVQA_REQUIRE(a->engine == 0, "savqa_graph_attn_bwd: unknown engine %d", a->engine);
  return attn_bwd_tc(a, static_cast<cudaStream_t>(stream_));
}
