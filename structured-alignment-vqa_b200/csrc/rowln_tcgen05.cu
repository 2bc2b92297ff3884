// Cluster GEMM with a ROW-WISE epilogue for the decoder's chain of M = B-row layers (AttModel_x3.py:141-154): one thread-block
// CLUSTER computes a [128 x N] output block, CTA r of the cluster owns the 64-column slab [64 r, 64 r + 64) -- so the N = C = 512
// GEMMs of the chain run on 8 SMs instead of 4, each streaming an eighth of the weight matrix -- and the reductions over a whole
// output row that the reference's layer_normalization needs (modules.py:62-65) are exchanged between the CTAs through
// DISTRIBUTED SHARED MEMORY: every CTA writes its per-row partial sums into every peer's shared memory, one cluster barrier, and
// each CTA merges the partials in rank order (identical arithmetic in all CTAs).  What used to be GEMM -> elementwise -> LayerNorm
// (three dependent launches of the launch-latency-bound chain, ~25 us) is one launch.
//
//   mode 0 (plain):   v = [gate](relu?(acc + bias)) + res                               -> out bf16 / fp32        (Q projection, conv1, ReLU-gated dgrad)
//   mode 1 (LN fwd):  a = relu?(acc + bias) [-> act_bf16, rounded to bf16]; pre = a * rowscale + res;
//                     y = gamma (pre - mean) / (sigma + eps) + beta, sigma unbiased         -> pre, y, y_bf16, on, stats = {mean, sigma}
//                     (one-token self-attention: LN(relu(x Wv^T + bv) * query_mask + x), modules.py:119-207 with one key;
//                      feedforward conv2: LN(h W2^T + b2 + x), modules.py:439-447)
//   mode 2 (LN bwd):  dy = acc + res;  dx = LN'(pre)[dy] (formula of csrc/layernorm.cu)    -> dx, dx_bf16, dxg_bf16 = gate(dx * rowscale),
//                     dgamma += sum_r dy c / s, dbeta += sum_r dy, dxsum += sum_r dx        (a dgrad GEMM that feeds a LayerNorm backward)
//
// Per CTA: warp 0 = TMA producer, warp 1 = TMEM allocator + tcgen05.mma issuer (M = 128, N = 64), warps 2..9 = epilogue (one
// thread per output row and 32-column half of the slab, its accumulator columns in registers).  A [128 x K] is streamed in 64-column k-blocks by every CTA
// (L2-resident: the previous link wrote it), B is the CTA's own [64 x K] slab of the weights.
#include "common.cuh"

namespace savqa {
namespace {

constexpr int BM = 128;
constexpr int BN = 64;   // columns per CTA
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int EC = 32;   // columns per epilogue thread (two epilogue warps per TMEM lane quadrant)
constexpr int kThreads = 320;
constexpr int kStages = 8;
constexpr int kMaxCluster = 8;
constexpr int kABytes = BM * BK * 2;
constexpr int kBBytes = BN * BK * 2;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kMaxSlab = 2 * kMaxCluster;             // 32-column partials per output row
constexpr int kPartBytes = 2 * kMaxSlab * BM * 8;     // two exchange rounds x slabs x rows x {float, float}
constexpr int kSmemBytes = kStages * kStageBytes + kPartBytes + 1024;

struct RowLnParams {
  int M, N, K, num_kb, cluster;
  savqa_rowln_args_t a;
};

__device__ __forceinline__ void st_cluster_f32x2(uint32_t cluster_addr, float x, float y) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(x), "f"(y) : "memory");
}

// every CTA of the cluster receives this thread's pair at part[round][slab][row]
__device__ __forceinline__ void share_pair(float2* part, int round, int slab, int cluster, int row, float x, float y) {
  const uint32_t local = smem_u32(part + (round * kMaxSlab + slab) * BM + row);
  for (int r = 0; r < cluster; ++r) st_cluster_f32x2(mapa_shared(local, static_cast<uint32_t>(r)), x, y);
}

// this thread's pair -> part[round][slab][row] of every CTA of the cluster (or of this CTA alone when nothing is exchanged)
__device__ __forceinline__ void put_pair(float2* part, bool exchange, int round, int slab, int cluster, int row, float x, float y) {
  if (exchange) share_pair(part, round, slab, cluster, row, x, y);
  else part[(round * kMaxSlab + slab) * BM + row] = make_float2(x, y);
}

template <bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
rowln_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const RowLnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float2* part = reinterpret_cast<float2*>(smem + kStages * kStageBytes);
  const savqa_rowln_args_t& a = p.a;
  const int n0 = blockIdx.x * BN;    // this CTA's column slab
  const int m_blk = blockIdx.y;
  const bool exchange = a.mode != 0 && p.cluster > 1;
  const uint32_t my_rank = p.cluster > 1 ? cluster_ctarank() : 0u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(&tmem_base_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (exchange) cluster_sync_all();  // every CTA of the cluster is running: its shared memory may be written from now on
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kStageBytes;
        uint8_t* sb = sa + kABytes;
        mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
        tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
        if constexpr (!B_MN) tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
        else tma_load_2d(sb, &tmB, &full_bar[stage], n0, kb * BK);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, false, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t adesc = umma_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
          const uint64_t bdesc = B_MN ? umma_smem_desc(sb + k * (UMMA_K * 128), 8192, 1024) : umma_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
          umma_bf16_ss(tmem_base, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(&tmem_full_bar);
    }
  }

  __syncwarp();  // lanes 1..31 of the two single-thread role warps wait for their lane 0: the barriers below are warp-aligned

  // ===================== epilogue (warps 2..9) interleaved with the cluster barriers every warp takes =====================
  // Two warps per TMEM lane quadrant, each thread one output row and EC = 32 of the CTA's 64 columns: the row epilogue is
  // straight-line code that every thread runs once (instruction fetch is its first stall reason, profiles/r2_ncu_rowln.txt) --
  // half the columns per thread is half the code, and the two halves run side by side.  A row's statistics are merged from
  // 2 * cluster partials of 32 columns each (slab index 2 * rank + half), in slab order, by every thread that needs them.
  const bool epi = warp >= 2;
  const int quad = warp & 3;
  const int half = epi ? ((warp - 2) >> 2) : 0;
  const int row = quad * 32 + lane;           // row of the 128-row block this thread owns (epilogue warps)
  const long grow = static_cast<long>(m_blk) * BM + row;
  const bool row_ok = epi && grow < p.M;
  const int nc0 = n0 + EC * half;             // first global column of this thread
  const int slab = 2 * static_cast<int>(my_rank) + half;
  const int nslab = 2 * p.cluster;
  float v[EC];
  float c_[EC];                                // mode 2: pre - mean
  float mean = 0.0f, sigma = 0.0f, inv = 0.0f;
  if (epi) {
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    {
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32(t_row + EC * half, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    if (a.mode != 2) {
      // ---- forward value: relu?(acc + bias), optional bf16 activation output, gate, row scale, residual ----
      if (a.bias) {
#pragma unroll
        for (int j = 0; j < EC; j += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + nc0 + j));
          v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
      }
      if (a.relu) {
#pragma unroll
        for (int j = 0; j < EC; ++j) v[j] = fmaxf(v[j], 0.0f);
      }
      if (a.gate_bf16 && row_ok) {
        const uint4* g4 = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(a.gate_bf16) + grow * a.ld_gate + nc0);
#pragma unroll
        for (int j = 0; j < EC / 8; ++j) {
          const uint4 g = __ldg(g4 + j);
          const uint32_t w[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = unpack_bf16x2(w[q]);
            if (!(f.x > 0.0f)) v[8 * j + 2 * q] = 0.0f;
            if (!(f.y > 0.0f)) v[8 * j + 2 * q + 1] = 0.0f;
          }
        }
      }
      if (a.act_bf16) {
        // the bf16 activation the next GEMM / the backward reads; the fp32 path continues from its ROUNDED value (what the unfused
        // path does: relu output staged in bf16, then multiplied by the query mask, functional.TokenSelfAttentionFn)
        uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a.act_bf16) + grow * a.ld_act + nc0);
#pragma unroll
        for (int j = 0; j < EC / 8; ++j) {
          const uint4 pk = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                      pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          if (row_ok) o4[j] = pk;
          const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = unpack_bf16x2(w[q]);
            v[8 * j + 2 * q] = f.x;
            v[8 * j + 2 * q + 1] = f.y;
          }
        }
      }
      if (a.mode == 1) {
        if (a.rowscale) {
          const float rs = row_ok ? __ldg(a.rowscale + grow) : 0.0f;
#pragma unroll
          for (int j = 0; j < EC; ++j) v[j] *= rs;
        }
        if (a.res && row_ok) {
#pragma unroll
          for (int j = 0; j < EC; j += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(a.res + grow * a.ld_res + nc0 + j));
            v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
          }
        }
        if (!row_ok) {
#pragma unroll
          for (int j = 0; j < EC; ++j) v[j] = 0.0f;
        }
        if (a.pre && row_ok) {
#pragma unroll
          for (int j = 0; j < EC; j += 4)
            *reinterpret_cast<float4*>(a.pre + grow * a.ld_pre + nc0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        // slab statistics: (sum, M2 about the slab mean) -> exact pairwise merge over the cluster (Chan et al.)
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < EC; ++j) s += v[j];
        const float ml = s * (1.0f / EC);
        float m2 = 0.0f;
#pragma unroll
        for (int j = 0; j < EC; ++j) m2 = fmaf(v[j] - ml, v[j] - ml, m2);
        put_pair(part, exchange, 0, slab, p.cluster, row, s, m2);
      } else {
        // ---- plain epilogue: (+ residual) store ----
        if (a.res && row_ok) {
#pragma unroll
          for (int j = 0; j < EC; j += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(a.res + grow * a.ld_res + nc0 + j));
            v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
          }
        }
        if (row_ok) {
          if (a.y) {
#pragma unroll
            for (int j = 0; j < EC; j += 4)
              *reinterpret_cast<float4*>(a.y + grow * a.ld_y + nc0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (a.y_bf16) {
            uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a.y_bf16) + grow * a.ld_yb + nc0);
#pragma unroll
            for (int j = 0; j < EC / 8; ++j)
              o4[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
        }
      }
    } else {
      // ---- mode 2: dy = acc + res; g = dy * gamma; c = pre - mean (statistics saved by the forward) ----
      if (a.res && row_ok) {
#pragma unroll
        for (int j = 0; j < EC; j += 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(a.res + grow * a.ld_res + nc0 + j));
          v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
        }
      }
      if (row_ok) {
        mean = __ldg(a.stats + 2 * grow);
        sigma = __ldg(a.stats + 2 * grow + 1);
      }
      inv = 1.0f / (sigma + a.eps);
      float sg = 0.0f, dot = 0.0f;
#pragma unroll
      for (int j = 0; j < EC; j += 4) {
        float4 pr = make_float4(mean, mean, mean, mean);
        if (row_ok) pr = __ldg(reinterpret_cast<const float4*>(a.pre + grow * a.ld_pre + nc0 + j));
        const float4 gm = __ldg(reinterpret_cast<const float4*>(a.gamma + nc0 + j));
        c_[j] = pr.x - mean; c_[j + 1] = pr.y - mean; c_[j + 2] = pr.z - mean; c_[j + 3] = pr.w - mean;
        if (!row_ok) { v[j] = 0.0f; v[j + 1] = 0.0f; v[j + 2] = 0.0f; v[j + 3] = 0.0f; }
        const float g0 = v[j] * gm.x, g1 = v[j + 1] * gm.y, g2 = v[j + 2] * gm.z, g3 = v[j + 3] * gm.w;
        sg += (g0 + g1) + (g2 + g3);
        dot = fmaf(g0, c_[j], fmaf(g1, c_[j + 1], fmaf(g2, c_[j + 2], fmaf(g3, c_[j + 3], dot))));
      }
      put_pair(part, exchange, 0, slab, p.cluster, row, sg, dot);
    }
  }
  if (a.mode == 0) {
    // nothing crosses CTAs (or warps) in plain mode
  } else {
    __syncwarp();
    if (exchange) cluster_sync_all(); else __syncthreads();
    float ysum = 0.0f;
    if (epi) {
      if (a.mode == 1) {
        // merge the slabs' (sum, M2) in slab order: mean, unbiased sigma of the whole row
        float tot = 0.0f;
        for (int r = 0; r < nslab; ++r) tot += part[r * BM + row].x;
        mean = tot / static_cast<float>(p.N);
        float m2 = 0.0f;
        for (int r = 0; r < nslab; ++r) {
          const float2 q = part[r * BM + row];
          const float d = q.x * (1.0f / EC) - mean;
          m2 += q.y + static_cast<float>(EC) * d * d;
        }
        sigma = sqrtf(m2 / static_cast<float>(p.N - 1));
        inv = 1.0f / (sigma + a.eps);
#pragma unroll
        for (int j = 0; j < EC; j += 4) {
          const float4 gm = __ldg(reinterpret_cast<const float4*>(a.gamma + nc0 + j));
          const float4 bt = __ldg(reinterpret_cast<const float4*>(a.beta + nc0 + j));
          v[j] = gm.x * (v[j] - mean) * inv + bt.x;
          v[j + 1] = gm.y * (v[j + 1] - mean) * inv + bt.y;
          v[j + 2] = gm.z * (v[j + 2] - mean) * inv + bt.z;
          v[j + 3] = gm.w * (v[j + 3] - mean) * inv + bt.w;
          ysum += (v[j] + v[j + 1]) + (v[j + 2] + v[j + 3]);
        }
        if (row_ok) {
          if (a.y) {
#pragma unroll
            for (int j = 0; j < EC; j += 4)
              *reinterpret_cast<float4*>(a.y + grow * a.ld_y + nc0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (a.y_bf16) {
            uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a.y_bf16) + grow * a.ld_yb + nc0);
#pragma unroll
            for (int j = 0; j < EC / 8; ++j)
              o4[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
          if (a.stats && slab == 0) {
            a.stats[2 * grow] = mean;
            a.stats[2 * grow + 1] = sigma;
          }
        }
        if (a.on) put_pair(part, exchange, 1, slab, p.cluster, row, ysum, 0.0f);
      } else {
        // mode 2: totals of (sum g, sum g c) over the row, then dx and the column sums of this warp's 32 rows x 32 columns
        float sg = 0.0f, dot = 0.0f;
        for (int r = 0; r < nslab; ++r) {
          const float2 q = part[r * BM + row];
          sg += q.x;
          dot += q.y;
        }
        const float mg = sg / static_cast<float>(p.N);
        const float sden = sigma + a.eps;
        const float k2 = (sigma > 0.0f) ? dot / (static_cast<float>(p.N - 1) * sigma * sden * sden) : 0.0f;
        const float rs = (a.rowscale && row_ok) ? __ldg(a.rowscale + grow) : 1.0f;
        {
          // column sums over this warp's 32 rows, one quantity at a time (one 32-register scratch array)
          float t32[32];
          if (a.dbeta) {
#pragma unroll
            for (int j = 0; j < 32; ++j) t32[j] = v[j];
            const float t = warp_colsum32(t32, lane);
            atomicAdd(a.dbeta + nc0 + lane, t);
          }
          if (a.dgamma) {
#pragma unroll
            for (int j = 0; j < 32; ++j) t32[j] = v[j] * c_[j] * inv;
            const float t = warp_colsum32(t32, lane);
            atomicAdd(a.dgamma + nc0 + lane, t);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 gm = __ldg(reinterpret_cast<const float4*>(a.gamma + nc0 + j));
            const float gmv[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int jj = j + q;
              v[jj] = row_ok ? (v[jj] * gmv[q] - mg) * inv - c_[jj] * k2 : 0.0f;
            }
          }
          if (a.dxsum) {
#pragma unroll
            for (int j = 0; j < 32; ++j) t32[j] = v[j];
            const float t = warp_colsum32(t32, lane);
            atomicAdd(a.dxsum + nc0 + lane, t);
          }
        }
        if (row_ok) {
          if (a.y) {
#pragma unroll
            for (int j = 0; j < EC; j += 4)
              *reinterpret_cast<float4*>(a.y + grow * a.ld_y + nc0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (a.y_bf16) {
            uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a.y_bf16) + grow * a.ld_yb + nc0);
#pragma unroll
            for (int j = 0; j < EC / 8; ++j)
              o4[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
          if (a.dxg_bf16) {
            // ReLU backward of the layer in front of this LayerNorm, staged as the next GEMM's bf16 operand: (act > 0) ? dx * rowscale : 0
            const uint4* g4 = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(a.gate_bf16) + grow * a.ld_gate + nc0);
            uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a.dxg_bf16) + grow * a.ld_dxg + nc0);
#pragma unroll
            for (int j = 0; j < EC / 8; ++j) {
              const uint4 g = __ldg(g4 + j);
              const uint32_t w[4] = {g.x, g.y, g.z, g.w};
              float t[8];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 f = unpack_bf16x2(w[q]);
                t[2 * q] = f.x > 0.0f ? v[8 * j + 2 * q] * rs : 0.0f;
                t[2 * q + 1] = f.y > 0.0f ? v[8 * j + 2 * q + 1] * rs : 0.0f;
              }
              o4[j] = make_uint4(pack_bf16x2(t[0], t[1]), pack_bf16x2(t[2], t[3]), pack_bf16x2(t[4], t[5]), pack_bf16x2(t[6], t[7]));
            }
          }
        }
      }
    }
    if (a.mode == 1 && a.on) {
      __syncwarp();
      if (exchange) cluster_sync_all(); else __syncthreads();
      if (epi && row_ok && slab == 0) {
        float tot = 0.0f;
        for (int r = 0; r < nslab; ++r) tot += part[(kMaxSlab + r) * BM + row].x;
        a.on[grow] = (tot != 0.0f) ? 1.0f : 0.0f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc<BN>(tmem_base);
  }
}

template <bool B_MN>
int launch_rowln(const CUtensorMap& tmA, const CUtensorMap& tmB, const RowLnParams& p, cudaStream_t stream) {
  auto kern = rowln_gemm_kernel<B_MN>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), kSmemBytes, "savqa_gemm_rowln")) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.N / BN, (p.M + BM - 1) / BM);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(p.cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.cluster > 1 ? 1 : 0;
  count_launch(LK_ROWLN_GEMM);
  SAVQA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p));
  return SAVQA_OK;
}

bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

}  // namespace
}  // namespace savqa

using namespace savqa;

extern "C" int savqa_gemm_rowln(const savqa_rowln_args_t* a, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SAVQA_REQUIRE(a && a->A && a->B, "savqa_gemm_rowln: null operand");
  SAVQA_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->N % BN == 0, "savqa_gemm_rowln: bad shape M=%d N=%d K=%d (N must be a multiple of 64)", a->M,
                a->N, a->K);
  SAVQA_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0 && al16(a->A) && al16(a->B), "savqa_gemm_rowln: operands need 16-byte aligned rows");
  SAVQA_REQUIRE(a->mode >= 0 && a->mode <= 2, "savqa_gemm_rowln: mode %d", a->mode);
  RowLnParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.num_kb = (a->K + BK - 1) / BK;
  p.a = *a;
  const int slabs = a->N / BN;
  if (a->mode == 0) {
    p.cluster = 1;
    SAVQA_REQUIRE(a->y || a->y_bf16, "savqa_gemm_rowln: no output");
  } else {
    SAVQA_REQUIRE(slabs <= kMaxCluster && (slabs & (slabs - 1)) == 0, "savqa_gemm_rowln: a row-wise epilogue needs N in {64,128,256,512} (got %d)", a->N);
    p.cluster = slabs;
    SAVQA_REQUIRE(a->gamma && al16(a->gamma), "savqa_gemm_rowln: gamma");
    if (a->mode == 1) SAVQA_REQUIRE(a->beta && al16(a->beta) && (a->y || a->y_bf16), "savqa_gemm_rowln: LayerNorm forward needs beta and an output");
    if (a->mode == 2) SAVQA_REQUIRE(a->pre && a->stats && a->ld_pre % 4 == 0 && al16(a->pre), "savqa_gemm_rowln: LayerNorm backward needs pre and stats");
    if (a->mode == 2) SAVQA_REQUIRE(!a->dxg_bf16 || (a->gate_bf16 && a->ld_dxg % 8 == 0 && al16(a->dxg_bf16)), "savqa_gemm_rowln: dxg needs the gate activation");
  }
  SAVQA_REQUIRE(!a->bias || al16(a->bias), "savqa_gemm_rowln: bias alignment");
  SAVQA_REQUIRE(!a->res || (al16(a->res) && a->ld_res % 4 == 0), "savqa_gemm_rowln: res alignment");
  SAVQA_REQUIRE(!a->gate_bf16 || (al16(a->gate_bf16) && a->ld_gate % 8 == 0), "savqa_gemm_rowln: gate alignment");
  SAVQA_REQUIRE(!a->act_bf16 || (al16(a->act_bf16) && a->ld_act % 8 == 0), "savqa_gemm_rowln: act alignment");
  SAVQA_REQUIRE(!a->pre || (al16(a->pre) && a->ld_pre % 4 == 0), "savqa_gemm_rowln: pre alignment");
  SAVQA_REQUIRE(!a->y || (al16(a->y) && a->ld_y % 4 == 0), "savqa_gemm_rowln: y alignment");
  SAVQA_REQUIRE(!a->y_bf16 || (al16(a->y_bf16) && a->ld_yb % 8 == 0), "savqa_gemm_rowln: y_bf16 alignment");
  alignas(64) CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->M)};
    const uint64_t str[1] = {static_cast<uint64_t>(a->lda) * 2};
    const uint32_t box[2] = {BK, BM};
    if (int rc = make_tensor_map_bf16(&tmA, a->A, 2, dims, str, box, true)) return rc;
  }
  if (!a->b_mn_major) {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->N)};
    const uint64_t str[1] = {static_cast<uint64_t>(a->ldb) * 2};
    const uint32_t box[2] = {BK, BN};
    if (int rc = make_tensor_map_bf16(&tmB, a->B, 2, dims, str, box, true)) return rc;
    return launch_rowln<false>(tmA, tmB, p, stream);
  }
  const uint64_t dims[2] = {static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->K)};
  const uint64_t str[1] = {static_cast<uint64_t>(a->ldb) * 2};
  const uint32_t box[2] = {BN, BK};
  if (int rc = make_tensor_map_bf16(&tmB, a->B, 2, dims, str, box, true)) return rc;
  return launch_rowln<true>(tmA, tmB, p, stream);
}
