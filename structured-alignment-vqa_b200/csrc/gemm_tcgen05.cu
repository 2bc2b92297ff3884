// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32
// accumulators in TMEM, double buffered) -> fused epilogue straight from TMEM.
//
//   acc[m,n] = sum_k A[m,k] * B[n,k]          (nn.Linear of modules.py:227-229, 428-429; AttModel_x3.py:42-44)
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (each owns the 32-lane TMEM quadrant `warp_idx & 3`).
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), static persistent tile loop.
//
// Operand majors: K-major (k contiguous, the forward / dgrad case) or MN-major (m or n contiguous: the wgrad
// case dW = dY^T X, where both operands are read "transposed" straight from their row-major activations).
#include <stdlib.h>

#include "common.cuh"

namespace savqa {

int gemm2_launch(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int M, int N, int K,
                 const savqa_gemm_epilogue_t* epi, int split_k, cudaStream_t stream, bool* handled);  // gemm2_tcgen05.cu

int gemm2_launch_group(const savqa_gemm_problem_t* probs, int count, int a_mn, int b_mn, int N, int split_k, cudaStream_t stream,
                       bool* handled);  // gemm2_tcgen05.cu

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;

struct GemmParams {
  int M, N, K;
  int num_m, num_n, split_k, kb_per_split, num_kb;
  int vec_ok;
  savqa_gemm_epilogue_t e;
};

template <int BN>
struct Cfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;
  static constexpr int kTmemCols = 2 * BN;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C::kStages];
  __shared__ __align__(8) uint64_t empty_bar[C::kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(&tmem_base_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = tmem_base_slot;

  const int tiles_mn = p.num_m * p.num_n;
  const int num_tiles = tiles_mn * p.split_k;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.num_n;
        const int m_blk = (tile / p.num_n) % p.num_m;
        const int ks = tile / tiles_mn;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
          if constexpr (!A_MN) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, &full_bar[stage], m_blk * BM + c * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * 8192, &tmB, &full_bar[stage], n_blk * BN + c * 64, kb * BK);
          }
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int ks = tile / tiles_mn;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const int as = it & 1;
        mbar_wait(&tmem_empty_bar[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + C::kABytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = A_MN ? umma_smem_desc(sa + k * (UMMA_K * 128), 8192, 1024) : umma_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc(sb + k * (UMMA_K * 128), 8192, 1024) : umma_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
            umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[as]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const savqa_gemm_epilogue_t& e = p.e;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int n_blk = tile % p.num_n;
      const int m_blk = (tile / p.num_n) % p.num_m;
      const int as = it & 1;
      mbar_wait(&tmem_full_bar[as], (it >> 1) & 1);
      tc_fence_after();
      const long grow = static_cast<long>(m_blk) * BM + row_in_tile;
      const bool row_ok = grow < p.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * BN);
      const float* rowtab_row = (e.rowtab != nullptr && row_ok) ? e.rowtab + static_cast<long>(grow % e.rowtab_period) * e.ld_rowtab : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged here
        tmem_ld_32x32(t_row + c0, r);
        tmem_ld_wait();
        const int gcol = n_blk * BN + c0;
        float v[32];
        if (row_ok && gcol < p.N) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * e.alpha;
        const bool full = p.vec_ok && (gcol + 32 <= p.N);
        if (full) {
          if (e.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(e.bias + gcol);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          if (e.res) {
            const float4* r4 = reinterpret_cast<const float4*>(e.res + grow * e.ld_res + gcol);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(r4 + j);
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          if (rowtab_row) {
            const float4* r4 = reinterpret_cast<const float4*>(rowtab_row + gcol);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(r4 + j);
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          if (e.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
          }
          if (e.gate_bf16) {
            const uint4* g4 = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(e.gate_bf16) + grow * e.ld_gate + gcol);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
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
          if (e.out_f32) {
            float* o = e.out_f32 + grow * e.ld_out_f32 + gcol;
            if (e.accumulate == 2) {
#pragma unroll
              for (int j = 0; j < 8; ++j) red_add_v4(o + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else if (e.accumulate == 1) {
              float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 t = o4[j];
                t.x += v[4 * j]; t.y += v[4 * j + 1]; t.z += v[4 * j + 2]; t.w += v[4 * j + 3];
                o4[j] = t;
              }
            } else {
              float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
              for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
          if (e.out_bf16) {
            uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out_bf16) + grow * e.ld_out_bf16 + gcol);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o4[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
        } else {
          // ragged / unaligned edge: scalar, fully guarded
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = gcol + j;
            if (col >= p.N) {
              v[j] = 0.0f;
              continue;
            }
            float x = v[j];
            if (e.bias) x += e.bias[col];
            if (e.res) x += e.res[grow * e.ld_res + col];
            if (rowtab_row) x += rowtab_row[col];
            if (e.relu) x = fmaxf(x, 0.0f);
            if (e.gate_bf16) {
              const float g = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(e.gate_bf16)[grow * e.ld_gate + col]);
              if (!(g > 0.0f)) x = 0.0f;
            }
            if (e.out_f32) {
              float* o = e.out_f32 + grow * e.ld_out_f32 + col;
              if (e.accumulate == 2) atomicAdd(o, x);
              else if (e.accumulate == 1) *o += x;
              else *o = x;
            }
            if (e.out_bf16) reinterpret_cast<__nv_bfloat16*>(e.out_bf16)[grow * e.ld_out_bf16 + col] = __float2bfloat16_rn(x);
            v[j] = x;
          }
        }
        }  // row_ok
        if (e.colsum && gcol < p.N) {  // warp-uniform: bias gradient = column sums of the tile's valid rows
          if (!row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0f;
          }
          const float cs = warp_colsum32(v, lane);
          if (gcol + lane < p.N) atomicAdd(e.colsum + gcol + lane, cs);
        }
      }
      // all TMEM reads of this accumulator buffer are done -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int BN, bool A_MN, bool B_MN>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::kSmemBytes, "savqa_gemm_bf16")) return rc;
  const int tiles = p.num_m * p.num_n * p.split_k;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  count_launch(LK_GEMM_SINGLE);
  SAVQA_CHECK_CUDA(launch_kernel(true, kern, dim3(grid), dim3(kThreads), C::kSmemBytes, stream, tmA, tmB, p));
  return SAVQA_OK;
}

template <int BN>
int launch_major(bool a_mn, bool b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<BN, false, false>(tmA, tmB, p, s);
  if (a_mn && b_mn) return launch<BN, true, true>(tmA, tmB, p, s);
  if (!a_mn && b_mn) return launch<BN, false, true>(tmA, tmB, p, s);
  return launch<BN, true, false>(tmA, tmB, p, s);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// SAVQA_GEMM_ENGINE=1 pins the single-CTA kernel (A/B comparison of the two kernels in tests and microbenchmarks)
int engine_pref() {
  static int pref = -1;
  if (pref < 0) {
    const char* s = getenv("SAVQA_GEMM_ENGINE");
    pref = (s && s[0] == '1') ? 1 : 0;
  }
  return pref;
}

}  // namespace

}  // namespace savqa

extern "C" int savqa_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, int M, int N,
                               int K, const savqa_gemm_epilogue_t* epi, int split_k, savqa_stream_t stream_) {
  using namespace savqa;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SAVQA_REQUIRE(A && B && epi, "savqa_gemm_bf16: null operand");
  SAVQA_REQUIRE(M > 0 && N > 0 && K > 0, "savqa_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  SAVQA_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "savqa_gemm_bf16: lda=%lld ldb=%lld must be multiples of 8 (16-byte TMA pitch)",
                static_cast<long long>(lda), static_cast<long long>(ldb));
  SAVQA_REQUIRE(epi->out_f32 || epi->out_bf16, "savqa_gemm_bf16: no output");
  SAVQA_REQUIRE(split_k >= 1, "savqa_gemm_bf16: split_k must be >= 1");
  SAVQA_REQUIRE(split_k == 1 || (epi->accumulate == 2 && !epi->out_bf16 && !epi->relu && !epi->gate_bf16 && !epi->bias && !epi->res && !epi->rowtab),
                "savqa_gemm_bf16: split_k > 1 needs accumulate == 2 and a linear epilogue");
  SAVQA_REQUIRE(!epi->rowtab || epi->rowtab_period > 0, "savqa_gemm_bf16: rowtab needs a period");
  SAVQA_REQUIRE(!epi->colsum || split_k == 1, "savqa_gemm_bf16: colsum needs split_k == 1");

  // large problems: the CTA-pair kernel (cta_group::2, TMA-store epilogue); everything else: the single-CTA kernel below
  if (engine_pref() != 1) {
    bool handled = false;
    if (int rc = gemm2_launch(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, epi, split_k, stream, &handled)) return rc;
    if (handled) return SAVQA_OK;
  }

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.e = *epi;
  p.num_kb = (K + BK - 1) / BK;
  if (split_k > p.num_kb) split_k = p.num_kb;
  p.kb_per_split = (p.num_kb + split_k - 1) / split_k;
  p.split_k = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;

  // tile width: wide tiles amortise the A traffic, narrow ones fill the 148 SMs on small problems
  int BN = 128;
  const long tiles128 = static_cast<long>((M + BM - 1) / BM) * ((N + 127) / 128) * p.split_k;
  if (N <= 64) BN = 64;
  else if (N % 256 == 0 && tiles128 >= 4L * sm_count()) BN = 256;
  p.num_m = (M + BM - 1) / BM;
  p.num_n = (N + BN - 1) / BN;

  bool vec = true;
  if (epi->bias && !aligned16(epi->bias)) vec = false;
  if (epi->res && (!aligned16(epi->res) || epi->ld_res % 4)) vec = false;
  if (epi->rowtab && (!aligned16(epi->rowtab) || epi->ld_rowtab % 4)) vec = false;
  if (epi->gate_bf16 && (!aligned16(epi->gate_bf16) || epi->ld_gate % 8)) vec = false;
  if (epi->out_f32 && (!aligned16(epi->out_f32) || epi->ld_out_f32 % 4)) vec = false;
  if (epi->out_bf16 && (!aligned16(epi->out_bf16) || epi->ld_out_bf16 % 8)) vec = false;
  p.vec_ok = vec ? 1 : 0;

  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn_major) {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t str[1] = {static_cast<uint64_t>(lda) * 2};
    const uint32_t box[2] = {BK, BM};
    rc = make_tensor_map_bf16(&tmA, A, 2, dims, str, box, true);
  } else {
    const uint64_t dims[2] = {static_cast<uint64_t>(M), static_cast<uint64_t>(K)};
    const uint64_t str[1] = {static_cast<uint64_t>(lda) * 2};
    const uint32_t box[2] = {64, BK};
    rc = make_tensor_map_bf16(&tmA, A, 2, dims, str, box, true);
  }
  if (rc) return rc;
  if (!b_mn_major) {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t str[1] = {static_cast<uint64_t>(ldb) * 2};
    const uint32_t box[2] = {BK, static_cast<uint32_t>(BN)};
    rc = make_tensor_map_bf16(&tmB, B, 2, dims, str, box, true);
  } else {
    const uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(K)};
    const uint64_t str[1] = {static_cast<uint64_t>(ldb) * 2};
    const uint32_t box[2] = {64, BK};
    rc = make_tensor_map_bf16(&tmB, B, 2, dims, str, box, true);
  }
  if (rc) return rc;

  switch (BN) {
    case 64: return launch_major<64>(a_mn_major != 0, b_mn_major != 0, tmA, tmB, p, stream);
    case 256: return launch_major<256>(a_mn_major != 0, b_mn_major != 0, tmA, tmB, p, stream);
    default: return launch_major<128>(a_mn_major != 0, b_mn_major != 0, tmA, tmB, p, stream);
  }
}

namespace savqa { int gemm2_set_sm_limit(int sms); }
extern "C" int savqa_set_gemm_sm_limit(int sms) { return savqa::gemm2_set_sm_limit(sms); }

extern "C" int savqa_gemm_bf16_grouped(const savqa_gemm_problem_t* problems, int count, int a_mn_major, int b_mn_major, int N, int split_k,
                                       savqa_stream_t stream_) {
  using namespace savqa;
  SAVQA_REQUIRE(problems && count >= 1, "savqa_gemm_bf16_grouped: no problems");
  bool handled = false;
  if (count <= 2 && engine_pref() != 1) {
    bool ok = true;
    for (int i = 0; i < count; ++i) ok = ok && problems[i].A && problems[i].B && problems[i].M > 0 && problems[i].K > 0 &&
                                         problems[i].lda % 8 == 0 && problems[i].ldb % 8 == 0 && (!problems[i].epilogue.colsum || split_k == 1) &&
                                         (problems[i].epilogue.out_f32 || problems[i].epilogue.out_bf16);
    if (ok)
      if (int rc = gemm2_launch_group(problems, count, a_mn_major, b_mn_major, N, split_k, static_cast<cudaStream_t>(stream_), &handled)) return rc;
  }
  if (handled) return SAVQA_OK;
  for (int i = 0; i < count; ++i)
    if (int rc = savqa_gemm_bf16(problems[i].A, problems[i].lda, a_mn_major, problems[i].B, problems[i].ldb, b_mn_major, problems[i].M, N,
                                 problems[i].K, &problems[i].epilogue, split_k, stream_))
      return rc;
  return SAVQA_OK;
}
