// HBM-bound integer / byte / streaming kernels of the path: scene-graph mask construction, embedding gathers and
// their scatter-add backward, bf16 staging casts, activation-derived padding masks, bias-gradient column sums.
// All are coalesced, vectorised where the alignment allows, and sized in multiples of the SM count.
#include "common.cuh"

namespace savqa {
namespace {

// ------------------------------------------------------------------------------------------------------
// a4: masks.  One thread per output element of the [B,T,T] planes (two fp32 stores per thread, fully coalesced).
// Reference: AttModel_x3.py:103-122 / :229-247.
// ------------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void build_masks_kernel(const TIn* __restrict__ first_mask, const TIn* __restrict__ q_mask, const TIn* __restrict__ q_graph,
                                   const TIn* __restrict__ first_graph, int B, int V, int Q, int dec_mask_on,
                                   float* __restrict__ graph_diag, float* __restrict__ graph, float* __restrict__ dec_mask) {
  const int T = V + Q;
  const long total = static_cast<long>(B) * T * T;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % T);
    const int r = static_cast<int>((i / T) % T);
    const long b = i / (static_cast<long>(T) * T);
    float gd = 0.0f, g = 1.0f;  // off-diagonal blocks: 1 - block_diag(..) = 1
    if (r >= V && c >= V) {
      const long qi = (b * Q + (r - V)) * Q + (c - V);
      gd = static_cast<float>(q_mask[qi]);
      g = static_cast<float>(q_graph[qi]);
    } else if (r < V && c < V) {
      g = first_graph ? static_cast<float>(first_graph[(b * V + r) * V + c]) : 1.0f;
    }
    graph_diag[i] = gd;
    graph[i] = g;
  }
  // dec_mask[b,0,j] = (sum_k mask[b,j,k] != 0); the sum is taken in fp32 like the reference (exact for 0/1 inputs in any
  // order).  One warp per mask row: coalesced reads, one shuffle reduction.
  const long rows = static_cast<long>(B) * T;
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  for (long i = warp0; i < rows; i += nwarps) {
    const int j = static_cast<int>(i % T);
    const long b = i / T;
    float on = 0.0f;
    if (dec_mask_on) {
      const TIn* row = (j < V) ? first_mask + (b * V + j) * V : q_mask + (b * Q + (j - V)) * Q;
      const int len = (j < V) ? V : Q;
      float s = 0.0f;
      for (int k = lane; k < len; k += 32) s += static_cast<float>(row[k]);
      s = warp_sum(s);
      on = (s != 0.0f) ? 1.0f : 0.0f;
    }
    if (lane == 0) dec_mask[i] = on;
  }
}

// One warp per (row, word): a coalesced 128-byte read, one ballot, one 4-byte store.
__global__ void pack_graph_bits_kernel(const float* __restrict__ graph, long rows, int Tk, uint32_t* __restrict__ bits, int wpr) {
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  const long total = rows * wpr;
  for (long i = warp0; i < total; i += nwarps) {
    const long r = i / wpr;
    const int col = static_cast<int>(i % wpr) * 32 + lane;
    const bool on = col < Tk && graph[r * Tk + col] != 0.0f;
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if (lane == 0) bits[i] = word;
  }
}

// ------------------------------------------------------------------------------------------------------
// a1/a2: row gather.  One warp per output row; 16-byte loads when width % 4 == 0 (300 = 75 float4).
// ------------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float* __restrict__ table, long table_rows, int width, const int64_t* __restrict__ idx,
                                   long n_idx, float scale, float* __restrict__ out_f32, long ld_f32,
                                   __nv_bfloat16* __restrict__ out_bf16, long ld_bf16, int pad_to) {
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  const bool vec = (width % 4 == 0) && ((reinterpret_cast<uintptr_t>(table) & 15) == 0) &&
                   (!out_f32 || (ld_f32 % 4 == 0 && (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0)) &&
                   (!out_bf16 || (ld_bf16 % 4 == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0));
  for (long r = warp0; r < n_idx; r += nwarps) {
    const long src = idx[r];
    const bool ok = src >= 0 && src < table_rows;
    const float* trow = table + (ok ? src : 0) * static_cast<long>(width);
    if (vec) {
      for (int c = lane * 4; c < width; c += 128) {
        float4 v = ok ? __ldg(reinterpret_cast<const float4*>(trow + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (scale != 1.0f) { v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale; }
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + r * ld_f32 + c) = v;
        if (out_bf16) *reinterpret_cast<uint2*>(out_bf16 + r * ld_bf16 + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      }
    } else {
      for (int c = lane; c < width; c += 32) {
        float v = ok ? __ldg(trow + c) : 0.0f;
        if (scale != 1.0f) v *= scale;
        if (out_f32) out_f32[r * ld_f32 + c] = v;
        if (out_bf16) out_bf16[r * ld_bf16 + c] = __float2bfloat16_rn(v);
      }
    }
    if (out_bf16)
      for (int c = width + lane; c < pad_to; c += 32) out_bf16[r * ld_bf16 + c] = __float2bfloat16_rn(0.0f);
  }
}

__global__ void scatter_add_rows_kernel(float* __restrict__ dtable, long table_rows, int width, const int64_t* __restrict__ idx,
                                        long n_idx, const float* __restrict__ dout, long ld_dout, float scale, long skip_row) {
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  for (long r = warp0; r < n_idx; r += nwarps) {
    const long dst = idx[r];
    if (dst < 0 || dst >= table_rows || dst == skip_row) continue;
    float* trow = dtable + dst * static_cast<long>(width);
    const float* g = dout + r * ld_dout;
    for (int c = lane; c < width; c += 32) atomicAdd(trow + c, g[c] * scale);
  }
}

// Row gradients of a data-parallel step: every rank accumulates the SAME gathered (row id, gradient row) lists, and the replicas must
// stay bit-identical although the order of the atomics is not. Integer addition is associative, float addition is not: the accumulator
// is a signed Q15.48 fixed-point table (resolution 2^-48 = 3.6e-15, range +-32768; a float scaled by a power of two is exact, so the
// only rounding is the one to 2^-48). savqa_adam_rows reads it back (grad_q48 = 1).
constexpr float kQ48 = 281474976710656.0f;          // 2^48
constexpr float kQ48Inv = 1.0f / 281474976710656.0f;

__global__ void scatter_add_rows_q48_kernel(long long* __restrict__ acc, long table_rows, int width, const int64_t* __restrict__ idx,
                                            long n_idx, const float* __restrict__ dout, long ld_dout, float scale, long skip_row) {
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  for (long r = warp0; r < n_idx; r += nwarps) {
    const long dst = idx[r];
    if (dst < 0 || dst >= table_rows || dst == skip_row) continue;
    unsigned long long* trow = reinterpret_cast<unsigned long long*>(acc) + dst * static_cast<long>(width);
    const float* g = dout + r * ld_dout;
    for (int c = lane; c < width; c += 32) {
      const long long q = __float2ll_rn(g[c] * scale * kQ48);  // saturates; two's complement wrap-around makes the unsigned add signed
      if (q != 0) atomicAdd(trow + c, static_cast<unsigned long long>(q));
    }
  }
}

__device__ __forceinline__ float load_row_grad(const float* g, long i) { return g[i]; }
__device__ __forceinline__ float load_row_grad(const long long* g, long i) { return __ll2float_rn(g[i]) * kQ48Inv; }
__device__ __forceinline__ void clear_row_grad(float* g, long i) { g[i] = 0.0f; }
__device__ __forceinline__ void clear_row_grad(long long* g, long i) { g[i] = 0; }

// ------------------------------------------------------------------------------------------------------
// staging casts
// ------------------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ src, long ld_src, __nv_bfloat16* __restrict__ dst, long ld_dst, long rows,
                                 int cols, int pad_to) {
  const long per_row = (pad_to + 3) / 4;
  const long total = rows * per_row;
  const bool vec = (ld_src % 4 == 0) && (ld_dst % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / per_row;
    const int c = static_cast<int>(i % per_row) * 4;
    if (vec && c + 4 <= cols) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * ld_src + c));
      *reinterpret_cast<uint2*>(dst + r * ld_dst + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    } else {
      for (int j = 0; j < 4 && c + j < pad_to; ++j)
        dst[r * ld_dst + c + j] = __float2bfloat16_rn(c + j < cols ? src[r * ld_src + c + j] : 0.0f);
    }
  }
}

// dst[c, r] = src[r, c]; 32x32 smem tile, coalesced on both sides.
__global__ void cast_transpose_bf16_kernel(const float* __restrict__ src, long ld_src, __nv_bfloat16* __restrict__ dst, long ld_dst,
                                           long rows, int cols, int pad_to) {
  __shared__ float tile[32][33];
  const long r0 = static_cast<long>(blockIdx.y) * 32;
  const int c0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long r = r0 + j;
    const int c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? src[r * ld_src + c] : 0.0f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;             // dst row
    const long r = r0 + threadIdx.x;  // dst col
    if (c < cols && r < pad_to) dst[static_cast<long>(c) * ld_dst + r] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}

// out[r, :] = (act[r, :] > 0) ? dy[src(r), :] : 0 in bf16, with src(r) = (r / group_rows) * group_stride + r % group_rows: the rows of
// dy may sit in groups inside a larger tensor (the node rows of every sample inside d[B, T, 2048]) -- no contiguous copy first.
// Eight columns per thread (16-byte accesses) when the pointers / pitches allow it.
template <typename TDy>
__global__ void relu_gate_kernel(const TDy* __restrict__ dy, long ld_dy, const __nv_bfloat16* __restrict__ act, long ld_act,
                                 __nv_bfloat16* __restrict__ out, long ld_out, long rows, int cols, long group_rows, long group_stride,
                                 int vec) {
  if (vec) {
    const int cv = cols >> 3;
    const long total = rows * cv;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
      const long r = i / cv;
      const int c = static_cast<int>(i - r * cv) << 3;
      const long sr = (r / group_rows) * group_stride + (r % group_rows);
      const uint4 a4 = __ldg(reinterpret_cast<const uint4*>(act + r * ld_act + c));
      float g[8];
      if constexpr (sizeof(TDy) == 4) {
        const float4 lo = __ldcs(reinterpret_cast<const float4*>(dy + sr * ld_dy + c));
        const float4 hi = __ldcs(reinterpret_cast<const float4*>(dy + sr * ld_dy + c) + 1);
        g[0] = lo.x; g[1] = lo.y; g[2] = lo.z; g[3] = lo.w; g[4] = hi.x; g[5] = hi.y; g[6] = hi.z; g[7] = hi.w;
      } else {
        const uint4 d4 = __ldcs(reinterpret_cast<const uint4*>(dy + sr * ld_dy + c));
        const uint32_t w[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16x2(w[q]);
          g[2 * q] = f.x;
          g[2 * q + 1] = f.y;
        }
      }
      const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w};
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack_bf16x2(aw[q]);
        o[q] = pack_bf16x2(f.x > 0.0f ? g[2 * q] : 0.0f, f.y > 0.0f ? g[2 * q + 1] : 0.0f);
      }
      *reinterpret_cast<uint4*>(out + r * ld_out + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    return;
  }
  const long total = rows * cols;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / cols;
    const int c = static_cast<int>(i % cols);
    const long sr = (r / group_rows) * group_stride + (r % group_rows);
    const float g = static_cast<float>(dy[sr * ld_dy + c]);
    const float a = __bfloat162float(act[r * ld_act + c]);
    out[r * ld_out + c] = __float2bfloat16_rn(a > 0.0f ? g : 0.0f);
  }
}

// 16 bytes per thread (the group widths and pitches are multiples of 16 bytes): a vector either lies inside a group's copied part
// or inside its zero part.
__global__ void regroup_cols_vec_kernel(const uint4* __restrict__ in, long ld_in16, uint4* __restrict__ out, long ld_out16, long rows, int groups,
                                        int w_in16, int w_out16) {
  const long per_row = static_cast<long>(groups) * w_out16;
  const long total = rows * per_row;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / per_row;
    const int rem = static_cast<int>(i - r * per_row);
    const int gidx = rem / w_out16, c = rem - gidx * w_out16;
    out[r * ld_out16 + rem] = c < w_in16 ? __ldg(in + r * ld_in16 + gidx * w_in16 + c) : make_uint4(0u, 0u, 0u, 0u);
  }
}

template <typename T>
__global__ void regroup_cols_kernel(const T* __restrict__ in, long ld_in, T* __restrict__ out, long ld_out, long rows, int groups, int w_in,
                                    int w_out) {
  const long per_row = static_cast<long>(groups) * w_out;
  const long total = rows * per_row;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / per_row;
    const int rem = static_cast<int>(i - r * per_row);
    const int gidx = rem / w_out, c = rem - gidx * w_out;
    out[r * ld_out + rem] = c < w_in ? in[r * ld_in + gidx * w_in + c] : T(0);
  }
}

// Zero fill with a BOUNDED grid: a kernel of tens of thousands of blocks keeps the block scheduler from dispatching the kernels
// other streams launch behind it until its last wave (seen in the step trace: the 356 MB gradient zero-fill "next to" the forward
// pass delayed it by its full 50 us); a few persistent blocks trickle through HBM underneath them instead.
__global__ void __launch_bounds__(256) fill_zero_kernel(uint4* __restrict__ p, long n16, uint8_t* __restrict__ tail, long n_tail) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n16; i += static_cast<long>(gridDim.x) * blockDim.x) p[i] = z;
  if (blockIdx.x == 0)
    for (long i = threadIdx.x; i < n_tail; i += blockDim.x) tail[i] = 0;
}

// One warp per row: fp32 row sum (the reference's sum(x,-1)), on = (sum != 0); optional bf16 copy of the row.
__global__ void row_nonzero_kernel(const float* __restrict__ x, long ld, long rows, int cols, float* __restrict__ on,
                                   __nv_bfloat16* __restrict__ x_bf16, long ld_bf16) {
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  const bool vec = (cols % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                   (!x_bf16 || (ld_bf16 % 4 == 0 && (reinterpret_cast<uintptr_t>(x_bf16) & 7) == 0));
  for (long r = warp0; r < rows; r += nwarps) {
    const float* row = x + r * ld;
    float s = 0.0f;
    if (vec) {
      for (int c = lane * 4; c < cols; c += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row + c));
        s += (v.x + v.y) + (v.z + v.w);
        if (x_bf16) *reinterpret_cast<uint2*>(x_bf16 + r * ld_bf16 + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      }
    } else {
      for (int c = lane; c < cols; c += 32) {
        const float v = row[c];
        s += v;
        if (x_bf16) x_bf16[r * ld_bf16 + c] = __float2bfloat16_rn(v);
      }
    }
    s = warp_sum(s);
    if (lane == 0 && on) on[r] = (s != 0.0f) ? 1.0f : 0.0f;
  }
}

// out[c] += sum_r x[r,c].  Block = 32 x 8 threads over a 64-column strip (bf16x2 per thread), rows strided over
// blockIdx.y; one atomicAdd per column per block.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long ld, long rows, int cols, float* __restrict__ out,
                                                          int vec) {
  // block = 8 warps over one 256-column strip: lane l owns columns [8 l, 8 l + 8) (one 16-byte load per row), every warp strides
  // over the rows with four loads in flight; the eight warps' partial sums meet in shared memory, one atomic per column per block
  __shared__ float part[8][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  long r = static_cast<long>(blockIdx.y) * 8 + w;
  const long stride = static_cast<long>(gridDim.y) * 8;
  if (vec && c + 8 <= cols) {
    for (; r + 3 * stride < rows; r += 4 * stride) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldcs(reinterpret_cast<const uint4*>(x + (r + k * stride) * ld + c));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t q[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(q[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
    }
    for (; r < rows; r += stride) {
      const uint4 u = __ldcs(reinterpret_cast<const uint4*>(x + r * ld + c));
      const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(q[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  } else {
    for (; r < rows; r += stride)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c + j < cols) acc[j] += __bfloat162float(x[r * ld + c + j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[w][lane * 8 + j] = acc[j];
  __syncthreads();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col < cols) {
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += part[j][threadIdx.x];
    atomicAdd(out + col, sum);
  }
}

// ------------------------------------------------------------------------------------------------------
// a10: three-head label-smoothed loss and its gradient.  One block per sample.
// ------------------------------------------------------------------------------------------------------
__device__ float block_reduce(float v, bool is_max, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : (is_max ? -INFINITY : 0.0f);
  if (w == 0) {
    r = is_max ? warp_max(r) : warp_sum(r);
    if (lane == 0) sh[0] = r;
  }
  __syncthreads();
  return sh[0];
}

__global__ void answer_loss_kernel(const float* __restrict__ lc, const float* __restrict__ lv, const float* __restrict__ ls,
                                   const int64_t* __restrict__ answer, int B, int ncls, float eps, float grad_scale,
                                   float* __restrict__ loss, float* __restrict__ dc, float* __restrict__ dv, float* __restrict__ ds) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const float* L[3] = {lc + static_cast<long>(b) * ncls, lv + static_cast<long>(b) * ncls, ls + static_cast<long>(b) * ncls};
  float* D[3] = {dc ? dc + static_cast<long>(b) * ncls : nullptr, dv ? dv + static_cast<long>(b) * ncls : nullptr,
                 ds ? ds + static_cast<long>(b) * ncls : nullptr};
  const int ans = static_cast<int>(answer[b]);
  const float t_off = eps / ncls, t_on = (1.0f - eps) + eps / ncls;
  float sample_loss = 0.0f;
  for (int h = 0; h < 3; ++h) {
    float m = -INFINITY;
    for (int c = threadIdx.x; c < ncls; c += blockDim.x) m = fmaxf(m, L[h][c]);
    m = block_reduce(m, true, sh);
    float z = 0.0f;
    for (int c = threadIdx.x; c < ncls; c += blockDim.x) z += expf(L[h][c] - m);
    z = block_reduce(z, false, sh);
    const float lse = m + logf(z);
    float acc = 0.0f;
    for (int c = threadIdx.x; c < ncls; c += blockDim.x) {
      const float t = (c == ans) ? t_on : t_off;
      const float lsm = L[h][c] - lse;
      acc += t * lsm;
      // d/dlogit of -(1/3B) sum_c t_c lsm_c = (softmax_c * sum_c t_c - t_c) / (3B);  sum_c t_c == 1
      if (D[h]) D[h][c] = grad_scale * (expf(lsm) - t) / (3.0f * B);
    }
    acc = block_reduce(acc, false, sh);
    sample_loss -= acc / 3.0f;
  }
  if (threadIdx.x == 0) atomicAdd(loss, sample_loss / B);
}

// f3: Adam (torch.optim.Adam defaults: no weight decay, no amsgrad), bias correction as in torch.
// `dyn` (device, optional) = {lr / bias_correction1, sqrt(bias_correction2), step}: lets a captured CUDA graph see the
// per-step scalars without re-capture.
__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, float b1, float b2, float eps, float step_size,
                                          float bc2_sqrt) {
  m = b1 * m + (1.0f - b1) * g;
  v = b2 * v + (1.0f - b2) * g * g;
  p -= step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
  return p;
}

// 16-byte accesses over the flat buffers (28 B of HBM traffic per parameter, +2 B with the bf16 mirror); scalar tail.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long n, float lr, float b1, float b2, float eps, float bc1,
                                                   float bc2_sqrt, const float* __restrict__ dyn, __nv_bfloat16* __restrict__ pb, int vec) {
  float step_size = lr / bc1;
  if (dyn) {
    step_size = dyn[0];
    bc2_sqrt = dyn[1];
  }
  const long tid = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  long done = 0;
  if (vec) {
    const long n4 = n >> 2;
    for (long i = tid; i < n4; i += nthreads) {
      float4 pi = reinterpret_cast<float4*>(p)[i];
      const float4 gi = __ldcs(reinterpret_cast<const float4*>(g) + i);
      float4 mi = reinterpret_cast<float4*>(m)[i];
      float4 vi = reinterpret_cast<float4*>(v)[i];
      adam_one(pi.x, gi.x, mi.x, vi.x, b1, b2, eps, step_size, bc2_sqrt);
      adam_one(pi.y, gi.y, mi.y, vi.y, b1, b2, eps, step_size, bc2_sqrt);
      adam_one(pi.z, gi.z, mi.z, vi.z, b1, b2, eps, step_size, bc2_sqrt);
      adam_one(pi.w, gi.w, mi.w, vi.w, b1, b2, eps, step_size, bc2_sqrt);
      reinterpret_cast<float4*>(p)[i] = pi;
      reinterpret_cast<float4*>(m)[i] = mi;
      reinterpret_cast<float4*>(v)[i] = vi;
      if (pb) reinterpret_cast<uint2*>(pb)[i] = make_uint2(pack_bf16x2(pi.x, pi.y), pack_bf16x2(pi.z, pi.w));
    }
    done = n4 << 2;
  }
  for (long i = done + tid; i < n; i += nthreads) {
    float pi = p[i], mi = m[i], vi = v[i];
    adam_one(pi, g[i], mi, vi, b1, b2, eps, step_size, bc2_sqrt);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    if (pb) pb[i] = __float2bfloat16_rn(pi);
  }
}

// One thread: advances the device-resident step counter and derives the step's scalars from it, so that a replayed CUDA graph
// (and a host that runs several steps ahead of the GPU) always sees the values of ITS step: dyn = {lr / (1 - b1^s), sqrt(1 - b2^s), s}.
__global__ void adam_advance_kernel(float* __restrict__ dyn, float lr, float b1, float b2) {
  const double s = static_cast<double>(dyn[2]) + 1.0;
  dyn[0] = static_cast<float>(static_cast<double>(lr) / (1.0 - pow(static_cast<double>(b1), s)));
  dyn[1] = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(b2), s)));
  dyn[2] = static_cast<float>(s);
}

// Row-sparse Adam with the semantics of DENSE torch.optim.Adam ("deferred" Adam): stamp[row] = the last step the row is current
// through (0: never touched, moments zero).  A row named at step t is first brought up to date -- the steps it missed are
// replayed with a zero gradient, exactly what the dense optimizer did to it in the meantime (m *= b1, v *= b2,
// p -= lr_s m / (sqrt(v) / sqrt(bc2_s) + eps)) -- and then (apply != 0) takes its step-t update with the accumulated gradient,
// which is zeroed again.  apply == 0 is the catch-up alone (through step t - 1), run before the step's gathers read the row.
// The update of a zero-gradient step shrinks by ~b1 / sqrt(b2) per step, so the replay stops as soon as one step leaves every
// element of the row (chunk) unchanged -- from then on dense Adam's updates are lost in fp32 rounding as well -- and at the
// latest after kReplayMax steps; the remaining decay of the moments is applied in closed form.
// idx == nullptr: every row of the table (flush before a checkpoint / evaluation).
constexpr int kReplayMax = 256;
constexpr int kRowChunk = 10;  // elements per lane held in registers during a replay: 320 columns per pass (word rows: 300)

template <typename G>
__global__ void adam_rows_kernel(float* __restrict__ p, G* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 int* __restrict__ stamp, long table_rows, int width, const int64_t* __restrict__ idx, long n_idx, float lr,
                                 float b1, float b2, float eps, int step, const float* __restrict__ dyn, int apply) {
  if (dyn) step = static_cast<int>(dyn[2]);
  const int target = apply ? step : step - 1;   // the row is current through `target` when this kernel is done with it
  if (target < 1) return;
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  for (long r = warp0; r < n_idx; r += nwarps) {
    const long row = idx ? idx[r] : r;
    if (row < 0 || row >= table_rows) continue;
    int old = 0;
    if (lane == 0) {
      old = stamp[row];
      // a never-touched row (moments zero) has nothing to catch up on and keeps its stamp 0: it must not turn into a row that is
      // replayed (read and rewritten) at every later step; everything else is claimed by exactly one occurrence
      if (apply || old != 0) old = atomicMax(stamp + row, target);
    }
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old >= target) continue;  // another occurrence of this row took it (or it is already current)
    if (old == 0 && !apply) continue;
    const long base = row * static_cast<long>(width);
    // steps old+1 .. last_zero see a zero gradient; step `target` sees g when apply
    const int last_zero = apply ? target - 1 : target;
    const int first = (old == 0) ? last_zero + 1 : old + 1;  // never touched: moments are zero, nothing to replay
    const int replay_to = (last_zero - first + 1 > kReplayMax) ? first + kReplayMax - 1 : last_zero;
    float step_apply = 0.0f, bc2_apply = 1.0f;
    if (apply) {
      const double sd = static_cast<double>(target);
      step_apply = dyn ? dyn[0] : static_cast<float>(static_cast<double>(lr) / (1.0 - pow(static_cast<double>(b1), sd)));
      bc2_apply = dyn ? dyn[1] : static_cast<float>(sqrt(1.0 - pow(static_cast<double>(b2), sd)));
    }
    for (int c0 = 0; c0 < width; c0 += 32 * kRowChunk) {
      float pi[kRowChunk], mi[kRowChunk], vi[kRowChunk];
#pragma unroll
      for (int j = 0; j < kRowChunk; ++j) {
        const int c = c0 + j * 32 + lane;
        const bool ok = c < width;
        pi[j] = ok ? p[base + c] : 0.0f;
        mi[j] = ok ? m[base + c] : 0.0f;
        vi[j] = ok ? v[base + c] : 0.0f;
      }
      int s_done = first - 1;  // last zero-gradient step applied to this chunk
      bool live = first <= replay_to;
      for (int s0 = first; live && s0 <= replay_to; s0 += 32) {
        // per-step scalars of steps s0 .. s0+31, one per lane (the same double-precision formulas as adam_advance_kernel)
        const double sd = static_cast<double>(s0 + lane);
        const float ss_l = static_cast<float>(static_cast<double>(lr) / (1.0 - pow(static_cast<double>(b1), sd)));
        const float bc_l = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(b2), sd)));
        for (int k = 0; k < 32 && s0 + k <= replay_to; ++k) {
          const float ss = __shfl_sync(0xffffffffu, ss_l, k);
          const float bc = __shfl_sync(0xffffffffu, bc_l, k);
          bool changed = false;
#pragma unroll
          for (int j = 0; j < kRowChunk; ++j) {
            const float before = pi[j];
            adam_one(pi[j], 0.0f, mi[j], vi[j], b1, b2, eps, ss, bc);
            changed = changed || (pi[j] != before);
          }
          s_done = s0 + k;
          if (!__any_sync(0xffffffffu, changed)) {  // every later zero-gradient update is smaller still: lost in rounding too
            live = false;
            break;
          }
        }
      }
      if (s_done < last_zero && old != 0) {
        const double rest = static_cast<double>(last_zero - s_done);
        const float dm = static_cast<float>(pow(static_cast<double>(b1), rest));
        const float dv = static_cast<float>(pow(static_cast<double>(b2), rest));
#pragma unroll
        for (int j = 0; j < kRowChunk; ++j) {
          mi[j] *= dm;
          vi[j] *= dv;
        }
      }
#pragma unroll
      for (int j = 0; j < kRowChunk; ++j) {
        const int c = c0 + j * 32 + lane;
        if (c >= width) continue;
        if (apply) {
          adam_one(pi[j], load_row_grad(g, base + c), mi[j], vi[j], b1, b2, eps, step_apply, bc2_apply);
          clear_row_grad(g, base + c);
        }
        p[base + c] = pi[j];
        m[base + c] = mi[j];
        v[base + c] = vi[j];
      }
    }
  }
}

// The same update for rows of up to 384 columns (a multiple of 4; the word rows are 300): one warp per row occurrence, 16-byte
// accesses, and every load of the row (parameters, both moments, the gradient) is issued BEFORE the claim on the row's stamp is
// known -- a duplicate occurrence throws its loads away, every other one has paid a single memory latency instead of three.
// The scalar kernel above keeps one block per SM busy (137 registers) and three dependent round trips per row: 2 TB/s on 110 k
// rows; this one runs the gathered lists of 8 ranks at HBM speed.
constexpr int kVecChunks = 3;  // float4 groups per lane: 3 x 32 x 4 = 384 columns

__device__ __forceinline__ float4 load_row_grad4(const float* g, long i4) { return reinterpret_cast<const float4*>(g)[i4]; }
__device__ __forceinline__ float4 load_row_grad4(const long long* g, long i4) {
  const longlong2 a = reinterpret_cast<const longlong2*>(g)[2 * i4];
  const longlong2 b = reinterpret_cast<const longlong2*>(g)[2 * i4 + 1];
  return make_float4(__ll2float_rn(a.x) * kQ48Inv, __ll2float_rn(a.y) * kQ48Inv, __ll2float_rn(b.x) * kQ48Inv, __ll2float_rn(b.y) * kQ48Inv);
}
__device__ __forceinline__ void clear_row_grad4(float* g, long i4) { reinterpret_cast<float4*>(g)[i4] = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void clear_row_grad4(long long* g, long i4) {
  reinterpret_cast<longlong2*>(g)[2 * i4] = make_longlong2(0, 0);
  reinterpret_cast<longlong2*>(g)[2 * i4 + 1] = make_longlong2(0, 0);
}
__device__ __forceinline__ bool adam_zero4(float4& p, float4& m, float4& v, float b1, float b2, float eps, float ss, float bc) {
  const float4 before = p;
  adam_one(p.x, 0.0f, m.x, v.x, b1, b2, eps, ss, bc);
  adam_one(p.y, 0.0f, m.y, v.y, b1, b2, eps, ss, bc);
  adam_one(p.z, 0.0f, m.z, v.z, b1, b2, eps, ss, bc);
  adam_one(p.w, 0.0f, m.w, v.w, b1, b2, eps, ss, bc);
  return p.x != before.x || p.y != before.y || p.z != before.z || p.w != before.w;
}

template <typename G>
__global__ void __launch_bounds__(256, 2)
    adam_rows_vec_kernel(float* __restrict__ p, G* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int* __restrict__ stamp,
                         long table_rows, int width, const int64_t* __restrict__ idx, long n_idx, float lr, float b1, float b2, float eps,
                         int step, const float* __restrict__ dyn, int apply) {
  if (dyn) step = static_cast<int>(dyn[2]);
  const int target = apply ? step : step - 1;
  if (target < 1) return;
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  const int w4 = width >> 2;
  float step_apply = 0.0f, bc2_apply = 1.0f;
  if (apply) {
    const double sd = static_cast<double>(target);
    step_apply = dyn ? dyn[0] : static_cast<float>(static_cast<double>(lr) / (1.0 - pow(static_cast<double>(b1), sd)));
    bc2_apply = dyn ? dyn[1] : static_cast<float>(sqrt(1.0 - pow(static_cast<double>(b2), sd)));
  }
  for (long r = warp0; r < n_idx; r += nwarps) {
    const long row = idx ? idx[r] : r;
    if (row < 0 || row >= table_rows) continue;
    const long base4 = row * static_cast<long>(w4);
    int old = 0;
    if (lane == 0) old = stamp[row];
    float4 pi[kVecChunks], mi[kVecChunks], vi[kVecChunks], gi[kVecChunks];
#pragma unroll
    for (int j = 0; j < kVecChunks; ++j) {
      const int c = j * 32 + lane;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      pi[j] = mi[j] = vi[j] = gi[j] = z;
      if (c < w4) {
        pi[j] = reinterpret_cast<const float4*>(p)[base4 + c];
        mi[j] = reinterpret_cast<const float4*>(m)[base4 + c];
        vi[j] = reinterpret_cast<const float4*>(v)[base4 + c];
        if (apply) gi[j] = load_row_grad4(g, base4 + c);
      }
    }
    // a never-touched row keeps its stamp 0 in a catch-up; everything else is claimed by exactly one occurrence
    if (lane == 0 && (apply || old != 0)) old = atomicMax(stamp + row, target);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old >= target) continue;
    if (old == 0 && !apply) continue;
    const int last_zero = apply ? target - 1 : target;
    const int first = (old == 0) ? last_zero + 1 : old + 1;
    const int replay_to = (last_zero - first + 1 > kReplayMax) ? first + kReplayMax - 1 : last_zero;
    int s_done = first - 1;
    bool live = first <= replay_to;
    for (int s0 = first; live && s0 <= replay_to; s0 += 32) {
      const double sd = static_cast<double>(s0 + lane);
      const float ss_l = static_cast<float>(static_cast<double>(lr) / (1.0 - pow(static_cast<double>(b1), sd)));
      const float bc_l = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(b2), sd)));
      for (int k = 0; k < 32 && s0 + k <= replay_to; ++k) {
        const float ss = __shfl_sync(0xffffffffu, ss_l, k);
        const float bc = __shfl_sync(0xffffffffu, bc_l, k);
        bool changed = false;
#pragma unroll
        for (int j = 0; j < kVecChunks; ++j) changed = adam_zero4(pi[j], mi[j], vi[j], b1, b2, eps, ss, bc) || changed;
        s_done = s0 + k;
        if (!__any_sync(0xffffffffu, changed)) {
          live = false;
          break;
        }
      }
    }
    if (s_done < last_zero && old != 0) {
      const double rest = static_cast<double>(last_zero - s_done);
      const float dm = static_cast<float>(pow(static_cast<double>(b1), rest));
      const float dv = static_cast<float>(pow(static_cast<double>(b2), rest));
#pragma unroll
      for (int j = 0; j < kVecChunks; ++j) {
        mi[j].x *= dm; mi[j].y *= dm; mi[j].z *= dm; mi[j].w *= dm;
        vi[j].x *= dv; vi[j].y *= dv; vi[j].z *= dv; vi[j].w *= dv;
      }
    }
#pragma unroll
    for (int j = 0; j < kVecChunks; ++j) {
      const int c = j * 32 + lane;
      if (c >= w4) continue;
      if (apply) {
        adam_one(pi[j].x, gi[j].x, mi[j].x, vi[j].x, b1, b2, eps, step_apply, bc2_apply);
        adam_one(pi[j].y, gi[j].y, mi[j].y, vi[j].y, b1, b2, eps, step_apply, bc2_apply);
        adam_one(pi[j].z, gi[j].z, mi[j].z, vi[j].z, b1, b2, eps, step_apply, bc2_apply);
        adam_one(pi[j].w, gi[j].w, mi[j].w, vi[j].w, b1, b2, eps, step_apply, bc2_apply);
        clear_row_grad4(g, base4 + c);
      }
      reinterpret_cast<float4*>(p)[base4 + c] = pi[j];
      reinterpret_cast<float4*>(m)[base4 + c] = mi[j];
      reinterpret_cast<float4*>(v)[base4 + c] = vi[j];
    }
  }
}

inline int grid_for(long work_items, int threads, int per_sm = 8) {
  long blocks = (work_items + threads - 1) / threads;
  const long cap = static_cast<long>(sm_count()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace

int colsum_bf16(const void* x, int64_t ld, int64_t rows, int cols, float* out, cudaStream_t stream);
}  // namespace savqa

using namespace savqa;

extern "C" int savqa_build_masks(const void* first_mask, const void* q_mask, const void* q_graph, const void* first_graph,
                                 int in_is_float, int B, int V, int Q, int dec_mask_on, float* graph_diag, float* graph,
                                 float* dec_mask, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SAVQA_REQUIRE(B >= 0 && V >= 0 && Q >= 0, "savqa_build_masks: negative extent");
  if (B == 0 || V + Q == 0) return SAVQA_OK;
  SAVQA_REQUIRE(graph_diag && graph && dec_mask, "savqa_build_masks: null output");
  SAVQA_REQUIRE((V == 0 || first_mask) && (Q == 0 || (q_mask && q_graph)), "savqa_build_masks: null input");
  const long total = static_cast<long>(B) * (V + Q) * (V + Q);
  const int grid = grid_for(total, 256);
  if (in_is_float)
    build_masks_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(first_mask), static_cast<const float*>(q_mask),
                                                        static_cast<const float*>(q_graph), static_cast<const float*>(first_graph), B, V,
                                                        Q, dec_mask_on, graph_diag, graph, dec_mask);
  else
    build_masks_kernel<int><<<grid, 256, 0, stream>>>(static_cast<const int*>(first_mask), static_cast<const int*>(q_mask),
                                                      static_cast<const int*>(q_graph), static_cast<const int*>(first_graph), B, V, Q,
                                                      dec_mask_on, graph_diag, graph, dec_mask);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_pack_graph_bits(const float* graph, int64_t rows, int Tk, uint32_t* bits, int words_per_row, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows == 0 || Tk == 0) return SAVQA_OK;
  SAVQA_REQUIRE(graph && bits && rows > 0 && Tk > 0 && words_per_row * 32 >= Tk, "savqa_pack_graph_bits: bad argument");
  pack_graph_bits_kernel<<<grid_for(rows * words_per_row * 32, 256), 256, 0, stream>>>(graph, rows, Tk, bits, words_per_row);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_gather_rows(const float* table, int64_t table_rows, int width, const int64_t* idx, int64_t n_idx, float scale,
                                 float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16, int pad_to, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_idx == 0) return SAVQA_OK;
  SAVQA_REQUIRE(table && idx && width > 0 && n_idx > 0, "savqa_gather_rows: bad argument");
  SAVQA_REQUIRE(out_f32 || out_bf16, "savqa_gather_rows: no output");
  SAVQA_REQUIRE(!out_f32 || ld_f32 >= width, "savqa_gather_rows: ld_f32 < width");
  SAVQA_REQUIRE(!out_bf16 || (ld_bf16 >= width && ld_bf16 >= pad_to), "savqa_gather_rows: ld_bf16 too small");
  gather_rows_kernel<<<grid_for(n_idx * 32, 256), 256, 0, stream>>>(table, table_rows, width, idx, n_idx, scale, out_f32, ld_f32,
                                                                     static_cast<__nv_bfloat16*>(out_bf16), ld_bf16, pad_to);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_scatter_add_rows(float* dtable, int64_t table_rows, int width, const int64_t* idx, int64_t n_idx, const float* dout,
                                      int64_t ld_dout, float scale, int64_t skip_row, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_idx == 0) return SAVQA_OK;
  SAVQA_REQUIRE(dtable && idx && dout && width > 0 && ld_dout >= width, "savqa_scatter_add_rows: bad argument");
  scatter_add_rows_kernel<<<grid_for(n_idx * 32, 256), 256, 0, stream>>>(dtable, table_rows, width, idx, n_idx, dout, ld_dout, scale,
                                                                          skip_row);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_scatter_add_rows_q48(int64_t* acc, int64_t table_rows, int width, const int64_t* idx, int64_t n_idx,
                                          const float* dout, int64_t ld_dout, float scale, int64_t skip_row, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_idx == 0) return SAVQA_OK;
  SAVQA_REQUIRE(acc && idx && dout && width > 0 && ld_dout >= width, "savqa_scatter_add_rows_q48: bad argument");
  scatter_add_rows_q48_kernel<<<grid_for(n_idx * 32, 256), 256, 0, stream>>>(reinterpret_cast<long long*>(acc), table_rows, width, idx,
                                                                              n_idx, dout, ld_dout, scale, skip_row);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_cast_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int cols, int pad_to,
                               savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows == 0) return SAVQA_OK;
  if (pad_to < cols) pad_to = cols;
  SAVQA_REQUIRE(src && dst && cols > 0 && ld_src >= cols && ld_dst >= pad_to, "savqa_cast_bf16: bad argument");
  const long work = rows * ((pad_to + 3) / 4);
  cast_bf16_kernel<<<grid_for(work, 256), 256, 0, stream>>>(src, ld_src, static_cast<__nv_bfloat16*>(dst), ld_dst, rows, cols, pad_to);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_cast_transpose_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int cols, int pad_to,
                                         savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows == 0 || cols == 0) return SAVQA_OK;
  if (pad_to < rows) pad_to = static_cast<int>(rows);
  SAVQA_REQUIRE(src && dst && ld_src >= cols && ld_dst >= pad_to, "savqa_cast_transpose_bf16: bad argument");
  dim3 grid((cols + 31) / 32, static_cast<unsigned>((pad_to + 31) / 32));
  SAVQA_REQUIRE(grid.y <= 65535, "savqa_cast_transpose_bf16: too many rows");
  cast_transpose_bf16_kernel<<<grid, dim3(32, 8), 0, stream>>>(src, ld_src, static_cast<__nv_bfloat16*>(dst), ld_dst, rows, cols, pad_to);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_row_nonzero(const float* x, int64_t ld, int64_t rows, int cols, float* on, void* x_bf16, int64_t ld_bf16,
                                 savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows == 0) return SAVQA_OK;
  SAVQA_REQUIRE(x && cols > 0 && ld >= cols && (on || x_bf16), "savqa_row_nonzero: bad argument");
  row_nonzero_kernel<<<grid_for(rows * 32, 256), 256, 0, stream>>>(x, ld, rows, cols, on, static_cast<__nv_bfloat16*>(x_bf16), ld_bf16);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_relu_gate_bf16(const void* dy, int dy_is_f32, int64_t ld_dy, const void* act, int64_t ld_act, void* out, int64_t ld_out,
                                    int64_t rows, int cols, int64_t group_rows, int64_t group_stride, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows == 0 || cols == 0) return SAVQA_OK;
  SAVQA_REQUIRE(dy && act && out && ld_dy >= cols && ld_act >= cols && ld_out >= cols, "savqa_relu_gate_bf16: bad argument");
  if (group_rows <= 0) {
    group_rows = rows;
    group_stride = rows;
  }
  SAVQA_REQUIRE(group_stride >= group_rows, "savqa_relu_gate_bf16: group_stride < group_rows");
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int vec = (cols % 8 == 0 && ld_act % 8 == 0 && ld_out % 8 == 0 && ld_dy % (dy_is_f32 ? 4 : 8) == 0 && a16(dy) && a16(act) && a16(out)) ? 1 : 0;
  const int grid = grid_for(vec ? rows * (cols / 8) : rows * cols, 256);
  if (dy_is_f32)
    relu_gate_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(dy), ld_dy, static_cast<const __nv_bfloat16*>(act), ld_act,
                                                      static_cast<__nv_bfloat16*>(out), ld_out, rows, cols, group_rows, group_stride, vec);
  else
    relu_gate_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dy), ld_dy,
                                                              static_cast<const __nv_bfloat16*>(act), ld_act,
                                                              static_cast<__nv_bfloat16*>(out), ld_out, rows, cols, group_rows, group_stride, vec);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_regroup_cols(const void* in, int64_t ld_in, void* out, int64_t ld_out, int64_t rows, int groups, int w_in, int w_out,
                                  int elem_bytes, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows == 0) return SAVQA_OK;
  SAVQA_REQUIRE(in && out && groups > 0 && w_in > 0 && w_out > 0 && (elem_bytes == 2 || elem_bytes == 4), "savqa_regroup_cols: bad argument");
  SAVQA_REQUIRE(ld_in >= static_cast<int64_t>(groups) * w_in && ld_out >= static_cast<int64_t>(groups) * w_out, "savqa_regroup_cols: pitch");
  const int per16 = 16 / elem_bytes;
  if (w_in % per16 == 0 && w_out % per16 == 0 && ld_in % per16 == 0 && ld_out % per16 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    regroup_cols_vec_kernel<<<grid_for(rows * groups * (w_out / per16), 256), 256, 0, stream>>>(
        static_cast<const uint4*>(in), ld_in / per16, static_cast<uint4*>(out), ld_out / per16, rows, groups, w_in / per16, w_out / per16);
    SAVQA_CHECK_CUDA(cudaGetLastError());
    return SAVQA_OK;
  }
  const int grid = grid_for(rows * groups * w_out, 256);
  if (elem_bytes == 2)
    regroup_cols_kernel<uint16_t><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(in), ld_in, static_cast<uint16_t*>(out), ld_out, rows, groups,
                                                            w_in, w_out);
  else
    regroup_cols_kernel<uint32_t><<<grid, 256, 0, stream>>>(static_cast<const uint32_t*>(in), ld_in, static_cast<uint32_t*>(out), ld_out, rows, groups,
                                                            w_in, w_out);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_fill_zero(void* p, int64_t bytes, int max_blocks, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (bytes == 0) return SAVQA_OK;
  SAVQA_REQUIRE(p && bytes > 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0, "savqa_fill_zero: needs a 16-byte aligned buffer");
  const long n16 = bytes / 16;
  long blocks = (n16 + 255) / 256;
  const long cap = max_blocks > 0 ? max_blocks : static_cast<long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  fill_zero_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<uint4*>(p), n16, static_cast<uint8_t*>(p) + n16 * 16, bytes - n16 * 16);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_colsum_bf16(const void* x, int64_t ld, int64_t rows, int cols, float* out, savqa_stream_t stream_) {
  return savqa::colsum_bf16(x, ld, rows, cols, out, static_cast<cudaStream_t>(stream_));
}

int savqa::colsum_bf16(const void* x, int64_t ld, int64_t rows, int cols, float* out, cudaStream_t stream) {
  if (rows == 0 || cols == 0) return SAVQA_OK;
  SAVQA_REQUIRE(x && out && ld >= cols, "savqa_colsum_bf16: bad argument");
  const int strips = (cols + 255) / 256;
  long by = (rows + 31) / 32;  // >= 4 rows per warp
  const long cap = (static_cast<long>(sm_count()) * 4 + strips - 1) / strips;
  if (by > cap) by = cap;
  if (by < 1) by = 1;
  const int vec = (ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) ? 1 : 0;
  colsum_bf16_kernel<<<dim3(strips, static_cast<unsigned>(by)), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ld, rows, cols, out, vec);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_answer_loss(const float* lc, const float* lv, const float* ls, const int64_t* answer, int B, int ncls, float epsilon,
                                 float grad_scale, float* loss, float* dc, float* dv, float* ds, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SAVQA_REQUIRE(lc && lv && ls && answer && loss && B > 0 && ncls > 0, "savqa_answer_loss: bad argument");
  SAVQA_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  answer_loss_kernel<<<B, 256, 0, stream>>>(lc, lv, ls, answer, B, ncls, epsilon, grad_scale, loss, dc, dv, ds);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                               float beta2, float eps, int step, const float* dyn, void* param_bf16, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n == 0) return SAVQA_OK;
  SAVQA_REQUIRE(param && grad && exp_avg && exp_avg_sq && step >= 1, "savqa_adam_step: bad argument");
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int vec = (a16(param) && a16(grad) && a16(exp_avg) && a16(exp_avg_sq) && (!param_bf16 || (reinterpret_cast<uintptr_t>(param_bf16) & 7) == 0)) ? 1 : 0;
  // 6 resident blocks per SM (1536 of 2048 threads): the stream is HBM-bound long before that, and the word tables' row updates on
  // the helper stream find room next to it instead of waiting for its last block (step trace: they started 370 us late)
  adam_kernel<<<grid_for((n + 3) / 4, 256, 6), 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2),
                                                                  dyn, static_cast<__nv_bfloat16*>(param_bf16), vec);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_adam_advance(float* dyn, float lr, float beta1, float beta2, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SAVQA_REQUIRE(dyn && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "savqa_adam_advance: bad argument");
  adam_advance_kernel<<<1, 1, 0, stream>>>(dyn, lr, beta1, beta2);
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_adam_rows(float* param, void* grad, int grad_q48, float* exp_avg, float* exp_avg_sq, int32_t* row_stamp,
                               int64_t table_rows, int width, const int64_t* idx, int64_t n_idx, float lr, float beta1, float beta2,
                               float eps, int step, const float* dyn, int apply, savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (idx == nullptr) n_idx = table_rows;
  if (n_idx == 0) return SAVQA_OK;
  SAVQA_REQUIRE(param && exp_avg && exp_avg_sq && row_stamp && width > 0 && step >= 1 && (grad || !apply), "savqa_adam_rows: bad argument");
  SAVQA_REQUIRE(idx || !apply, "savqa_adam_rows: the whole-table form is the catch-up alone (apply == 0)");
  const bool vec = (width % 4 == 0) && width <= 128 * kVecChunks && (reinterpret_cast<uintptr_t>(param) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(exp_avg) % 16 == 0) && (reinterpret_cast<uintptr_t>(exp_avg_sq) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(grad) % 16 == 0);
  const int grid = grid_for(n_idx * 32, 256);
#define SAVQA_ADAM_ROWS(KERNEL, G)                                                                                                  \
  KERNEL<G><<<grid, 256, 0, stream>>>(param, static_cast<G*>(grad), exp_avg, exp_avg_sq, row_stamp, table_rows, width, idx, n_idx, lr, \
                                      beta1, beta2, eps, step, dyn, apply)
  if (vec) {
    if (grad_q48) SAVQA_ADAM_ROWS(adam_rows_vec_kernel, long long);
    else SAVQA_ADAM_ROWS(adam_rows_vec_kernel, float);
  } else {
    if (grad_q48) SAVQA_ADAM_ROWS(adam_rows_kernel, long long);
    else SAVQA_ADAM_ROWS(adam_rows_kernel, float);
  }
#undef SAVQA_ADAM_ROWS
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}
