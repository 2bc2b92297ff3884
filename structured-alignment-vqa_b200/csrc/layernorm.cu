// a6: fused residual + the reference's layer_normalization (modules.py:62-65) and its backward.
//   y = gamma * (x - mean) / (sigma + eps) + beta,  sigma = sqrt(sum (x-mean)^2 / (C-1))   (unbiased, eps added to sigma)
// HBM-streaming kernels: one warp per row, the row lives in registers (16-byte loads, warp-shuffle reductions),
// every byte is read once and written once.
#include "common.cuh"

namespace savqa {
namespace {

constexpr int kMaxV4 = 8;  // row cached in registers up to C = 8 * 128 = 1024

#ifndef SAVQA_LN_FWD_BLOCKS
#define SAVQA_LN_FWD_BLOCKS 1  // resident 256-thread blocks per SM the forward kernel is compiled for (register cap)
#endif
template <int NV>  // NV float4 per lane, C == NV * 128
__global__ void __launch_bounds__(256, SAVQA_LN_FWD_BLOCKS) ln_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                         long rows, float* __restrict__ pre, float* __restrict__ y,
                                                         __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ on,
                                                         float* __restrict__ stats) {
  constexpr int C = NV * 128;
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  float4 g[NV], b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    b[i] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
  }
  for (long r = warp0; r < rows; r += nwarps) {
    float4 v[NV];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = __ldcs(reinterpret_cast<const float4*>(x + r * C) + lane + 32 * i);
      if (res) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(res + r * C) + lane + 32 * i);
        v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (pre) *(reinterpret_cast<float4*>(pre + r * C) + lane + 32 * i) = v[i];
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float sigma = sqrtf(warp_sum(q) * (1.0f / (C - 1)));
    const float inv = 1.0f / (sigma + eps);
    if (stats && lane == 0) {
      stats[2 * r] = mean;
      stats[2 * r + 1] = sigma;
    }
    float ys = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = g[i].x * v[i].x * inv + b[i].x;
      o.y = g[i].y * v[i].y * inv + b[i].y;
      o.z = g[i].z * v[i].z * inv + b[i].z;
      o.w = g[i].w * v[i].w * inv + b[i].w;
      ys += (o.x + o.y) + (o.z + o.w);
      *(reinterpret_cast<float4*>(y + r * C) + lane + 32 * i) = o;
      if (y_bf16) *(reinterpret_cast<uint2*>(y_bf16 + r * C) + lane + 32 * i) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
    if (on) {
      ys = warp_sum(ys);
      if (lane == 0) on[r] = (ys != 0.0f) ? 1.0f : 0.0f;
    }
  }
}

// generic width (e.g. C = 64 in the small golden model): one warp per row, three L1-resident passes
__global__ void ln_fwd_generic_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float eps, long rows, int C, float* __restrict__ pre,
                                      float* __restrict__ y, __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ on,
                                      float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  for (long r = warp0; r < rows; r += nwarps) {
    float s = 0.0f;
    for (int c = lane; c < C; c += 32) s += x[r * C + c] + (res ? res[r * C + c] : 0.0f);
    const float mean = warp_sum(s) / C;
    float q = 0.0f;
    for (int c = lane; c < C; c += 32) {
      const float p = x[r * C + c] + (res ? res[r * C + c] : 0.0f);
      if (pre) pre[r * C + c] = p;
      q += (p - mean) * (p - mean);
    }
    const float sigma = sqrtf(warp_sum(q) / (C - 1));
    const float inv = 1.0f / (sigma + eps);
    if (stats && lane == 0) {
      stats[2 * r] = mean;
      stats[2 * r + 1] = sigma;
    }
    float ys = 0.0f;
    for (int c = lane; c < C; c += 32) {
      const float p = x[r * C + c] + (res ? res[r * C + c] : 0.0f);
      const float o = gamma[c] * (p - mean) * inv + beta[c];
      ys += o;
      y[r * C + c] = o;
      if (y_bf16) y_bf16[r * C + c] = __float2bfloat16_rn(o);
    }
    if (on) {
      ys = warp_sum(ys);
      if (lane == 0) on[r] = (ys != 0.0f) ? 1.0f : 0.0f;
    }
  }
}

// backward: dx = (g - mean g)/s - c * dot(g,c) / ((C-1) sigma s^2)  [second term dropped when sigma == 0, like autograd]
//           dgamma += sum_rows dy * c / s ; dbeta += sum_rows dy
template <int NV>
__global__ void __launch_bounds__(256, 2) ln_bwd_vec_kernel(const float* __restrict__ dy, const float* __restrict__ pre,
                                                         const float* __restrict__ gamma, float eps, long rows,
                                                         const float* __restrict__ dres_in, float* __restrict__ dx,
                                                         __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta, float* __restrict__ dxsum) {
  constexpr int C = NV * 128;
  __shared__ __align__(16) float red[8][NV * 128 + 4];  // per-warp partials of one parameter gradient at a time
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  float4 gm[NV], ag[NV], ab[NV], ax[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    gm[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ax[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long r = warp0; r < rows; r += nwarps) {
    float4 c[NV], d[NV];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      c[i] = __ldcs(reinterpret_cast<const float4*>(pre + r * C) + lane + 32 * i);
      d[i] = __ldcs(reinterpret_cast<const float4*>(dy + r * C) + lane + 32 * i);
      s += (c[i].x + c[i].y) + (c[i].z + c[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      c[i].x -= mean; c[i].y -= mean; c[i].z -= mean; c[i].w -= mean;
      q += (c[i].x * c[i].x + c[i].y * c[i].y) + (c[i].z * c[i].z + c[i].w * c[i].w);
    }
    const float sigma = sqrtf(warp_sum(q) * (1.0f / (C - 1)));
    const float sden = sigma + eps;
    const float inv = 1.0f / sden;
    float sg = 0.0f, dot = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      ab[i].x += d[i].x; ab[i].y += d[i].y; ab[i].z += d[i].z; ab[i].w += d[i].w;
      ag[i].x += d[i].x * c[i].x * inv; ag[i].y += d[i].y * c[i].y * inv;
      ag[i].z += d[i].z * c[i].z * inv; ag[i].w += d[i].w * c[i].w * inv;
      d[i].x *= gm[i].x; d[i].y *= gm[i].y; d[i].z *= gm[i].z; d[i].w *= gm[i].w;  // g = dy * gamma
      sg += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      dot += (d[i].x * c[i].x + d[i].y * c[i].y) + (d[i].z * c[i].z + d[i].w * c[i].w);
    }
    const float mg = warp_sum(sg) * (1.0f / C);
    dot = warp_sum(dot);
    const float k2 = (sigma > 0.0f) ? dot / (static_cast<float>(C - 1) * sigma * sden * sden) : 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = (d[i].x - mg) * inv - c[i].x * k2;
      o.y = (d[i].y - mg) * inv - c[i].y * k2;
      o.z = (d[i].z - mg) * inv - c[i].z * k2;
      o.w = (d[i].w - mg) * inv - c[i].w * k2;
      if (dres_in) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(dres_in + r * C) + lane + 32 * i);
        o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
      }
      *(reinterpret_cast<float4*>(dx + r * C) + lane + 32 * i) = o;
      if (dx_bf16) *(reinterpret_cast<uint2*>(dx_bf16 + r * C) + lane + 32 * i) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      ax[i].x += o.x; ax[i].y += o.y; ax[i].z += o.z; ax[i].w += o.w;
    }
  }
  // block reduction of the parameter gradients (one after the other through the same buffer), then one atomic per
  // column per block
  const int nw = blockDim.x >> 5;
#pragma unroll 1
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dxsum);
    if (!dst) continue;  // uniform over the block
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 a = which == 0 ? ag[i] : (which == 1 ? ab[i] : ax[i]);
      *reinterpret_cast<float4*>(&red[w][(lane + 32 * i) * 4]) = a;
    }
    __syncthreads();
    for (int col = threadIdx.x; col < C; col += blockDim.x) {
      float a = 0.0f;
      for (int j = 0; j < nw; ++j) a += red[j][col];
      atomicAdd(dst + col, a);
    }
  }
}

__global__ void ln_bwd_generic_kernel(const float* __restrict__ dy, const float* __restrict__ pre, const float* __restrict__ gamma,
                                      float eps, long rows, int C, const float* __restrict__ dres_in, float* __restrict__ dx,
                                      __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                      float* __restrict__ dxsum) {
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  for (long r = warp0; r < rows; r += nwarps) {
    float s = 0.0f;
    for (int c = lane; c < C; c += 32) s += pre[r * C + c];
    const float mean = warp_sum(s) / C;
    float q = 0.0f, sg = 0.0f, dot = 0.0f;
    for (int c = lane; c < C; c += 32) {
      const float cc = pre[r * C + c] - mean;
      const float g = dy[r * C + c] * gamma[c];
      q += cc * cc;
      sg += g;
      dot += g * cc;
    }
    const float sigma = sqrtf(warp_sum(q) / (C - 1));
    const float sden = sigma + eps, inv = 1.0f / sden;
    const float mg = warp_sum(sg) / C;
    dot = warp_sum(dot);
    const float k2 = (sigma > 0.0f) ? dot / (static_cast<float>(C - 1) * sigma * sden * sden) : 0.0f;
    for (int c = lane; c < C; c += 32) {
      const float cc = pre[r * C + c] - mean;
      const float d = dy[r * C + c];
      float o = (d * gamma[c] - mg) * inv - cc * k2;
      if (dres_in) o += dres_in[r * C + c];
      dx[r * C + c] = o;
      if (dx_bf16) dx_bf16[r * C + c] = __float2bfloat16_rn(o);
      if (dgamma) atomicAdd(dgamma + c, d * cc * inv);
      if (dbeta) atomicAdd(dbeta + c, d);
      if (dxsum) atomicAdd(dxsum + c, o);
    }
  }
}

inline int ln_grid(long rows, int per_sm) {
  long blocks = (rows + 7) / 8;  // 8 warps per 256-thread block
  const long cap = static_cast<long>(sm_count()) * per_sm;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

inline bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace savqa

using namespace savqa;

extern "C" int savqa_residual_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, float eps,
                                            int64_t rows, int C, float* pre, float* y, void* y_bf16, float* on, float* stats,
                                            savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows == 0) return SAVQA_OK;
  SAVQA_REQUIRE(x && gamma && beta && y && C >= 2, "savqa_residual_layernorm_fwd: bad argument");
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y_bf16);
  const bool vec = (C % 128 == 0) && (C / 128 <= kMaxV4) && a16(x) && a16(gamma) && a16(beta) && a16(y) && (!res || a16(res)) &&
                   (!pre || a16(pre)) && (!yb || a16(yb));
  const int grid = ln_grid(rows, 8);
  if (vec) {
    switch (C / 128) {
#define LN_CASE(NV) case NV: SAVQA_CHECK_CUDA(launch_kernel(true, ln_fwd_vec_kernel<NV>, dim3(grid), dim3(256), 0, stream, x, res, gamma, beta, eps, static_cast<long>(rows), pre, y, yb, on, stats)); break;
      LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
#undef LN_CASE
    }
  } else {
    ln_fwd_generic_kernel<<<grid, 256, 0, stream>>>(x, res, gamma, beta, eps, rows, C, pre, y, yb, on, stats);
  }
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}

extern "C" int savqa_layernorm_bwd(const float* dy, const float* pre, const float* gamma, float eps, int64_t rows, int C,
                                   const float* dres_in, float* dx, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum,
                                   savqa_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows == 0) return SAVQA_OK;
  SAVQA_REQUIRE(dy && pre && gamma && dx && C >= 2, "savqa_layernorm_bwd: bad argument");
  __nv_bfloat16* db = static_cast<__nv_bfloat16*>(dx_bf16);
  const bool vec = (C % 128 == 0) && (C / 128 <= 4) && a16(dy) && a16(pre) && a16(gamma) && a16(dx) && (!dres_in || a16(dres_in)) &&
                   (!db || a16(db));
  if (vec) {
    const int grid = ln_grid(rows, 2);  // few, fat blocks: each ends with C atomics per parameter (4 per SM measured slower)
    switch (C / 128) {
#define LN_CASE(NV) case NV: SAVQA_CHECK_CUDA(launch_kernel(true, ln_bwd_vec_kernel<NV>, dim3(grid), dim3(256), 0, stream, dy, pre, gamma, eps, static_cast<long>(rows), dres_in, dx, db, dgamma, dbeta, dxsum)); break;
      LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4)
#undef LN_CASE
    }
  } else {
    ln_bwd_generic_kernel<<<ln_grid(rows, 8), 256, 0, stream>>>(dy, pre, gamma, eps, rows, C, dres_in, dx, db, dgamma, dbeta, dxsum);
  }
  SAVQA_CHECK_CUDA(cudaGetLastError());
  return SAVQA_OK;
}
