// Conv2d-subsampling front-end kernels (reference: neurst/layers/modalities/audio_modalities.py:84-109).
//
// conv1 (Cin=1, K=9) is bandwidth-bound: a direct CUDA-core convolution with the channel LayerNorm + ReLU
// fused in (one warp per output position, channels across lanes, warp-shuffle statistics).  Its backward
// recomputes the pre-LN activation from the fbank tile instead of storing 164 M activations.
// conv2 is lowered to a tcgen05 GEMM through im2col / col2im (3x3, stride 2, pad 1).
#include "kernels.cuh"

namespace b200st {

#define DISPATCH_DTYPE(dt, T, ...)                                   \
  do {                                                               \
    if ((dt) == F32) { using T = float; __VA_ARGS__; }               \
    else if ((dt) == BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else B200ST_FAIL("bad dtype");                                   \
  } while (0)

constexpr int CONV_MAX_CPL = 16;   // channels per lane (C <= 512)

// ---------------------------------------------------------------------------------------------
// conv1 + LN + ReLU forward
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv1_fwd_kernel(const float* __restrict__ src, const float* __restrict__ w,
                                                         const float* __restrict__ bias, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float eps, T* __restrict__ y, int B,
                                                         int Tn, int F, int Cin, int C, int T1, int F1, int use_ln) {
  extern __shared__ float sw[];   // [9*Cin][C] weights, then bias, gamma, beta
  const int nw = 9 * Cin * C;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
  float* sb = sw + nw; float* sg = sb + C; float* sbe = sg + C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sb[i] = bias[i]; sg[i] = use_ln ? gamma[i] : 1.f; sbe[i] = use_ln ? beta[i] : 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t npos = (int64_t)B * T1 * F1;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t pos = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pos < npos; pos += warps_total) {
    const int f1 = (int)(pos % F1);
    const int t1 = (int)((pos / F1) % T1);
    const int b = (int)(pos / ((int64_t)F1 * T1));
    float acc[CONV_MAX_CPL];
#pragma unroll
    for (int i = 0; i < CONV_MAX_CPL; ++i) { const int c = lane + 32 * i; acc[i] = c < C ? sb[c] : 0.f; }
    for (int kh = 0; kh < 3; ++kh) {
      const int t = 2 * t1 + kh - 1;
      if (t < 0 || t >= Tn) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int f = 2 * f1 + kw - 1;
        if (f < 0 || f >= F) continue;
        for (int ci = 0; ci < Cin; ++ci) {
          const float xv = __ldg(&src[(((int64_t)b * Tn + t) * F + f) * Cin + ci]);
          const float* wr = sw + ((kh * 3 + kw) * Cin + ci) * C;
#pragma unroll
          for (int i = 0; i < CONV_MAX_CPL; ++i) { const int c = lane + 32 * i; if (c < C) acc[i] = fmaf(xv, wr[c], acc[i]); }
        }
      }
    }
    float mean = 0.f, rstd = 1.f;
    if (use_ln) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < CONV_MAX_CPL; ++i) if (lane + 32 * i < C) sum += acc[i];
      mean = warp_sum(sum) / C;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < CONV_MAX_CPL; ++i) if (lane + 32 * i < C) { const float dl = acc[i] - mean; sq += dl * dl; }
      rstd = 1.0f / sqrtf(warp_sum(sq) / C + eps);
    }
    T* yr = y + pos * C;
#pragma unroll
    for (int i = 0; i < CONV_MAX_CPL; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        float v = use_ln ? (acc[i] - mean) * rstd * sg[c] + sbe[c] : acc[i];
        yr[c] = from_f32<T>(fmaxf(v, 0.f));
      }
    }
  }
}

int conv1_ln_relu_fwd(const float* src, const float* w, const float* b, const float* gamma, const float* beta, float eps,
                      void* y1, int y_dtype, int B, int T, int F, int Cin, int C, int use_ln, cudaStream_t s) {
  B200ST_CHECK(C <= 32 * CONV_MAX_CPL, "conv channels must be <= 512");
  const int T1 = (T + 1) / 2, F1 = (F + 1) / 2;
  const int64_t npos = (int64_t)B * T1 * F1;
  if (npos == 0) return 0;
  const size_t smem = (size_t)(9 * Cin * C + 3 * C) * sizeof(float);
  B200ST_CHECK(smem <= 200 * 1024, "conv1 weights do not fit shared memory");
  int grid = (int)((npos + 63) / 64);
  if (grid > 148 * 4) grid = 148 * 4;
  DISPATCH_DTYPE(y_dtype, TT, {
    auto kern = conv1_fwd_kernel<TT>;
    if (smem > 48 * 1024) B200ST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, s>>>(src, w, b, gamma, beta, eps, (TT*)y1, B, T, F, Cin, C, T1, F1, use_ln);
  });
  ++g_kernel_launches;
  B200ST_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// conv1 + LN + ReLU backward (recompute): dW[9*Cin][C], db[C], dgamma[C], dbeta[C]
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv1_bwd_kernel(const float* __restrict__ src, const float* __restrict__ w,
                                                         const float* __restrict__ bias, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float eps, const T* __restrict__ y,
                                                         const T* __restrict__ dy, float* __restrict__ dw,
                                                         float* __restrict__ db, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta, int B, int Tn, int F, int Cin, int C,
                                                         int T1, int F1, int use_ln) {
  extern __shared__ float sw[];   // weights [9*Cin][C] | bias | gamma | beta | accum: dw [9*Cin][C] | db | dgamma | dbeta
  const int nw = 9 * Cin * C;
  float* sb = sw + nw; float* sg = sb + C; float* sbe = sg + C;
  float* adw = sbe + C; float* adb = adw + nw; float* adg = adb + C; float* adbe = adg + C;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) { sw[i] = w[i]; adw[i] = 0.f; }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    sb[i] = bias[i]; sg[i] = use_ln ? gamma[i] : 1.f; sbe[i] = use_ln ? beta[i] : 0.f;
    adb[i] = 0.f; adg[i] = 0.f; adbe[i] = 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t npos = (int64_t)B * T1 * F1;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t pos = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pos < npos; pos += warps_total) {
    const int f1 = (int)(pos % F1);
    const int t1 = (int)((pos / F1) % T1);
    const int b = (int)(pos / ((int64_t)F1 * T1));
    float acc[CONV_MAX_CPL];
#pragma unroll
    for (int i = 0; i < CONV_MAX_CPL; ++i) { const int c = lane + 32 * i; acc[i] = c < C ? sb[c] : 0.f; }
    for (int kh = 0; kh < 3; ++kh) {
      const int t = 2 * t1 + kh - 1;
      if (t < 0 || t >= Tn) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int f = 2 * f1 + kw - 1;
        if (f < 0 || f >= F) continue;
        for (int ci = 0; ci < Cin; ++ci) {
          const float xv = __ldg(&src[(((int64_t)b * Tn + t) * F + f) * Cin + ci]);
          const float* wr = sw + ((kh * 3 + kw) * Cin + ci) * C;
#pragma unroll
          for (int i = 0; i < CONV_MAX_CPL; ++i) { const int c = lane + 32 * i; if (c < C) acc[i] = fmaf(xv, wr[c], acc[i]); }
        }
      }
    }
    float mean = 0.f, rstd = 1.f;
    if (use_ln) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < CONV_MAX_CPL; ++i) if (lane + 32 * i < C) sum += acc[i];
      mean = warp_sum(sum) / C;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < CONV_MAX_CPL; ++i) if (lane + 32 * i < C) { const float dl = acc[i] - mean; sq += dl * dl; }
      rstd = 1.0f / sqrtf(warp_sum(sq) / C + eps);
    }
    // dz = LN'(dy * relu') ; acc[] is reused to hold dz
    float c1 = 0.f, c2 = 0.f;
    float dl_[CONV_MAX_CPL];
#pragma unroll
    for (int i = 0; i < CONV_MAX_CPL; ++i) {
      const int c = lane + 32 * i;
      dl_[i] = 0.f;
      if (c < C) {
        float d = to_f32(dy[pos * C + c]);
        if (!(to_f32(y[pos * C + c]) > 0.f)) d = 0.f;      // ReLU mask from the stored output
        dl_[i] = d;
        if (use_ln) {
          const float xh = (acc[i] - mean) * rstd;
          const float g = d * sg[c];
          c1 += g; c2 += g * xh;
          atomicAdd(&adg[c], d * xh);
          atomicAdd(&adbe[c], d);
        }
      }
    }
    if (use_ln) { c1 = warp_sum(c1) / C; c2 = warp_sum(c2) / C; }
#pragma unroll
    for (int i = 0; i < CONV_MAX_CPL; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        float dz = dl_[i];
        if (use_ln) { const float xh = (acc[i] - mean) * rstd; dz = rstd * (dl_[i] * sg[c] - c1 - xh * c2); }
        acc[i] = dz;
        atomicAdd(&adb[c], dz);
      } else acc[i] = 0.f;
    }
    for (int kh = 0; kh < 3; ++kh) {
      const int t = 2 * t1 + kh - 1;
      if (t < 0 || t >= Tn) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int f = 2 * f1 + kw - 1;
        if (f < 0 || f >= F) continue;
        for (int ci = 0; ci < Cin; ++ci) {
          const float xv = __ldg(&src[(((int64_t)b * Tn + t) * F + f) * Cin + ci]);
          float* ar = adw + ((kh * 3 + kw) * Cin + ci) * C;
#pragma unroll
          for (int i = 0; i < CONV_MAX_CPL; ++i) { const int c = lane + 32 * i; if (c < C) atomicAdd(&ar[c], xv * acc[i]); }
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nw; i += blockDim.x) atomicAdd(&dw[i], adw[i]);
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&db[i], adb[i]);
    if (use_ln) { atomicAdd(&dgamma[i], adg[i]); atomicAdd(&dbeta[i], adbe[i]); }
  }
}


// Fast path for Cin == 1 (log-mel fbank): per-lane register accumulators for dW / db / dgamma / dbeta over all
// positions a warp visits; one shared-memory flush per warp at the end.
template <typename T, int CPL>
__global__ void __launch_bounds__(256) conv1_bwd_cin1_kernel(const float* __restrict__ src, const float* __restrict__ w,
                                                              const float* __restrict__ bias, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, const T* __restrict__ y,
                                                              const T* __restrict__ dy, float* __restrict__ dw,
                                                              float* __restrict__ db, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, int B, int Tn, int F, int C, int T1,
                                                              int F1, int use_ln) {
  extern __shared__ float sw[];   // accum: dw [9][C] | db | dgamma | dbeta
  const int nw = 9 * C;
  float* adw = sw; float* adb = adw + nw; float* adg = adb + C; float* adbe = adg + C;
  for (int i = threadIdx.x; i < nw + 3 * C; i += blockDim.x) sw[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float wreg[9][CPL], breg[CPL], greg[CPL], bereg[CPL];
  float a_dw[9][CPL], a_db[CPL], a_dg[CPL], a_dbe[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = lane + 32 * i;
    const bool ok = c < C;
    breg[i] = ok ? bias[c] : 0.f; greg[i] = (ok && use_ln) ? gamma[c] : 1.f; bereg[i] = (ok && use_ln) ? beta[c] : 0.f;
    a_db[i] = a_dg[i] = a_dbe[i] = 0.f;
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) { wreg[tp][i] = ok ? w[tp * C + c] : 0.f; a_dw[tp][i] = 0.f; }
  }
  const int64_t npos = (int64_t)B * T1 * F1;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t pos = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pos < npos; pos += warps_total) {
    const int f1 = (int)(pos % F1);
    const int t1 = (int)((pos / F1) % T1);
    const int b = (int)(pos / ((int64_t)F1 * T1));
    float xv[9];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      const int t = 2 * t1 + tp / 3 - 1, f = 2 * f1 + tp % 3 - 1;
      xv[tp] = (t >= 0 && t < Tn && f >= 0 && f < F) ? __ldg(&src[((int64_t)b * Tn + t) * F + f]) : 0.f;
    }
    float z[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      float a = breg[i];
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) a = fmaf(xv[tp], wreg[tp][i], a);
      z[i] = a;
    }
    float mean = 0.f, rstd = 1.f;
    if (use_ln) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) if (lane + 32 * i < C) sum += z[i];
      mean = warp_sum(sum) / C;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) if (lane + 32 * i < C) { const float dl = z[i] - mean; sq += dl * dl; }
      rstd = 1.0f / sqrtf(warp_sum(sq) / C + eps);
    }
    float d[CPL];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = lane + 32 * i;
      d[i] = 0.f;
      if (c < C) {
        float dd = to_f32(dy[pos * C + c]);
        if (!(to_f32(y[pos * C + c]) > 0.f)) dd = 0.f;
        d[i] = dd;
        if (use_ln) {
          const float xh = (z[i] - mean) * rstd;
          const float g = dd * greg[i];
          c1 += g; c2 += g * xh;
          a_dg[i] += dd * xh; a_dbe[i] += dd;
        }
      }
    }
    if (use_ln) { c1 = warp_sum(c1) / C; c2 = warp_sum(c2) / C; }
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      float dz = d[i];
      if (use_ln) { const float xh = (z[i] - mean) * rstd; dz = rstd * (d[i] * greg[i] - c1 - xh * c2); }
      if (lane + 32 * i >= C) dz = 0.f;
      a_db[i] += dz;
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) a_dw[tp][i] = fmaf(xv[tp], dz, a_dw[tp][i]);
    }
  }
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      atomicAdd(&adb[c], a_db[i]); atomicAdd(&adg[c], a_dg[i]); atomicAdd(&adbe[c], a_dbe[i]);
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) atomicAdd(&adw[tp * C + c], a_dw[tp][i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nw; i += blockDim.x) atomicAdd(&dw[i], adw[i]);
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&db[i], adb[i]);
    if (use_ln) { atomicAdd(&dgamma[i], adg[i]); atomicAdd(&dbeta[i], adbe[i]); }
  }
}

int conv1_ln_relu_bwd(const float* src, const float* w, const float* b, const float* gamma, const float* beta, float eps,
                      const void* y1, const void* dy1, int dtype, float* dw, float* db, float* dgamma, float* dbeta,
                      int B, int T, int F, int Cin, int C, int use_ln, cudaStream_t s) {
  B200ST_CHECK(C <= 32 * CONV_MAX_CPL, "conv channels must be <= 512");
  const int T1 = (T + 1) / 2, F1 = (F + 1) / 2;
  const int64_t npos = (int64_t)B * T1 * F1;
  if (npos == 0) return 0;
  const size_t smem = (size_t)(2 * (9 * Cin * C + 3 * C)) * sizeof(float);
  B200ST_CHECK(smem <= 200 * 1024, "conv1 weights do not fit shared memory");
  int grid = (int)((npos + 63) / 64);
  if (grid > 148 * 2) grid = 148 * 2;
  if (Cin == 1 && C <= 256) {
    const size_t smem1 = (size_t)(9 * C + 3 * C) * sizeof(float);
    DISPATCH_DTYPE(dtype, TT, {
      if (C <= 32) conv1_bwd_cin1_kernel<TT, 1><<<grid, 256, smem1, s>>>(src, w, b, gamma, beta, eps, (const TT*)y1, (const TT*)dy1, dw, db, dgamma, dbeta, B, T, F, C, T1, F1, use_ln);
      else if (C <= 64) conv1_bwd_cin1_kernel<TT, 2><<<grid, 256, smem1, s>>>(src, w, b, gamma, beta, eps, (const TT*)y1, (const TT*)dy1, dw, db, dgamma, dbeta, B, T, F, C, T1, F1, use_ln);
      else if (C <= 128) conv1_bwd_cin1_kernel<TT, 4><<<grid, 256, smem1, s>>>(src, w, b, gamma, beta, eps, (const TT*)y1, (const TT*)dy1, dw, db, dgamma, dbeta, B, T, F, C, T1, F1, use_ln);
      else conv1_bwd_cin1_kernel<TT, 8><<<grid, 256, smem1, s>>>(src, w, b, gamma, beta, eps, (const TT*)y1, (const TT*)dy1, dw, db, dgamma, dbeta, B, T, F, C, T1, F1, use_ln);
    });
    ++g_kernel_launches;
    B200ST_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_DTYPE(dtype, TT, {
    auto kern = conv1_bwd_kernel<TT>;
    if (smem > 48 * 1024) B200ST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, s>>>(src, w, b, gamma, beta, eps, (const TT*)y1, (const TT*)dy1, dw, db, dgamma, dbeta, B, T, F,
                                 Cin, C, T1, F1, use_ln);
  });
  ++g_kernel_launches;
  B200ST_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// im2col / col2im for the 3x3 stride-2 pad-1 conv2 (NHWC)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ y1, T* __restrict__ col, int B, int T1, int F1,
                                                      int C, int T2, int F2) {
  const int64_t nchunks = (int64_t)B * T2 * F2 * 9;   // one (row, tap) chunk of C channels per warp
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t ch = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ch < nchunks; ch += warps_total) {
    const int tap = (int)(ch % 9);
    const int64_t row = ch / 9;
    const int f2 = (int)(row % F2);
    const int t2 = (int)((row / F2) % T2);
    const int b = (int)(row / ((int64_t)F2 * T2));
    const int t = 2 * t2 + tap / 3 - 1, f = 2 * f2 + tap % 3 - 1;
    T* dst = col + (row * 9 + tap) * C;
    if (t < 0 || t >= T1 || f < 0 || f >= F1) {
      for (int c = lane; c < C; c += 32) dst[c] = from_f32<T>(0.f);
    } else {
      const T* srcp = y1 + (((int64_t)b * T1 + t) * F1 + f) * C;
      for (int c = lane; c < C; c += 32) dst[c] = srcp[c];
    }
  }
}
int im2col_3x3s2(const void* y1, void* col, int dtype, int B, int T1, int F1, int C, cudaStream_t s) {
  const int T2 = (T1 + 1) / 2, F2 = (F1 + 1) / 2;
  const int64_t nchunks = (int64_t)B * T2 * F2 * 9;
  if (nchunks == 0) return 0;
  int64_t g = (nchunks + 7) / 8;
  const int grid = (int)(g > 148 * 16 ? 148 * 16 : g);
  DISPATCH_DTYPE(dtype, TT, (im2col_kernel<TT><<<grid, 256, 0, s>>>((const TT*)y1, (TT*)col, B, T1, F1, C, T2, F2)));
  ++g_kernel_launches;
  B200ST_LAUNCH_CHECK();
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(256) col2im_kernel(const T* __restrict__ dcol, T* __restrict__ dy1, int B, int T1, int F1,
                                                      int C, int T2, int F2) {
  const int64_t npos = (int64_t)B * T1 * F1;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t pos = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pos < npos; pos += warps_total) {
    const int f1 = (int)(pos % F1);
    const int t1 = (int)((pos / F1) % T1);
    const int b = (int)(pos / ((int64_t)F1 * T1));
    for (int c = lane; c < C; c += 32) {
      float acc = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int tt = t1 + 1 - kh;
        if (tt < 0 || (tt & 1)) continue;
        const int t2 = tt >> 1;
        if (t2 >= T2) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int ff = f1 + 1 - kw;
          if (ff < 0 || (ff & 1)) continue;
          const int f2 = ff >> 1;
          if (f2 >= F2) continue;
          const int64_t row = ((int64_t)b * T2 + t2) * F2 + f2;
          acc += to_f32(dcol[(row * 9 + kh * 3 + kw) * C + c]);
        }
      }
      dy1[pos * C + c] = from_f32<T>(acc);
    }
  }
}
int col2im_3x3s2(const void* dcol, void* dy1, int dtype, int B, int T1, int F1, int C, cudaStream_t s) {
  const int T2 = (T1 + 1) / 2, F2 = (F1 + 1) / 2;
  const int64_t npos = (int64_t)B * T1 * F1;
  if (npos == 0) return 0;
  int64_t g = (npos + 7) / 8;
  const int grid = (int)(g > 148 * 16 ? 148 * 16 : g);
  DISPATCH_DTYPE(dtype, TT, (col2im_kernel<TT><<<grid, 256, 0, s>>>((const TT*)dcol, (TT*)dy1, B, T1, F1, C, T2, F2)));
  ++g_kernel_launches;
  B200ST_LAUNCH_CHECK();
  return 0;
}

}  // namespace b200st
