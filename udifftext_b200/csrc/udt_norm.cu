// udt_norm.cu — K1 GroupNorm(+SiLU) and K6 LayerNorm on NHWC fp16 (HBM-bound kernels, fp32/fp64 statistics).
//
// GroupNorm runs as two coalesced passes over the pixel-major tensor:
//   pass 1 (stats)  : each CTA owns a slab of pixels of one image, every thread keeps per-channel partial
//                     sums for a fixed 8-channel vector (16-byte loads), partials are folded per channel in
//                     shared memory, then per group, and added to an fp64 [image, group] accumulator;
//   pass 2 (apply)  : 16-byte loads, normalise with the group statistics, affine, optional SiLU, 16-byte stores.
// The second read mostly hits the 126 MB L2.  Two sources (x0 ++ x1 on the channel axis) implement the
// UNet skip concatenation without materialising th.cat.
#include "udt_common.cuh"
#include "udt_host.h"

namespace {

using namespace udt;

constexpr int kGnThreads = 256;
constexpr int kGnMaxC = 4096;

struct GnArgs {
  const __half* x0;
  const __half* x1;
  __half* y;
  const float* gamma;
  const float* beta;
  double* stats;  // [NB, groups, 2]
  int C0, C1, C, NB, HW, groups, rows_per_cta, silu;
  float eps;
};

__device__ __forceinline__ uint4 gn_load_vec(const GnArgs& a, size_t pix, int vc) {
  const int c = vc * 8;
  if (c < a.C0) return __ldg(reinterpret_cast<const uint4*>(a.x0 + pix * a.C0 + c));
  return __ldg(reinterpret_cast<const uint4*>(a.x1 + pix * a.C1 + (c - a.C0)));
}

__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const GnArgs a) {
  extern __shared__ float sm[];  // [2*C]
  float* s_sum = sm;
  float* s_sq = sm + a.C;
  const int n = blockIdx.y;
  const int row0 = blockIdx.x * a.rows_per_cta;
  const int row1 = min(a.HW, row0 + a.rows_per_cta);
  for (int i = threadIdx.x; i < 2 * a.C; i += blockDim.x) sm[i] = 0.0f;
  __syncthreads();

  const int VC = a.C / 8;
  const int vcols = min(VC, static_cast<int>(blockDim.x));
  const int rpi = max(1, static_cast<int>(blockDim.x) / VC);
  const int r = threadIdx.x / vcols;
  if (r < rpi) {
    for (int vc = threadIdx.x % vcols; vc < VC; vc += vcols) {
      float s[8], q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.0f;
      for (int row = row0 + r; row < row1; row += rpi) {
        const uint4 v = gn_load_vec(a, static_cast<size_t>(n) * a.HW + row, vc);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          s[2 * j] += f.x;
          q[2 * j] += f.x * f.x;
          s[2 * j + 1] += f.y;
          q[2 * j + 1] += f.y * f.y;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_sum[vc * 8 + j], s[j]);
        atomicAdd(&s_sq[vc * 8 + j], q[j]);
      }
    }
  }
  __syncthreads();
  const int cpg = a.C / a.groups;
  for (int g = threadIdx.x; g < a.groups; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      ds += static_cast<double>(s_sum[c]);
      dq += static_cast<double>(s_sq[c]);
    }
    double* st = a.stats + (static_cast<size_t>(n) * a.groups + g) * 2;
    atomicAdd(st, ds);
    atomicAdd(st + 1, dq);
  }
}

__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const GnArgs a) {
  __shared__ float s_mean[64];
  __shared__ float s_rstd[64];
  const int n = blockIdx.y;
  const int cpg = a.C / a.groups;
  if (threadIdx.x < a.groups) {
    const double* st = a.stats + (static_cast<size_t>(n) * a.groups + threadIdx.x) * 2;
    const double cnt = static_cast<double>(a.HW) * cpg;
    const double mean = st[0] / cnt;
    double var = st[1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
  }
  __syncthreads();
  const int VC = a.C / 8;
  const int row0 = blockIdx.x * a.rows_per_cta;
  const int row1 = min(a.HW, row0 + a.rows_per_cta);
  const int total = (row1 - row0) * VC;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int row = row0 + i / VC;
    const int vc = i % VC;
    const size_t pix = static_cast<size_t>(n) * a.HW + row;
    const uint4 v = gn_load_vec(a, pix, vc);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __half22float2(h[j]);
      f[2 * j] = t.x;
      f[2 * j + 1] = t.y;
    }
    const int c0 = vc * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const int g = c / cpg;
      float o = (f[j] - s_mean[g]) * s_rstd[g] * __ldg(a.gamma + c) + __ldg(a.beta + c);
      if (a.silu) o = silu_f(o);
      f[j] = o;
    }
    uint4 ov;
    ov.x = pack_half2(f[0], f[1]);
    ov.y = pack_half2(f[2], f[3]);
    ov.z = pack_half2(f[4], f[5]);
    ov.w = pack_half2(f[6], f[7]);
    *reinterpret_cast<uint4*>(a.y + pix * a.C + c0) = ov;
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm
constexpr int kLnWarps = 8;
constexpr int kLnMaxVec = 8;  // C <= 32 * 8 * 8 = 2048 kept in registers

__global__ void __launch_bounds__(kLnWarps * 32) layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                                  int rows, int C, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float eps) {
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int VC = C / 8;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * C);
  uint4 v[kLnMaxVec];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int vc = lane + i * 32;
    if (vc < VC) {
      v[i] = __ldg(xr + vc);
      const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        s += f.x + f.y;
      }
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(C);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int vc = lane + i * 32;
    if (vc < VC) {
      const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        const float dx = f.x - mean, dy = f.y - mean;
        q += dx * dx + dy * dy;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(C) + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * C);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int vc = lane + i * 32;
    if (vc < VC) {
      const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
      const int c0 = vc * 8;
      float o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        o[2 * j] = (f.x - mean) * rstd * __ldg(gamma + c0 + 2 * j) + __ldg(beta + c0 + 2 * j);
        o[2 * j + 1] = (f.y - mean) * rstd * __ldg(gamma + c0 + 2 * j + 1) + __ldg(beta + c0 + 2 * j + 1);
      }
      uint4 ov;
      ov.x = pack_half2(o[0], o[1]);
      ov.y = pack_half2(o[2], o[3]);
      ov.z = pack_half2(o[4], o[5]);
      ov.w = pack_half2(o[6], o[7]);
      yr[vc] = ov;
    }
  }
}

}  // namespace

extern "C" int udt_groupnorm_nhwc(const void* x0, int32_t C0, const void* x1, int32_t C1, void* y, int32_t NB,
                                  int32_t HW, int32_t groups, const float* gamma, const float* beta, float eps,
                                  int32_t silu, void* stats_ws, void* stream) {
  using namespace udt_host;
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (x1 == nullptr) C1 = 0;
  const int C = C0 + C1;
  if (C0 % 8 != 0 || C1 % 8 != 0 || C % groups != 0 || groups > 64 || C > kGnMaxC || C < 8)
    return fail(UDT_ERR_SHAPE, "udt_groupnorm_nhwc: C0=%d C1=%d groups=%d unsupported", C0, C1, groups);
  if (NB < 1 || HW < 1) return fail(UDT_ERR_SHAPE, "udt_groupnorm_nhwc: NB=%d HW=%d", NB, HW);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GnArgs a;
  a.x0 = reinterpret_cast<const __half*>(x0);
  a.x1 = reinterpret_cast<const __half*>(x1);
  a.y = reinterpret_cast<__half*>(y);
  a.gamma = gamma;
  a.beta = beta;
  a.stats = reinterpret_cast<double*>(stats_ws);
  a.C0 = C0;
  a.C1 = C1;
  a.C = C;
  a.NB = NB;
  a.HW = HW;
  a.groups = groups;
  a.silu = silu;
  a.eps = eps;
  int rows = 65536 / (C * 2);  // ~64 KB of activations per CTA
  if (rows < 4) rows = 4;
  if (rows > HW) rows = HW;
  a.rows_per_cta = rows;
  const int chunks = (HW + rows - 1) / rows;
  cudaError_t e = cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * NB * groups, st);
  if (e != cudaSuccess) return fail(UDT_ERR_LAUNCH, "udt_groupnorm_nhwc memset: %s", cudaGetErrorString(e));
  dim3 grid(chunks, NB);
  gn_stats_kernel<<<grid, kGnThreads, 2 * C * sizeof(float), st>>>(a);
  gn_apply_kernel<<<grid, kGnThreads, 0, st>>>(a);
  return check_launch("udt_groupnorm_nhwc");
}

extern "C" int udt_layernorm(const void* x, void* y, int32_t rows, int32_t C, const float* gamma, const float* beta,
                             float eps, void* stream) {
  using namespace udt_host;
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (C % 8 != 0 || C > 32 * kLnMaxVec * 8 || C < 8 || rows < 1)
    return fail(UDT_ERR_SHAPE, "udt_layernorm: rows=%d C=%d unsupported (C %% 8 == 0, C <= %d)", rows, C, 32 * kLnMaxVec * 8);
  const int grid = (rows + kLnWarps - 1) / kLnWarps;
  layernorm_kernel<<<grid, kLnWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), rows, C, gamma, beta, eps);
  return check_launch("udt_layernorm");
}
