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
constexpr int kGnMaxGroups = 32;

struct GnArgs {
  const __half* x0;
  const __half* x1;
  __half* y;
  const float* gamma;
  const float* beta;
  double* partial;  // [NB, groups, chunks, 2] per-CTA partial (sum, sum of squares)
  int C0, C1, C, NB, HW, groups, rows_per_cta, chunks, silu;
  float eps;
};

__device__ __forceinline__ uint4 gn_load_vec(const GnArgs& a, size_t pix, int vc) {
  const int c = vc * 8;
  if (c < a.C0) return __ldg(reinterpret_cast<const uint4*>(a.x0 + pix * a.C0 + c));
  return __ldg(reinterpret_cast<const uint4*>(a.x1 + pix * a.C1 + (c - a.C0)));
}

__device__ __forceinline__ void gn_accum(const uint4& v, float (&s)[8], float (&q)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h[j]);
    s[2 * j] += f.x;
    q[2 * j] = fmaf(f.x, f.x, q[2 * j]);
    s[2 * j + 1] += f.y;
    q[2 * j + 1] = fmaf(f.y, f.y, q[2 * j + 1]);
  }
}

// pass 1: deterministic per-CTA partial statistics (no atomics anywhere)
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const GnArgs a) {
  extern __shared__ float sm[];  // [rpi][C] sums, then [rpi][C] sums of squares
  const int n = blockIdx.y;
  const int row0 = blockIdx.x * a.rows_per_cta;
  const int row1 = min(a.HW, row0 + a.rows_per_cta);
  const int VC = a.C / 8;
  const int vcols = min(VC, static_cast<int>(blockDim.x));
  const int rpi = max(1, static_cast<int>(blockDim.x) / VC);
  float* s_sum = sm;
  float* s_sq = sm + rpi * a.C;
  const int r = threadIdx.x / vcols;
  if (r < rpi) {
    const size_t pix0 = static_cast<size_t>(n) * a.HW;
    for (int vc = threadIdx.x % vcols; vc < VC; vc += vcols) {
      float s[8], q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.0f;
      int row = row0 + r;
      for (; row + 3 * rpi < row1; row += 4 * rpi) {  // 4 independent 16-byte loads in flight
        const uint4 v0 = gn_load_vec(a, pix0 + row, vc);
        const uint4 v1 = gn_load_vec(a, pix0 + row + rpi, vc);
        const uint4 v2 = gn_load_vec(a, pix0 + row + 2 * rpi, vc);
        const uint4 v3 = gn_load_vec(a, pix0 + row + 3 * rpi, vc);
        gn_accum(v0, s, q);
        gn_accum(v1, s, q);
        gn_accum(v2, s, q);
        gn_accum(v3, s, q);
      }
      for (; row < row1; row += rpi) gn_accum(gn_load_vec(a, pix0 + row, vc), s, q);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s_sum[r * a.C + vc * 8 + j] = s[j];
        s_sq[r * a.C + vc * 8 + j] = q[j];
      }
    }
  }
  __syncthreads();
  const int cpg = a.C / a.groups;
  for (int g = threadIdx.x; g < a.groups; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    for (int rr = 0; rr < rpi; ++rr)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        ds += static_cast<double>(s_sum[rr * a.C + c]);
        dq += static_cast<double>(s_sq[rr * a.C + c]);
      }
    double* st = a.partial + ((static_cast<size_t>(n) * a.groups + g) * a.chunks + blockIdx.x) * 2;
    st[0] = ds;
    st[1] = dq;
  }
}

// pass 2: fold the partials in a fixed order, then y = x * A[c] + B[c] (+SiLU) with per-channel A/B in smem
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const GnArgs a) {
  extern __shared__ float sm[];  // A[C], B[C]
  __shared__ double s_red[kGnMaxGroups][8][2];
  __shared__ float s_mean[kGnMaxGroups];
  __shared__ float s_rstd[kGnMaxGroups];
  float* sA = sm;
  float* sB = sm + a.C;
  const int n = blockIdx.y;
  const int cpg = a.C / a.groups;
  {
    const int g = threadIdx.x >> 3, j = threadIdx.x & 7;
    if (g < a.groups) {
      const double* st = a.partial + (static_cast<size_t>(n) * a.groups + g) * a.chunks * 2;
      double ds = 0.0, dq = 0.0;
      for (int c = j; c < a.chunks; c += 8) {
        ds += st[2 * c];
        dq += st[2 * c + 1];
      }
      s_red[g][j][0] = ds;
      s_red[g][j][1] = dq;
    }
  }
  __syncthreads();
  if (threadIdx.x < a.groups) {
    double ds = 0.0, dq = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ds += s_red[threadIdx.x][j][0];
      dq += s_red[threadIdx.x][j][1];
    }
    const double cnt = static_cast<double>(a.HW) * cpg;
    const double mean = ds / cnt;
    double var = dq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    const int g = c / cpg;
    const float A = s_rstd[g] * __ldg(a.gamma + c);
    sA[c] = A;
    sB[c] = __ldg(a.beta + c) - s_mean[g] * A;
  }
  __syncthreads();
  const int VC = a.C / 8;
  const int row0 = blockIdx.x * a.rows_per_cta;
  const int row1 = min(a.HW, row0 + a.rows_per_cta);
  const int total = (row1 - row0) * VC;
  const size_t pix0 = static_cast<size_t>(n) * a.HW + row0;
  constexpr int U = 4;
  for (int i0 = threadIdx.x; i0 < total; i0 += U * kGnThreads) {
    uint4 v[U];
    int vcs[U];
    size_t pix[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * kGnThreads;
      if (i < total) {
        const int row = i / VC;
        vcs[u] = i - row * VC;
        pix[u] = pix0 + row;
        v[u] = gn_load_vec(a, pix[u], vcs[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * kGnThreads;
      if (i < total) {
        const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
        const int c0 = vcs[u] * 8;
        const float4 A0 = *reinterpret_cast<const float4*>(sA + c0);
        const float4 A1 = *reinterpret_cast<const float4*>(sA + c0 + 4);
        const float4 B0 = *reinterpret_cast<const float4*>(sB + c0);
        const float4 B1 = *reinterpret_cast<const float4*>(sB + c0 + 4);
        const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]);
        const float2 f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
        float o[8];
        o[0] = fmaf(f0.x, A0.x, B0.x);
        o[1] = fmaf(f0.y, A0.y, B0.y);
        o[2] = fmaf(f1.x, A0.z, B0.z);
        o[3] = fmaf(f1.y, A0.w, B0.w);
        o[4] = fmaf(f2.x, A1.x, B1.x);
        o[5] = fmaf(f2.y, A1.y, B1.y);
        o[6] = fmaf(f3.x, A1.z, B1.z);
        o[7] = fmaf(f3.y, A1.w, B1.w);
        if (a.silu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = __fdividef(o[j], 1.0f + __expf(-o[j]));
        }
        uint4 ov;
        ov.x = pack_half2(o[0], o[1]);
        ov.y = pack_half2(o[2], o[3]);
        ov.z = pack_half2(o[4], o[5]);
        ov.w = pack_half2(o[6], o[7]);
        *reinterpret_cast<uint4*>(a.y + pix[u] * a.C + c0) = ov;
      }
    }
  }
}

inline int gn_rows_per_cta(int C, int HW) {
  int rows = 65536 / (C * 2);  // ~64 KB of activations per CTA
  if (rows < 4) rows = 4;
  if (rows > HW) rows = HW;
  return rows;
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

extern "C" int64_t udt_groupnorm_ws_bytes(int32_t NB, int32_t HW, int32_t C, int32_t groups) {
  if (NB < 1 || HW < 1 || C < 8 || groups < 1) return 0;
  const int rows = gn_rows_per_cta(C, HW);
  const int chunks = (HW + rows - 1) / rows;
  return static_cast<int64_t>(NB) * groups * chunks * 2 * static_cast<int64_t>(sizeof(double));
}

extern "C" int udt_groupnorm_nhwc(const void* x0, int32_t C0, const void* x1, int32_t C1, void* y, int32_t NB,
                                  int32_t HW, int32_t groups, const float* gamma, const float* beta, float eps,
                                  int32_t silu, void* stats_ws, void* stream) {
  using namespace udt_host;
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (x1 == nullptr) C1 = 0;
  const int C = C0 + C1;
  if (C0 % 8 != 0 || C1 % 8 != 0 || C % groups != 0 || groups > kGnMaxGroups || C > kGnMaxC || C < 8)
    return fail(UDT_ERR_SHAPE, "udt_groupnorm_nhwc: C0=%d C1=%d groups=%d unsupported", C0, C1, groups);
  if (NB < 1 || HW < 1) return fail(UDT_ERR_SHAPE, "udt_groupnorm_nhwc: NB=%d HW=%d", NB, HW);
  if (stats_ws == nullptr || (reinterpret_cast<uintptr_t>(stats_ws) & 7)) return fail(UDT_ERR_ALIGN, "udt_groupnorm_nhwc: workspace");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GnArgs a;
  a.x0 = reinterpret_cast<const __half*>(x0);
  a.x1 = reinterpret_cast<const __half*>(x1);
  a.y = reinterpret_cast<__half*>(y);
  a.gamma = gamma;
  a.beta = beta;
  a.partial = reinterpret_cast<double*>(stats_ws);
  a.C0 = C0;
  a.C1 = C1;
  a.C = C;
  a.NB = NB;
  a.HW = HW;
  a.groups = groups;
  a.silu = silu;
  a.eps = eps;
  a.rows_per_cta = gn_rows_per_cta(C, HW);
  a.chunks = (HW + a.rows_per_cta - 1) / a.rows_per_cta;
  const int VC = C / 8;
  const int rpi = (kGnThreads / VC) > 1 ? (kGnThreads / VC) : 1;
  const int smem_stats = 2 * rpi * C * static_cast<int>(sizeof(float));
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kGnMaxC * 4 * 2);
    attr_set = true;
  }
  dim3 grid(a.chunks, NB);
  gn_stats_kernel<<<grid, kGnThreads, smem_stats, st>>>(a);
  gn_apply_kernel<<<grid, kGnThreads, 2 * C * sizeof(float), st>>>(a);
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
