// udt_norm.cu — K1 GroupNorm(+SiLU) and K6 LayerNorm on NHWC fp16 (HBM-bound kernels, fp32/fp64 statistics).
//
// GroupNorm has two schedules:
//   one pass  (a sample's tensor fits the shared memory of a cluster of <= 8 CTAs, i.e. every UNet level except the
//              64x64 one at >= 320 channels): each CTA of the cluster owns a slab of pixels, loads it ONCE into shared
//              memory while accumulating per-channel sums, the CTAs exchange their per-group partial sums through
//              distributed shared memory in a fixed order (deterministic), then normalise out of shared memory.
//              One launch, 4 B/element of traffic.
//   two pass  (large tensors) as two coalesced passes over the pixel-major tensor:
//   pass 1 (stats)  : each CTA owns a slab of pixels of one image, every thread keeps per-channel partial
//                     sums for a fixed 8-channel vector (16-byte loads), partials are folded per channel in
//                     shared memory, then per group, and added to an fp64 [image, group] accumulator;
//   pass 2 (apply)  : 16-byte loads, normalise with the group statistics, affine, optional SiLU, 16-byte stores.
// The second read mostly hits the 126 MB L2.  Two sources (x0 ++ x1 on the channel axis) implement the
// UNet skip concatenation without materialising th.cat.
#include "udt_common.cuh"
#include "udt_host.h"
#include <stdlib.h>

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
  double* folded;   // [NB, groups, 2] (sum, sum of squares) folded by gn_fold_kernel when there are many chunks, else NULL
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

// SiLU with one MUFU op: x * sigmoid(x) = 0.5 x (1 + tanh(0.5 x)); tanh.approx error (2^-11) is below fp16 resolution
__device__ __forceinline__ float silu_tanh(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return 0.5f * x * (1.0f + t);
}

// fold per-channel partial sums [rpi][C] of group g with `nl` cooperating lanes (lane j of the group), fp64, fixed order
__device__ __forceinline__ void gn_group_fold(const float* ps, const float* pq, int rpi, int C, int cpg, int g, int j, int nl,
                                              double& ds, double& dq) {
  ds = 0.0;
  dq = 0.0;
  for (int rr = 0; rr < rpi; ++rr)
    for (int cc = j; cc < cpg; cc += nl) {
      ds += static_cast<double>(ps[rr * C + g * cpg + cc]);
      dq += static_cast<double>(pq[rr * C + g * cpg + cc]);
    }
  for (int o = nl >> 1; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dq += __shfl_xor_sync(0xffffffffu, dq, o);
  }
}

// pass 1: deterministic per-CTA partial statistics (no atomics anywhere).  The CTA's slab of pixels is fetched with
// bulk asynchronous copies (the whole slab in flight at once, one round trip) and summed out of shared memory.
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const GnArgs a) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  extern __shared__ __align__(16) uint8_t gsm2[];
  __shared__ uint64_t s_bar;
  const int n = blockIdx.y;
  const int row0 = blockIdx.x * a.rows_per_cta;
  const int nrows = min(a.HW, row0 + a.rows_per_cta) - row0;
  const int VC = a.C / 8;
  const int vcols = min(VC, static_cast<int>(blockDim.x));
  const int rpi = max(1, static_cast<int>(blockDim.x) / VC);
  const uint4* slab = reinterpret_cast<const uint4*>(gsm2);
  float* s_sum = reinterpret_cast<float*>(gsm2 + static_cast<size_t>(a.rows_per_cta) * a.C * 2);
  float* s_sq = s_sum + rpi * a.C;
  const size_t pix0 = static_cast<size_t>(n) * a.HW + row0;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  griddep_wait();     // PDL: wait for the producers of our inputs
  if (threadIdx.x == 0) mbar_expect_tx(&s_bar, static_cast<uint32_t>(nrows) * static_cast<uint32_t>(a.C) * 2u);
  if (a.C1 == 0) {
    // single source: the slab is one contiguous range of global memory -> a few large bulk copies (16 KB pieces)
    const uint32_t total = static_cast<uint32_t>(nrows) * static_cast<uint32_t>(a.C) * 2u;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(a.x0 + pix0 * a.C0);
    for (uint32_t off = threadIdx.x * 16384u; off < total; off += blockDim.x * 16384u)
      bulk_load_1d(gsm2 + off, src + off, min(16384u, total - off), &s_bar);
  } else {
    for (int row = threadIdx.x; row < nrows; row += blockDim.x) {
      uint8_t* dst = gsm2 + static_cast<size_t>(row) * a.C * 2;
      bulk_load_1d(dst, a.x0 + (pix0 + row) * a.C0, static_cast<uint32_t>(a.C0) * 2u, &s_bar);
      bulk_load_1d(dst + a.C0 * 2, a.x1 + (pix0 + row) * a.C1, static_cast<uint32_t>(a.C1) * 2u, &s_bar);
    }
  }
  mbar_wait(&s_bar, 0);
  const int r = threadIdx.x / vcols;
  if (r < rpi) {
    for (int vc = threadIdx.x % vcols; vc < VC; vc += vcols) {
      float s[8], q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.0f;
      for (int row = r; row < nrows; row += rpi) gn_accum(slab[row * VC + vc], s, q);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s_sum[r * a.C + vc * 8 + j] = s[j];
        s_sq[r * a.C + vc * 8 + j] = q[j];
      }
    }
  }
  __syncthreads();
  const int cpg = a.C / a.groups;
  {
    const int g = threadIdx.x >> 3, j = threadIdx.x & 7;   // 8 lanes per group (256 threads, <= 32 groups)
    double ds, dq;
    gn_group_fold(s_sum, s_sq, rpi, a.C, cpg, min(g, a.groups - 1), j, 8, ds, dq);
    if (j == 0 && g < a.groups) {
      double* st = a.partial + ((static_cast<size_t>(n) * a.groups + g) * a.chunks + blockIdx.x) * 2;
      st[0] = ds;
      st[1] = dq;
    }
  }
}

// large tensors (VAE resolutions: thousands of chunks per image): fold the per-CTA partials ONCE, in a fixed order, instead
// of in every CTA of the apply pass.  grid (groups, NB), 256 threads.
__global__ void __launch_bounds__(256) gn_fold_kernel(const GnArgs a) {
  griddep_launch();
  griddep_wait();
  __shared__ double s_s[256], s_q[256];
  const int g = blockIdx.x, n = blockIdx.y;
  const double* st = a.partial + (static_cast<size_t>(n) * a.groups + g) * a.chunks * 2;
  double ds = 0.0, dq = 0.0;
  for (int c = threadIdx.x; c < a.chunks; c += 256) {
    ds += st[2 * c];
    dq += st[2 * c + 1];
  }
  s_s[threadIdx.x] = ds;
  s_q[threadIdx.x] = dq;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_s[threadIdx.x] += s_s[threadIdx.x + o];
      s_q[threadIdx.x] += s_q[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a.folded[(static_cast<size_t>(n) * a.groups + g) * 2] = s_s[0];
    a.folded[(static_cast<size_t>(n) * a.groups + g) * 2 + 1] = s_q[0];
  }
}

// pass 2: fold the partials in a fixed order, then y = x * A[c] + B[c] (+SiLU); every thread owns fixed 8-channel
// vectors (its A / B coefficients live in registers) and walks the CTA's rows with 4 loads in flight
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const GnArgs a) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  __shared__ double s_red[kGnMaxGroups][8][2];
  __shared__ float s_mean[kGnMaxGroups];
  __shared__ float s_rstd[kGnMaxGroups];
  const int n = blockIdx.y;
  const int cpg = a.C / a.groups;
  griddep_wait();     // PDL: wait for the producers of our inputs
  {
    const int g = threadIdx.x >> 3, j = threadIdx.x & 7;
    if (g < a.groups) {
      double ds = 0.0, dq = 0.0;
      if (a.folded != nullptr) {            // many chunks: gn_fold_kernel has already summed them
        if (j == 0) {
          ds = a.folded[(static_cast<size_t>(n) * a.groups + g) * 2];
          dq = a.folded[(static_cast<size_t>(n) * a.groups + g) * 2 + 1];
        }
      } else {
        const double* st = a.partial + (static_cast<size_t>(n) * a.groups + g) * a.chunks * 2;
        for (int c = j; c < a.chunks; c += 8) {
          ds += st[2 * c];
          dq += st[2 * c + 1];
        }
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
  const int VC = a.C / 8;
  const int vcols = min(VC, static_cast<int>(blockDim.x));
  const int rpi = max(1, static_cast<int>(blockDim.x) / VC);
  const int r = threadIdx.x / vcols;
  if (r >= rpi) return;
  const int row0 = blockIdx.x * a.rows_per_cta;
  const int nrows = min(a.HW, row0 + a.rows_per_cta) - row0;
  const size_t pix0 = static_cast<size_t>(n) * a.HW + row0;
  for (int vc = threadIdx.x % vcols; vc < VC; vc += vcols) {
    float A[8], B[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = vc * 8 + j;
      const int g = c / cpg;
      A[j] = s_rstd[g] * __ldg(a.gamma + c);
      B[j] = __ldg(a.beta + c) - s_mean[g] * A[j];
    }
    auto finish = [&](const uint4& v, size_t pix) {
      const __half2* h = reinterpret_cast<const __half2*>(&v);
      float o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        o[2 * j] = fmaf(f.x, A[2 * j], B[2 * j]);
        o[2 * j + 1] = fmaf(f.y, A[2 * j + 1], B[2 * j + 1]);
      }
      if (a.silu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = silu_tanh(o[j]);
      }
      uint4 ov;
      ov.x = pack_half2(o[0], o[1]);
      ov.y = pack_half2(o[2], o[3]);
      ov.z = pack_half2(o[4], o[5]);
      ov.w = pack_half2(o[6], o[7]);
      *reinterpret_cast<uint4*>(a.y + pix * a.C + vc * 8) = ov;
    };
    int row = r;
    for (; row + 3 * rpi < nrows; row += 4 * rpi) {
      const uint4 v0 = gn_load_vec(a, pix0 + row, vc);
      const uint4 v1 = gn_load_vec(a, pix0 + row + rpi, vc);
      const uint4 v2 = gn_load_vec(a, pix0 + row + 2 * rpi, vc);
      const uint4 v3 = gn_load_vec(a, pix0 + row + 3 * rpi, vc);
      finish(v0, pix0 + row);
      finish(v1, pix0 + row + rpi);
      finish(v2, pix0 + row + 2 * rpi);
      finish(v3, pix0 + row + 3 * rpi);
    }
    for (; row < nrows; row += rpi) finish(gn_load_vec(a, pix0 + row, vc), pix0 + row);
  }
}

constexpr int kGn1Threads = 512;
constexpr int kGn1MaxCluster = 16;      // 16 = non-portable cluster size (one cluster per GPC); 8 if the device refuses it

// one-pass GroupNorm: grid (S, NB), cluster (S, 1, 1); CTA `rank` of a cluster owns rows [rank*R, rank*R + R) of sample n
__global__ void __launch_bounds__(kGn1Threads) gn_onepass_kernel(const GnArgs a) {
  griddep_launch();
  extern __shared__ __align__(16) uint8_t gsm[];
  __shared__ __align__(16) double s_recv[kGn1MaxCluster][kGnMaxGroups][2];   // per-group partial sums received from every rank
  __shared__ uint64_t s_xbar;
  __shared__ float s_mean[kGnMaxGroups];
  __shared__ float s_rstd[kGnMaxGroups];
  const int n = blockIdx.y;
  const int rank = blockIdx.x;
  const int S = gridDim.x;
  const int R = a.rows_per_cta;
  const int row0 = rank * R;
  const int nrows = max(0, min(a.HW, row0 + R) - row0);
  const int VC = a.C / 8;                       // <= 512 (C <= 4096)
  const int rpi = kGn1Threads / VC;             // row lanes
  const int r = threadIdx.x / VC;
  const int vc = threadIdx.x - r * VC;
  const bool act = r < rpi;
  uint4* slab = reinterpret_cast<uint4*>(gsm);                       // [R][VC] 16-byte vectors
  float* ps = reinterpret_cast<float*>(gsm + static_cast<size_t>(R) * a.C * 2);   // [rpi][C]
  float* pq = ps + rpi * a.C;
  const int cpg = a.C / a.groups;
  const size_t pix0 = static_cast<size_t>(n) * a.HW + row0;

  // the slab is fetched with bulk asynchronous copies (one per row and source): the whole slab is in flight at once
  __shared__ uint64_t s_bar;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_init(&s_xbar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  cluster_arrive_relaxed();   // "my exchange barrier exists"; waited for just before the first remote store
  griddep_wait();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&s_bar, static_cast<uint32_t>(nrows) * static_cast<uint32_t>(a.C) * 2u);
    mbar_expect_tx(&s_xbar, static_cast<uint32_t>(S) * static_cast<uint32_t>(a.groups) * 16u);
  }
  for (int row = threadIdx.x; row < nrows; row += kGn1Threads) {
    uint8_t* dst = gsm + static_cast<size_t>(row) * a.C * 2;
    bulk_load_1d(dst, a.x0 + (pix0 + row) * a.C0, static_cast<uint32_t>(a.C0) * 2u, &s_bar);
    if (a.C1 > 0) bulk_load_1d(dst + a.C0 * 2, a.x1 + (pix0 + row) * a.C1, static_cast<uint32_t>(a.C1) * 2u, &s_bar);
  }
  mbar_wait(&s_bar, 0);
  if (act) {
    float sv[8], qv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sv[j] = qv[j] = 0.0f;
    for (int row = r; row < nrows; row += rpi) gn_accum(slab[row * VC + vc], sv, qv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ps[r * a.C + vc * 8 + j] = sv[j];
      pq[r * a.C + vc * 8 + j] = qv[j];
    }
  }
  __syncthreads();
  {
    const int g = threadIdx.x >> 4, j = threadIdx.x & 15;   // 16 lanes per group (512 threads, <= 32 groups)
    double ds, dq;
    gn_group_fold(ps, pq, rpi, a.C, cpg, min(g, a.groups - 1), j, 16, ds, dq);
    cluster_wait();                                  // every CTA of the cluster has initialised its exchange barrier
    if (j == 0 && g < a.groups) {
      // push this CTA's partial sums of group g to every CTA of the cluster (itself included): asynchronous 16-byte
      // stores counted on the receiver's mbarrier — no cluster-wide memory fence, no second cluster barrier
      const uint32_t dst = smem_u32(&s_recv[rank][g][0]);
      const uint32_t bar = smem_u32(&s_xbar);
      for (int rk = 0; rk < S; ++rk)
        st_async_f64x2(mapa_u32(dst, static_cast<uint32_t>(rk)), ds, dq, mapa_u32(bar, static_cast<uint32_t>(rk)));
    }
  }
  mbar_wait(&s_xbar, 0);                             // all S contributions have landed in s_recv
  if (threadIdx.x < a.groups) {
    double ds = 0.0, dq = 0.0;
    for (int rk = 0; rk < S; ++rk) {               // fixed order over the cluster: deterministic
      ds += s_recv[rk][threadIdx.x][0];
      dq += s_recv[rk][threadIdx.x][1];
    }
    const double cnt = static_cast<double>(a.HW) * cpg;
    const double mean = ds / cnt;
    double var = dq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
  }
  __syncthreads();      // publishes s_mean / s_rstd (peers only ever WRITE to this CTA, and all of those writes have landed)
  if (act) {
    float A[8], B[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = vc * 8 + j;
      const int g = c / cpg;
      A[j] = s_rstd[g] * __ldg(a.gamma + c);
      B[j] = __ldg(a.beta + c) - s_mean[g] * A[j];
    }
    for (int row = r; row < nrows; row += rpi) {
      const uint4 v = slab[row * VC + vc];
      const __half2* h = reinterpret_cast<const __half2*>(&v);
      float o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        o[2 * j] = fmaf(f.x, A[2 * j], B[2 * j]);
        o[2 * j + 1] = fmaf(f.y, A[2 * j + 1], B[2 * j + 1]);
      }
      if (a.silu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = silu_tanh(o[j]);
      }
      uint4 ov;
      ov.x = pack_half2(o[0], o[1]);
      ov.y = pack_half2(o[2], o[3]);
      ov.z = pack_half2(o[4], o[5]);
      ov.w = pack_half2(o[6], o[7]);
      *reinterpret_cast<uint4*>(a.y + (pix0 + row) * a.C + vc * 8) = ov;
    }
  }
}

// ---------------------------------------------------------------------------------------------- group-owner GroupNorm
// Small / mid-size activations (the whole tensor is L2 resident: every UNet level at batch <= 8): one CTA owns ONE
// (sample, group) — its HW x cpg elements — so the statistics need no exchange between CTAs at all: no cluster, no
// DSMEM, no second kernel.  NB x 32 CTAs (256 at batch 4) fill the machine where a cluster per sample used 64 SMs.  The
// group's elements are short channel runs (cpg * 2 bytes per pixel, e.g. 20 B at 320 channels) strided by the pixel
// pitch; the neighbouring groups' CTAs run at the same time, so the partially used sectors are shared through L2.
// The slab is kept in shared memory between the statistics and the apply pass when it fits (else re-read from L2).
// Deterministic: fixed thread -> element mapping, shuffle tree, fp64 fold of the per-warp sums in warp order.
constexpr int kGgThreads = 480;   // a multiple of every vectors-per-pixel count in use (5, 10, 15, 2, 4, 8, ...): see below
constexpr int kGgMaxCpg = 128;
constexpr int kGgSlabBudget = 200 * 1024;

template <typename V>   // uint32_t / uint2 / uint4 = 2 / 4 / 8 channels per vector
__global__ void __launch_bounds__(kGgThreads, 2) gn_group_kernel(const GnArgs a, const int nvec, const int use_smem) {
  griddep_launch();
  constexpr int VW = static_cast<int>(sizeof(V)) / 2;
  constexpr int U = sizeof(V) == 16 ? 4 : 8;   // independent loads in flight per thread
  extern __shared__ __align__(16) uint8_t ggs[];
  __shared__ double s_ws[kGgThreads / 32][2];
  __shared__ float s_stat[2];
  V* slab = reinterpret_cast<V*>(ggs);
  const int g = blockIdx.x, n = blockIdx.y;
  const int cpg = a.C / a.groups;
  // 480 % nvec == 0: a thread always handles the same vector slot j of its pixels (pixels pl, pl + ppi, ...), so its
  // affine coefficients live in registers and the addresses advance by a constant
  const int ppi = kGgThreads / nvec;
  const int pl = static_cast<int>(threadIdx.x) / nvec, j = static_cast<int>(threadIdx.x) - pl * nvec;
  const int c = g * cpg + j * VW;
  const bool first = c < a.C0;
  const size_t pitch = first ? a.C0 : a.C1;
  const __half* src = (first ? a.x0 + c : a.x1 + (c - a.C0)) + (static_cast<size_t>(n) * a.HW + pl) * pitch;
  const size_t step = static_cast<size_t>(ppi) * pitch;
  const int niter = pl < a.HW ? (a.HW - pl + ppi - 1) / ppi : 0;
  griddep_wait();

  float sum = 0.0f, sq = 0.0f;
  for (int it = 0; it < niter; it += U) {
    V v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (it + u < niter) v[u] = __ldg(reinterpret_cast<const V*>(src + (it + u) * step));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (it + u < niter) {
        if (use_smem) slab[(it + u) * kGgThreads + threadIdx.x] = v[u];
        const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
        for (int k = 0; k < VW / 2; ++k) {
          const float2 f = __half22float2(h[k]);
          sum += f.x + f.y;
          sq = fmaf(f.x, f.x, fmaf(f.y, f.y, sq));
        }
      }
    }
  }
  sum = warp_sum(sum);
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) {
    s_ws[threadIdx.x >> 5][0] = static_cast<double>(sum);
    s_ws[threadIdx.x >> 5][1] = static_cast<double>(sq);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ds = 0.0, dq = 0.0;
    for (int w = 0; w < kGgThreads / 32; ++w) {   // fixed order
      ds += s_ws[w][0];
      dq += s_ws[w][1];
    }
    const double cnt = static_cast<double>(a.HW) * cpg;
    const double mean = ds / cnt;
    double var = dq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_stat[0] = static_cast<float>(mean);
    s_stat[1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
  }
  __syncthreads();
  float A[VW], B[VW];
  if (niter > 0) {
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      A[k] = s_stat[1] * __ldg(a.gamma + c + k);
      B[k] = __ldg(a.beta + c + k) - s_stat[0] * A[k];
    }
  }
  __half* dst = a.y + (static_cast<size_t>(n) * a.HW + pl) * a.C + c;
  const size_t dstep = static_cast<size_t>(ppi) * a.C;
  for (int it = 0; it < niter; it += U) {
    V v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (it + u < niter)
        v[u] = use_smem ? slab[(it + u) * kGgThreads + threadIdx.x] : __ldg(reinterpret_cast<const V*>(src + (it + u) * step));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (it + u < niter) {
        const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
        V o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int k = 0; k < VW / 2; ++k) {
          const float2 f = __half22float2(h[k]);
          float y0 = fmaf(f.x, A[2 * k], B[2 * k]), y1 = fmaf(f.y, A[2 * k + 1], B[2 * k + 1]);
          if (a.silu) {
            y0 = silu_tanh(y0);
            y1 = silu_tanh(y1);
          }
          ow[k] = pack_half2(y0, y1);
        }
        *reinterpret_cast<V*>(dst + (it + u) * dstep) = o;
      }
    }
  }
}

template <typename V>
int launch_gn_group(const GnArgs& a, int vw, cudaStream_t st) {
  const int cpg = a.C / a.groups;
  const int nvec = cpg / vw;
  const int ppi = kGgThreads / nvec;
  // slab slots: ceil(HW / ppi) rounds of kGgThreads vectors
  const size_t slab = static_cast<size_t>((a.HW + ppi - 1) / ppi) * kGgThreads * sizeof(V);
  const int use_smem = slab <= static_cast<size_t>(kGgSlabBudget);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gn_group_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGgSlabBudget);
    if (e != cudaSuccess) return udt_host::fail(UDT_ERR_LAUNCH, "cudaFuncSetAttribute(gn group smem): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  udt_host::launch_pdl(gn_group_kernel<V>, dim3(a.groups, a.NB), dim3(kGgThreads), use_smem ? slab : 0, st, a, nvec, use_smem);
  return udt_host::check_launch("udt_groupnorm_nhwc (group owner)");
}

constexpr int kGn1SlabBudget = 176 * 1024;   // + <= 32 KB of partial sums (<= 208 KB dynamic) + ~9 KB static < 227 KB

// cluster size of the one-pass schedule for this problem, 0 = use the two-pass schedule
inline int gn_onepass_cluster(int NB, int HW, int C, int max_cluster) {
  if (C / 8 > kGn1Threads) return 0;
  int s_fit = 0;
  for (int s = 1; s <= max_cluster; s <<= 1) {
    const long slab = static_cast<long>((HW + s - 1) / s) * C * 2;
    if (slab <= kGn1SlabBudget) {
      s_fit = s;
      break;
    }
  }
  if (s_fit == 0) return 0;
  int s = s_fit;
  while (s < 8 && NB * s < 96 && 2 * s <= HW) s <<= 1;   // spread small problems over more SMs (portable sizes only)
  return s;
}

inline int gn_rows_per_cta(int C, int HW) {
  int rows = 49152 / (C * 2);  // ~48 KB of activations per CTA
  if (rows < 4) rows = 4;
  if (rows > HW) rows = HW;
  return rows;
}

// ---------------------------------------------------------------------------------------------- LayerNorm
constexpr int kLnWarps = 8;
// (gamma / beta live in registers, so layernorm_rows_kernel runs ONE CTA per SM and its memory-level parallelism has to come
//  from the row passes U a warp keeps in flight: U = 4 when that still fills the GPU, else U = 2 — measured on B200:
//  32768 x 320 9.3 -> 8.1 us, 8192 x 640 5.5 -> 4.7 us, 262144 x 320 63.5 -> 56.1 us; 4096 x 640 would go 3.3 -> 4.7 us)
constexpr int kLnMaxVec = 8;  // C <= 32 * 8 * 8 = 2048 kept in registers

__global__ void __launch_bounds__(kLnWarps * 32) layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                                  int rows, int C, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float eps) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
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


// LayerNorm, sub-warp-per-row schedule: LPR lanes share one row, each lane owns VPL 16-byte vectors of it (C = 8*VPL*LPR)
// and keeps their gamma / beta in registers; a warp normalises 32/LPR rows per pass and two passes are in flight, the
// warps walk the rows grid-stride.  (C = 320 / 640 / 1280 -> LPR = 8 / 16 / 32 with VPL = 5; C = 2048 -> VPL = 8.)
template <int VPL, int LPR, int U>
__global__ void __launch_bounds__(kLnWarps * 32) layernorm_rows_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                                       int rows, const float* __restrict__ gamma,
                                                                       const float* __restrict__ beta, float eps) {
  griddep_launch();
  constexpr int RPW = 32 / LPR;
  constexpr int VC = VPL * LPR;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / LPR, j = lane % LPR;
  float g[VPL][8], b[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c0 = (j + i * LPR) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    g[i][0] = g0.x; g[i][1] = g0.y; g[i][2] = g0.z; g[i][3] = g0.w; g[i][4] = g1.x; g[i][5] = g1.y; g[i][6] = g1.z; g[i][7] = g1.w;
    b[i][0] = b0.x; b[i][1] = b0.y; b[i][2] = b0.z; b[i][3] = b0.w; b[i][4] = b1.x; b[i][5] = b1.y; b[i][6] = b1.z; b[i][7] = b1.w;
  }
  griddep_wait();
  const float inv_c = 1.0f / static_cast<float>(VC * 8);
  const int stride = gridDim.x * kLnWarps * RPW * U;
  for (int row0 = (blockIdx.x * kLnWarps + warp) * RPW * U; row0 < rows; row0 += stride) {
    uint4 v[U][VPL];
    int row[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      row[u] = row0 + u * RPW + sub;
      const uint4* xr = reinterpret_cast<const uint4*>(x) + static_cast<size_t>(min(row[u], rows - 1)) * VC + j;
#pragma unroll
      for (int i = 0; i < VPL; ++i) v[u][i] = __ldg(xr + i * LPR);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const __half2* h = reinterpret_cast<const __half2*>(&v[u][i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(h[k]);
          s += f.x + f.y;
        }
      }
#pragma unroll
      for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * inv_c;
      float q = 0.0f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const __half2* h = reinterpret_cast<const __half2*>(&v[u][i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(h[k]);
          const float dx = f.x - mean, dy = f.y - mean;
          q = fmaf(dx, dx, q);
          q = fmaf(dy, dy, q);
        }
      }
#pragma unroll
      for (int o = LPR >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q * inv_c + eps);
      if (row[u] < rows) {
        uint4* yr = reinterpret_cast<uint4*>(y) + static_cast<size_t>(row[u]) * VC + j;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const __half2* h = reinterpret_cast<const __half2*>(&v[u][i]);
          float o8[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(h[k]);
            o8[2 * k] = fmaf((f.x - mean) * rstd, g[i][2 * k], b[i][2 * k]);
            o8[2 * k + 1] = fmaf((f.y - mean) * rstd, g[i][2 * k + 1], b[i][2 * k + 1]);
          }
          uint4 ov;
          ov.x = pack_half2(o8[0], o8[1]);
          ov.y = pack_half2(o8[2], o8[3]);
          ov.z = pack_half2(o8[4], o8[5]);
          ov.w = pack_half2(o8[6], o8[7]);
          yr[i * LPR] = ov;
        }
      }
    }
  }
}

template <int VPL, int LPR>
void launch_ln_rows(const void* x, void* y, int rows, const float* gamma, const float* beta, float eps, cudaStream_t st) {
  constexpr int RPW = 32 / LPR;
  const int cap = udt_host::num_sms() * 3;
  const int need4 = (rows + kLnWarps * RPW * 4 - 1) / (kLnWarps * RPW * 4);
  if (VPL <= 5 && need4 >= 100) {   // four passes in flight still give (almost) every SM a CTA
    udt_host::launch_pdl(layernorm_rows_kernel<VPL, LPR, (VPL <= 5 ? 4 : 2)>, dim3(need4 < cap ? need4 : cap), dim3(kLnWarps * 32), 0, st,
                         reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), rows, gamma, beta, eps);
    return;
  }
  const int need = (rows + kLnWarps * RPW * 2 - 1) / (kLnWarps * RPW * 2);
  udt_host::launch_pdl(layernorm_rows_kernel<VPL, LPR, 2>, dim3(need < cap ? need : cap), dim3(kLnWarps * 32), 0, st,
                       reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), rows, gamma, beta, eps);
}

}  // namespace

extern "C" int64_t udt_groupnorm_ws_bytes(int32_t NB, int32_t HW, int32_t C, int32_t groups) {
  if (NB < 1 || HW < 1 || C < 8 || groups < 1) return 0;
  const int rows = gn_rows_per_cta(C, HW);
  const int chunks = (HW + rows - 1) / rows;
  return static_cast<int64_t>(NB) * groups * (chunks + 1) * 2 * static_cast<int64_t>(sizeof(double));   // + folded sums
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
  a.folded = nullptr;
  a.C0 = C0;
  a.C1 = C1;
  a.C = C;
  a.NB = NB;
  a.HW = HW;
  a.groups = groups;
  a.silu = silu;
  a.eps = eps;
  static const bool onepass_ok = udt_host::tune_int("UDT_GN_ONEPASS", 1) != 0;
  // largest usable cluster: 16 CTAs (non-portable size) when the device can co-schedule such a cluster with the
  // kernel's full shared-memory footprint, else the portable 8
  static int max_cluster = 0;
  if (max_cluster == 0) {
    cudaError_t e = cudaFuncSetAttribute(gn_onepass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    if (e != cudaSuccess) return fail(UDT_ERR_LAUNCH, "cudaFuncSetAttribute(gn one-pass smem): %s", cudaGetErrorString(e));
    max_cluster = 8;
    // opt-in (tuning builds): measured SLOWER than the two-pass schedule on B200 (8 co-resident clusters of 16 CTAs pull a
    // 2.6 MB sample at single-SM rates: 64x64x320 28 us vs 19 us), kept for experiments
    static const bool big_ok = udt_host::tune_int("UDT_GN_CLUSTER16", 0) != 0;
    if (big_ok && cudaFuncSetAttribute(gn_onepass_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3(16, 1, 1);
      q.blockDim = dim3(kGn1Threads, 1, 1);
      q.dynamicSmemBytes = 208 * 1024;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 16;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, gn_onepass_kernel, &q) == cudaSuccess && nclusters >= 4) max_cluster = 16;
    }
    (void)cudaGetLastError();
  }
  // group-owner schedule while the whole tensor is L2 resident (UDT_GN_GROUP=0: never)
  static const bool group_ok = udt_host::tune_int("UDT_GN_GROUP", 1) != 0;
  // tensor size up to which the group-owner schedule is used (tuning builds: UDT_GN_GROUP_MB)
  static const size_t group_limit = static_cast<size_t>(udt_host::tune_int("UDT_GN_GROUP_MB", 64)) << 20;
  {
    const int cpg = C / groups;
    const size_t bytes = static_cast<size_t>(NB) * HW * C * 2;
    const int vw = cpg % 8 == 0 ? 8 : (cpg % 4 == 0 ? 4 : 2);
    // measured on B200 (batch 4): wins where a pixel contributes a run of >= 40 bytes and the slab fits shared memory twice per
    // SM (32x32x640 13.2 -> 9.9 us, 16x16x1280 9.4 -> 5.0 us, 8x8x1280 7.1 -> 3.5 us); loses at 64x64x320 (20-byte runs:
    // 19.5 -> 23.4 us), which keeps the two-pass schedule
    const size_t slab = static_cast<size_t>(HW) * cpg * 2;
    if (group_ok && cpg % 2 == 0 && cpg * 2 >= 40 && cpg <= kGgMaxCpg && kGgThreads % (cpg / vw) == 0 && slab <= 128 * 1024 &&
        bytes <= group_limit) {
      if (vw == 8) return launch_gn_group<uint4>(a, 8, st);
      if (vw == 4) return launch_gn_group<uint2>(a, 4, st);
      return launch_gn_group<uint32_t>(a, 2, st);
    }
  }
  const int S = onepass_ok ? gn_onepass_cluster(NB, HW, C, max_cluster) : 0;
  if (S > 0) {
    a.rows_per_cta = (HW + S - 1) / S;
    a.chunks = S;
    const int rpi1 = kGn1Threads / (C / 8);
    const size_t smem1 = static_cast<size_t>(a.rows_per_cta) * C * 2 + static_cast<size_t>(2) * rpi1 * C * sizeof(float);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(S, NB, 1);
    cfg.blockDim = dim3(kGn1Threads, 1, 1);
    cfg.dynamicSmemBytes = smem1;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = S;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    int na = 1;
    na += pdl_attr(&attrs[na]);
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gn_onepass_kernel, a);
    if (e != cudaSuccess) return fail(UDT_ERR_LAUNCH, "udt_groupnorm_nhwc (one pass): %s", cudaGetErrorString(e));
    return UDT_OK;
  }
  a.rows_per_cta = gn_rows_per_cta(C, HW);
  a.chunks = (HW + a.rows_per_cta - 1) / a.rows_per_cta;
  const int VC = C / 8;
  const int rpi = (kGnThreads / VC) > 1 ? (kGnThreads / VC) : 1;
  const int smem_stats = a.rows_per_cta * C * 2 + 2 * rpi * C * static_cast<int>(sizeof(float));
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    attr_set = true;
  }
  dim3 grid(a.chunks, NB);
  a.folded = nullptr;
  udt_host::launch_pdl(gn_stats_kernel, dim3(grid), dim3(kGnThreads), smem_stats, st, a);
  if (a.chunks > 64) {
    a.folded = a.partial + static_cast<size_t>(NB) * groups * a.chunks * 2;
    udt_host::launch_pdl(gn_fold_kernel, dim3(groups, NB), dim3(256), 0, st, a);
  }
  udt_host::launch_pdl(gn_apply_kernel, dim3(grid), dim3(kGnThreads), 0, st, a);
  return check_launch("udt_groupnorm_nhwc");
}

extern "C" int udt_layernorm(const void* x, void* y, int32_t rows, int32_t C, const float* gamma, const float* beta,
                             float eps, void* stream) {
  using namespace udt_host;
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (C % 8 != 0 || C > 32 * kLnMaxVec * 8 || C < 8 || rows < 1)
    return fail(UDT_ERR_SHAPE, "udt_layernorm: rows=%d C=%d unsupported (C %% 8 == 0, C <= %d)", rows, C, 32 * kLnMaxVec * 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int VC = C / 8;
  const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                    reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
  if (al && (VC == 40 || VC == 80 || VC == 160 || VC == 256)) {
    if (VC == 40) launch_ln_rows<5, 8>(x, y, rows, gamma, beta, eps, st);
    else if (VC == 80) launch_ln_rows<5, 16>(x, y, rows, gamma, beta, eps, st);
    else if (VC == 160) launch_ln_rows<5, 32>(x, y, rows, gamma, beta, eps, st);
    else launch_ln_rows<8, 32>(x, y, rows, gamma, beta, eps, st);
    return check_launch("udt_layernorm");
  }
  const int grid = (rows + kLnWarps - 1) / kLnWarps;
  udt_host::launch_pdl(layernorm_kernel, dim3(grid), dim3(kLnWarps * 32), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), rows, C, gamma, beta, eps);
  return check_launch("udt_layernorm");
}
