// udt_elem.cu — HBM-bound glue kernels of the hot path: K5 short-context cross-attention, row softmax,
// K7 CFG pack / Euler step, nearest 2x upsample, the LabelEncoder / PARSeq small-sequence attentions and the
// NCHW fp32 <-> NHWC fp16 conversions at the API boundary.  All 16-byte vectorised where the layout allows.
#include "udt_common.cuh"
#include "udt_host.h"

namespace {

using namespace udt;

// ---------------------------------------------------------------------------------------------- K5
// One thread per (query pixel, head): q row (64 fp16 = 128 B) against L <= 16 context tokens whose K/V
// (L x 64 each) are staged in shared memory as fp32.  Replaces attention.py:147-174.
constexpr int kXaMaxL = 16;
constexpr int kXaThreads = 128;

__global__ void __launch_bounds__(kXaThreads) xattn_small_l_kernel(const __half* __restrict__ q,
                                                                   const __half* __restrict__ kc,
                                                                   const __half* __restrict__ vc,
                                                                   __half* __restrict__ o, float* __restrict__ probs,
                                                                   int N, int L, int heads, int ldq, int ldkv, int ldo,
                                                                   float scale) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
  __shared__ float sk[kXaMaxL * 64];
  __shared__ float sv[kXaMaxL * 64];
  const int b = blockIdx.z;
  const int h = blockIdx.y;
  for (int i = threadIdx.x; i < L * 64; i += blockDim.x) {
    const int l = i / 64, d = i % 64;
    const size_t off = (static_cast<size_t>(b) * L + l) * ldkv + h * 64 + d;
    sk[i] = __half2float(kc[off]);
    sv[i] = __half2float(vc[off]);
  }
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const size_t row = static_cast<size_t>(b) * N + n;
  float qf[64];
  const uint4* q4 = reinterpret_cast<const uint4*>(q + row * ldq + h * 64);
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const uint4 t = __ldg(q4 + v);
    const __half2* hh = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(hh[j]);
      qf[v * 8 + 2 * j] = f.x;
      qf[v * 8 + 2 * j + 1] = f.y;
    }
  }
  float s[kXaMaxL];
  float mx = -INFINITY;
#pragma unroll
  for (int l = 0; l < kXaMaxL; ++l) {
    if (l < L) {
      float acc = 0.0f;
#pragma unroll
      for (int d = 0; d < 64; ++d) acc += qf[d] * sk[l * 64 + d];
      s[l] = acc * scale;
      mx = fmaxf(mx, s[l]);
    }
  }
  if (L > 1) {
    float sum = 0.0f;
#pragma unroll
    for (int l = 0; l < kXaMaxL; ++l)
      if (l < L) {
        s[l] = __expf(s[l] - mx);
        sum += s[l];
      }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int l = 0; l < kXaMaxL; ++l)
      if (l < L) s[l] *= inv;
  } else {
    s[0] = 1.0f / (1.0f + __expf(-s[0]));  // sigmoid on a single token (attention.py:159-162)
  }
  if (probs != nullptr) {
    float* pr = probs + ((static_cast<size_t>(b) * heads + h) * N + n) * L;
    for (int l = 0; l < L; ++l) pr[l] = s[l];
  }
  float of[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) of[d] = 0.0f;
#pragma unroll
  for (int l = 0; l < kXaMaxL; ++l) {
    if (l < L) {
#pragma unroll
      for (int d = 0; d < 64; ++d) of[d] += s[l] * sv[l * 64 + d];
    }
  }
  uint4* o4 = reinterpret_cast<uint4*>(o + row * ldo + h * 64);
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    uint4 ov;
    ov.x = pack_half2(of[v * 8 + 0], of[v * 8 + 1]);
    ov.y = pack_half2(of[v * 8 + 2], of[v * 8 + 3]);
    ov.z = pack_half2(of[v * 8 + 4], of[v * 8 + 5]);
    ov.w = pack_half2(of[v * 8 + 6], of[v * 8 + 7]);
    o4[v] = ov;
  }
}

// ---------------------------------------------------------------------------------------------- folded cross-attention
// W1[s][h*L + l][c] = scale * sum_d K[s,l,h*64+d] * Wq[h*64+d][c];  W2[s][c][h*L + l] = sum_d Wo[c][h*64+d] * V[s,l,h*64+d]
// (once per request: the context is step-invariant).  One thread per output element, 64-term dot products.
__global__ void __launch_bounds__(256) xattn_fold_kernel(const __half* __restrict__ kc, const __half* __restrict__ vc, int ldkv,
                                                         const __half* __restrict__ wq, int ldwq,
                                                         const __half* __restrict__ wo, int ldwo, __half* __restrict__ w1,
                                                         __half* __restrict__ w2, int B, int L, int heads, int Npad,
                                                         float scale) {
  griddep_launch();
  griddep_wait();
  const int C = heads * 64;
  const long long per = static_cast<long long>(Npad) * C;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= 2 * per * B) return;
  const bool second = idx >= per * B;
  const long long i = second ? idx - per * B : idx;
  const int s = static_cast<int>(i / per);
  const int r = static_cast<int>(i - s * per);
  if (!second) {                       // W1[s][row = h*L + l][c]
    const int row = r / C, c = r - row * C;
    float acc = 0.0f;
    if (row < heads * L) {
      const int h = row / L, l = row - h * L;
      const __half* kp = kc + (static_cast<size_t>(s) * L + l) * ldkv + h * 64;
      for (int d = 0; d < 64; ++d) acc = fmaf(__half2float(kp[d]), __half2float(wq[static_cast<size_t>(h * 64 + d) * ldwq + c]), acc);
      acc *= scale;
    }
    w1[i] = __float2half_rn(acc);
  } else {                             // W2[s][c][col = h*L + l]
    const int c = r / Npad, col = r - c * Npad;
    float acc = 0.0f;
    if (col < heads * L) {
      const int h = col / L, l = col - h * L;
      const __half* vp = vc + (static_cast<size_t>(s) * L + l) * ldkv + h * 64;
      const __half* wp = wo + static_cast<size_t>(c) * ldwo + h * 64;
      for (int d = 0; d < 64; ++d) acc = fmaf(__half2float(wp[d]), __half2float(vp[d]), acc);
    }
    w2[i] = __float2half_rn(acc);
  }
}

// softmax over `groups` groups of L <= 16 consecutive columns of every row; one thread per (row, group)
__global__ void __launch_bounds__(256) softmax_groups_kernel(const __half* __restrict__ in, __half* __restrict__ out, int rows,
                                                             int cols, int ld, int groups, int L, float* __restrict__ probs,
                                                             int N) {
  griddep_launch();
  griddep_wait();
  const int gpr = groups + 1;          // the extra "group" of a row zeroes the pad columns
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * gpr) return;
  const int row = static_cast<int>(idx / gpr);
  const int g = static_cast<int>(idx - static_cast<long long>(row) * gpr);
  if (g == groups) {
    for (int c = groups * L; c < cols; ++c) out[static_cast<size_t>(row) * ld + c] = __float2half_rn(0.0f);
    return;
  }
  const __half* x = in + static_cast<size_t>(row) * ld + g * L;
  float v[16];
  float mx = -INFINITY;
  const bool vec = (L == 12) && ((ld & 3) == 0);     // 12 tokens = three 8-byte words (consecutive threads: consecutive words)
  if (vec) {
    const uint2* x2 = reinterpret_cast<const uint2*>(x);
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      const uint2 t = x2[w];
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
      v[4 * w] = a.x; v[4 * w + 1] = a.y; v[4 * w + 2] = b.x; v[4 * w + 3] = b.y;
    }
#pragma unroll
    for (int l = 12; l < 16; ++l) v[l] = -INFINITY;
#pragma unroll
    for (int l = 0; l < 12; ++l) mx = fmaxf(mx, v[l]);
  } else {
#pragma unroll
    for (int l = 0; l < 16; ++l) {
      v[l] = (l < L) ? __half2float(x[l]) : -INFINITY;
      mx = fmaxf(mx, v[l]);
    }
  }
  float sum = 0.0f;
  if (L > 1) {
#pragma unroll
    for (int l = 0; l < 16; ++l) {
      v[l] = (l < L) ? __expf(v[l] - mx) : 0.0f;
      sum += v[l];
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int l = 0; l < 16; ++l) v[l] *= inv;
  } else {
    v[0] = 1.0f / (1.0f + __expf(-v[0]));   // sigmoid on a single token (attention.py:159-162)
  }
  __half* y = out + static_cast<size_t>(row) * ld + g * L;
  if (vec) {
    uint2* y2 = reinterpret_cast<uint2*>(y);
#pragma unroll
    for (int w = 0; w < 3; ++w) y2[w] = make_uint2(pack_half2(v[4 * w], v[4 * w + 1]), pack_half2(v[4 * w + 2], v[4 * w + 3]));
  } else {
#pragma unroll
    for (int l = 0; l < 16; ++l)
      if (l < L) y[l] = __float2half_rn(v[l]);
  }
  if (probs != nullptr) {
    const int img = row / N, n = row - img * N;
    float* pr = probs + ((static_cast<size_t>(img) * groups + g) * N + n) * L;
#pragma unroll
    for (int l = 0; l < 16; ++l)
      if (l < L) pr[l] = v[l];
  }
}

// ---------------------------------------------------------------------------------------------- LabelEncoder pieces
// Character embedding + sinusoid positional encoding (encoders/modules.py:1160-1166, 1083-1085):
// out fp16 [B*L, D] = emb[idx[b,l], :] + pe[l, :]
__global__ void label_embed_kernel(const int32_t* __restrict__ idx, const float* __restrict__ emb,
                                   const float* __restrict__ pe, __half* __restrict__ out, float* __restrict__ out_f32,
                                   __half* __restrict__ out_lo, int rows, int L, int D) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
  const size_t total = static_cast<size_t>(rows) * D;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const int r = static_cast<int>(i / D);
    const float v = emb[static_cast<size_t>(idx[r]) * D + d] + pe[static_cast<size_t>(r % L) * D + d];
    const __half hi = __float2half_rn(v);
    out[i] = hi;
    if (out_f32 != nullptr) out_f32[i] = v;
    if (out_lo != nullptr) out_lo[i] = __float2half_rn(v - __half2float(hi));
  }
}

// ---- fp32-stream pieces of the LabelEncoder (its output conditions every step of every request, so its error is a
// systematic one: it is evaluated with an fp32 residual stream and hi + lo fp16 operand pairs, see label.py)
// y = LN( relu?( in0 + in1 + in2 ) + res ) per row (LN optional); y -> fp32 and / or the fp16 pair hi = rn(y), lo = rn(y - hi)
constexpr int kRsThreads = 256;
constexpr int kRsMaxPer = 32;   // C <= 256 * 32 = 8192

__global__ void __launch_bounds__(kRsThreads) rowsum_norm_split_kernel(
    const float* __restrict__ in0, const float* __restrict__ in1, const float* __restrict__ in2, const float* __restrict__ res,
    int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu, float* __restrict__ out_f32,
    __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  griddep_launch();
  griddep_wait();
  __shared__ float red[kRsThreads / 32];
  __shared__ float bc;
  const size_t base = static_cast<size_t>(blockIdx.x) * C;
  float v[kRsMaxPer];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < kRsMaxPer; ++i) {
    const int c = threadIdx.x + i * kRsThreads;
    float t = 0.0f;
    if (c < C) {
      t = in0[base + c];
      if (in1 != nullptr) t += in1[base + c];
      if (in2 != nullptr) t += in2[base + c];
      if (relu) t = fmaxf(t, 0.0f);
      if (res != nullptr) t += res[base + c];
    }
    v[i] = t;
    s += t;
  }
  auto block_sum = [&](float x) -> float {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.0f;
      for (int w = 0; w < kRsThreads / 32; ++w) t += red[w];
      bc = t;
    }
    __syncthreads();
    return bc;
  };
  float mean = 0.0f, rstd = 1.0f;
  if (gamma != nullptr) {
    mean = block_sum(s) / static_cast<float>(C);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < kRsMaxPer; ++i) {
      const int c = threadIdx.x + i * kRsThreads;
      if (c < C) {
        const float d = v[i] - mean;
        q += d * d;
      }
    }
    rstd = rsqrtf(block_sum(q) / static_cast<float>(C) + eps);
  }
#pragma unroll
  for (int i = 0; i < kRsMaxPer; ++i) {
    const int c = threadIdx.x + i * kRsThreads;
    if (c < C) {
      float y = v[i];
      if (gamma != nullptr) y = (y - mean) * rstd * gamma[c] + beta[c];
      if (out_f32 != nullptr) out_f32[base + c] = y;
      if (out_hi != nullptr) {
        const __half hi = __float2half_rn(y);
        out_hi[base + c] = hi;
        if (out_lo != nullptr) out_lo[base + c] = __float2half_rn(y - __half2float(hi));
      }
    }
  }
}

// fp32 variant of mha_small: qkv fp32 [B*L, ld] (q | k | v), output as the fp16 pair hi / lo
__global__ void __launch_bounds__(128) mha_small_f32_kernel(const float* __restrict__ qkv, __half* __restrict__ o_hi,
                                                            __half* __restrict__ o_lo, int L, int heads, int dh, int ld, int ldo,
                                                            float scale) {
  griddep_launch();
  griddep_wait();
  extern __shared__ float smf[];
  float* sq = smf;
  float* sk = sq + L * dh;
  float* sv = sk + L * dh;
  float* sp = sv + L * dh;
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = heads * dh;
  for (int i = threadIdx.x; i < L * dh; i += blockDim.x) {
    const int l = i / dh, d = i % dh;
    const size_t off = (static_cast<size_t>(b) * L + l) * ld + h * dh + d;
    sq[i] = qkv[off];
    sk[i] = qkv[off + D];
    sv[i] = qkv[off + 2 * D];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
    const int i = e / L, j = e % L;
    float acc = 0.0f;
    for (int d = 0; d < dh; ++d) acc = fmaf(sq[i * dh + d], sk[j * dh + d], acc);
    sp[e] = acc * scale;
  }
  __syncthreads();
  if (threadIdx.x < L) {
    float* row = sp + threadIdx.x * L;
    float mx = -INFINITY;
    for (int j = 0; j < L; ++j) mx = fmaxf(mx, row[j]);
    float sum = 0.0f;
    for (int j = 0; j < L; ++j) {
      row[j] = expf(row[j] - mx);
      sum += row[j];
    }
    const float inv = 1.0f / sum;
    for (int j = 0; j < L; ++j) row[j] *= inv;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < L * dh; e += blockDim.x) {
    const int i = e / dh, d = e % dh;
    float acc = 0.0f;
    for (int j = 0; j < L; ++j) acc = fmaf(sp[i * L + j], sv[j * dh + d], acc);
    const size_t off = (static_cast<size_t>(b) * L + i) * ldo + h * dh + d;
    const __half hi = __float2half_rn(acc);
    o_hi[off] = hi;
    o_lo[off] = __float2half_rn(acc - __half2float(hi));
  }
}

// Multi-head self-attention over a short sequence (L <= 16 tokens, head dim <= 256, no mask): the
// nn.TransformerEncoderLayer attention of the LabelEncoder (encoders/modules.py:1103-1104).  One CTA per
// (head, batch item); q/k/v are column blocks of the fused in_proj output [B*L, 3*D].
constexpr int kMhaMaxL = 16;
constexpr int kMhaMaxD = 256;
constexpr int kMhaThreads = 128;

__global__ void __launch_bounds__(kMhaThreads) mha_small_kernel(const __half* __restrict__ qkv, __half* __restrict__ o,
                                                                int L, int heads, int dh, int ld, int ldo, float scale) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
  __shared__ __half sq[kMhaMaxL * kMhaMaxD];
  __shared__ __half sk[kMhaMaxL * kMhaMaxD];
  __shared__ __half sv[kMhaMaxL * kMhaMaxD];
  __shared__ float sp[kMhaMaxL * kMhaMaxL];
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = heads * dh;
  for (int i = threadIdx.x; i < L * dh; i += blockDim.x) {
    const int l = i / dh, d = i % dh;
    const size_t off = (static_cast<size_t>(b) * L + l) * ld + h * dh + d;
    sq[i] = qkv[off];
    sk[i] = qkv[off + D];
    sv[i] = qkv[off + 2 * D];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
    const int i = e / L, j = e % L;
    float acc = 0.0f;
    for (int d = 0; d < dh; ++d) acc = fmaf(__half2float(sq[i * dh + d]), __half2float(sk[j * dh + d]), acc);
    sp[e] = acc * scale;
  }
  __syncthreads();
  if (threadIdx.x < L) {
    float* row = sp + threadIdx.x * L;
    float mx = -INFINITY;
    for (int j = 0; j < L; ++j) mx = fmaxf(mx, row[j]);
    float sum = 0.0f;
    for (int j = 0; j < L; ++j) {
      row[j] = __expf(row[j] - mx);
      sum += row[j];
    }
    const float inv = 1.0f / sum;
    for (int j = 0; j < L; ++j) row[j] *= inv;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < L * dh; e += blockDim.x) {
    const int i = e / dh, d = e % dh;
    float acc = 0.0f;
    for (int j = 0; j < L; ++j) acc = fmaf(sp[i * L + j], __half2float(sv[j * dh + d]), acc);
    o[(static_cast<size_t>(b) * L + i) * ldo + h * dh + d] = __float2half_rn(acc);
  }
}

// ---------------------------------------------------------------------------------------------- masked small MHA
// nn.MultiheadAttention over short sequences with an additive attention mask and a boolean key-padding mask: the self- and
// cross-attention of PARSeq's two-stream decoder layer (src/parseq/strhub/models/parseq/modules.py:57-75; OCR scoring,
// test.py:58-91).  One CTA per (head, batch item): this head's K / V rows staged in shared memory as fp32 (row pitch dh + 1:
// conflict-free when the lanes walk the keys), one warp per query row: lanes stride over the keys for the scores, warp
// softmax, lanes stride over the head dim for P V.  A fully masked row yields zeros.
constexpr int kMmThreads = 128;
constexpr int kMmMaxLk = 160;
constexpr int kMmMaxDh = 64;

__global__ void __launch_bounds__(kMmThreads) mha_masked_kernel(const __half* __restrict__ q, const __half* __restrict__ k,
                                                                const __half* __restrict__ v, __half* __restrict__ o, int Lq, int Lk,
                                                                int dh, int ldq, int ldk, int ldv, int ldo, float scale,
                                                                const float* __restrict__ mask, int ldm,
                                                                const uint8_t* __restrict__ kpm) {
  griddep_launch();
  griddep_wait();
  extern __shared__ float smm[];
  const int pitch = dh + 1;
  float* sk = smm;
  float* sv = sk + Lk * pitch;
  float* sq = sv + Lk * pitch;                 // [warps][dh]
  float* sp = sq + (kMmThreads / 32) * dh;     // [warps][Lk]
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < Lk * dh; i += blockDim.x) {
    const int j = i / dh, d = i - j * dh;
    sk[j * pitch + d] = __half2float(k[(static_cast<size_t>(b) * Lk + j) * ldk + h * dh + d]);
    sv[j * pitch + d] = __half2float(v[(static_cast<size_t>(b) * Lk + j) * ldv + h * dh + d]);
  }
  __syncthreads();
  float* myq = sq + warp * dh;
  float* myp = sp + warp * Lk;
  for (int i = warp; i < Lq; i += kMmThreads / 32) {
    const size_t qrow = static_cast<size_t>(b) * Lq + i;
    for (int d = lane; d < dh; d += 32) myq[d] = __half2float(q[qrow * ldq + h * dh + d]);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < Lk; j += 32) {
      float acc = 0.0f;
      for (int d = 0; d < dh; ++d) acc = fmaf(myq[d], sk[j * pitch + d], acc);
      acc *= scale;
      if (mask != nullptr) acc += mask[static_cast<size_t>(i) * ldm + j];
      if (kpm != nullptr && kpm[static_cast<size_t>(b) * Lk + j]) acc = -INFINITY;
      myp[j] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int j = lane; j < Lk; j += 32) {
      const float e = (mx == -INFINITY) ? 0.0f : __expf(myp[j] - mx);
      myp[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float acc = 0.0f;
      for (int j = 0; j < Lk; ++j) acc = fmaf(myp[j], sv[j * pitch + d], acc);
      o[qrow * ldo + h * dh + d] = __float2half_rn(acc * inv);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------- row softmax
// One CTA per row; cols up to 16384 (VAE attention N = 4096 / 9216), values cached in registers.
constexpr int kSmThreads = 256;
constexpr int kSmMaxVec = 8;  // 256 threads * 8 vec * 8 halves = 16384 columns

__global__ void __launch_bounds__(kSmThreads) softmax_rows_kernel(__half* __restrict__ x, int cols, int ld, float scale) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
  __shared__ float red[kSmThreads / 32];
  __half* xr = x + static_cast<size_t>(blockIdx.x) * ld;
  const int VC = cols / 8;
  uint4 v[kSmMaxVec];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kSmMaxVec; ++i) {
    const int vc = threadIdx.x + i * kSmThreads;
    if (vc < VC) {
      v[i] = *reinterpret_cast<const uint4*>(xr + vc * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        mx = fmaxf(mx, fmaxf(f.x, f.y));
      }
    }
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < kSmThreads / 32; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float e[kSmMaxVec][8];
  float sum = 0.0f;
  const float sl2 = scale * 1.4426950408889634f;
#pragma unroll
  for (int i = 0; i < kSmMaxVec; ++i) {
    const int vc = threadIdx.x + i * kSmThreads;
    if (vc < VC) {
      const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        e[i][2 * j] = exp2f((f.x - mx) * sl2);
        e[i][2 * j + 1] = exp2f((f.y - mx) * sl2);
        sum += e[i][2 * j] + e[i][2 * j + 1];
      }
    }
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kSmThreads / 32; ++i) sum += red[i];
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < kSmMaxVec; ++i) {
    const int vc = threadIdx.x + i * kSmThreads;
    if (vc < VC) {
      uint4 ov;
      ov.x = pack_half2(e[i][0] * inv, e[i][1] * inv);
      ov.y = pack_half2(e[i][2] * inv, e[i][3] * inv);
      ov.z = pack_half2(e[i][4] * inv, e[i][5] * inv);
      ov.w = pack_half2(e[i][6] * inv, e[i][7] * inv);
      *reinterpret_cast<uint4*>(xr + vc * 8) = ov;
    }
  }
}

// ---------------------------------------------------------------------------------------------- K7
// Sampler state stays in the reference's layout (x fp32 NCHW [B,4,HW], concat fp32 NCHW [B,5,HW]); the UNet side
// is NHWC.  Per-step scalars (c_in, sigma_next - sigma) are read from device memory so that one captured CUDA
// graph serves every step.
// unet_in[2B, HW, 16] fp16: rows [0,B) = uc half, [B,2B) = cond half (guiders.py:36: uc first).
__global__ void cfg_pack_kernel(const float* __restrict__ x, const float* __restrict__ cat_uc,
                                const float* __restrict__ cat_c, __half* __restrict__ out, int B, int HW,
                                const float* __restrict__ c_in_dev) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over 2B*HW pixels
  const int total = 2 * B * HW;
  if (idx >= total) return;
  const float c_in = __ldg(c_in_dev);
  const int half_sel = idx / (B * HW);
  const int pix = idx % (B * HW);
  const int b = pix / HW, p = pix % HW;
  const float* xb = x + static_cast<size_t>(b) * 4 * HW + p;
  const float* cc = (half_sel == 0 ? cat_uc : cat_c) + static_cast<size_t>(b) * 5 * HW + p;
  float f[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) f[j] = __ldg(xb + static_cast<size_t>(j) * HW) * c_in;
#pragma unroll
  for (int j = 0; j < 5; ++j) f[4 + j] = __ldg(cc + static_cast<size_t>(j) * HW);
#pragma unroll
  for (int j = 9; j < 16; ++j) f[j] = 0.0f;
  uint4* o4 = reinterpret_cast<uint4*>(out + static_cast<size_t>(idx) * 16);
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    uint4 ov;
    ov.x = pack_half2(f[v * 8 + 0], f[v * 8 + 1]);
    ov.y = pack_half2(f[v * 8 + 2], f[v * 8 + 3]);
    ov.z = pack_half2(f[v * 8 + 4], f[v * 8 + 5]);
    ov.w = pack_half2(f[v * 8 + 6], f[v * 8 + 7]);
    o4[v] = ov;
  }
}

// x[B,4,HW] (NCHW) += dsigma * (eps_u + s*(eps_c - eps_u)); eps2b fp32 NHWC [2B,HW,4]
__global__ void cfg_euler_kernel(float* __restrict__ x, const float* __restrict__ eps, int B, int HW, float s,
                                 const float* __restrict__ dsigma_dev, const float* __restrict__ scale_dev) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = B * HW;
  if (idx >= total) return;
  const float dsigma = __ldg(dsigma_dev);
  if (scale_dev != nullptr) s = __ldg(scale_dev);   // per-request guidance scale from device memory (graph-stable)
  const int b = idx / HW, p = idx % HW;
  const float4 eu = *reinterpret_cast<const float4*>(eps + static_cast<size_t>(idx) * 4);
  const float4 ec = *reinterpret_cast<const float4*>(eps + (static_cast<size_t>(total) + idx) * 4);
  float* xb = x + static_cast<size_t>(b) * 4 * HW + p;
  xb[0] += dsigma * (eu.x + s * (ec.x - eu.x));
  xb[static_cast<size_t>(HW)] += dsigma * (eu.y + s * (ec.y - eu.y));
  xb[static_cast<size_t>(2) * HW] += dsigma * (eu.z + s * (ec.z - eu.z));
  xb[static_cast<size_t>(3) * HW] += dsigma * (eu.w + s * (ec.w - eu.w));
}

// ---------------------------------------------------------------------------------------------- K10
// Conditioner tail: posterior sample of the masked-image latent for the c and the uc branch (two different
// noise draws over the same moments), latent scale, 1/8 bilinear mask, channel concat.
//   moments fp32 NHWC [B, HW, ldm] (mean 0..3, logvar 4..7); noise fp32 NCHW [B,4,HW]; mask fp32 [B,1,8h,8w]
//   concat_{c,uc} fp32 NCHW [B,5,HW] = cat(mask8, scale * (mean + exp(0.5*clamp(logvar,-30,20)) * noise))
__global__ void vae_sample_pack_kernel(const float* __restrict__ moments, int ldm, const float* __restrict__ noise_c,
                                       const float* __restrict__ noise_uc, const float* __restrict__ mask,
                                       float* __restrict__ cat_c, float* __restrict__ cat_uc, int B, int h, int w,
                                       float scale) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = h * w;
  if (idx >= B * HW) return;
  const int b = idx / HW, p = idx % HW;
  const int y = p / w, x = p % w;
  // F.interpolate(scale_factor=0.125, bilinear, align_corners=False): source coordinate 8*i + 3.5
  const int W8 = 8 * w;
  const float* mb = mask + static_cast<size_t>(b) * (8 * h) * W8 + static_cast<size_t>(8 * y + 3) * W8 + (8 * x + 3);
  const float m8 = 0.25f * (mb[0] + mb[1] + mb[W8] + mb[W8 + 1]);
  const float* mo = moments + static_cast<size_t>(idx) * ldm;
  const size_t o5 = static_cast<size_t>(b) * 5 * HW + p;
  const size_t o4 = static_cast<size_t>(b) * 4 * HW + p;
  cat_c[o5] = m8;
  cat_uc[o5] = m8;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float mean = mo[c];
    const float lv = fminf(fmaxf(mo[4 + c], -30.0f), 20.0f);
    const float sd = expf(0.5f * lv);
    cat_c[o5 + static_cast<size_t>(c + 1) * HW] = scale * (mean + sd * noise_c[o4 + static_cast<size_t>(c) * HW]);
    cat_uc[o5 + static_cast<size_t>(c + 1) * HW] = scale * (mean + sd * noise_uc[o4 + static_cast<size_t>(c) * HW]);
  }
}

// per-pixel small affine map (post_quant_conv, autoencoder.py:313-316 with the 1/scale_factor of
// diffusion.py:125 folded into `in_scale`): out fp16 NHWC [B,HW,Cpad] = Wm[Cout,Cin] * (x[B,Cin,HW] * in_scale) + bias
__global__ void pointwise_affine_kernel(const float* __restrict__ x, const float* __restrict__ Wm,
                                        const float* __restrict__ bias, __half* __restrict__ out, int B, int HW, int Cin,
                                        int Cout, int Cpad, float in_scale) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * HW) return;
  const int b = idx / HW, p = idx % HW;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = (k < Cin) ? x[(static_cast<size_t>(b) * Cin + k) * HW + p] * in_scale : 0.0f;
  __half* o = out + static_cast<size_t>(idx) * Cpad;
  for (int c = 0; c < Cpad; ++c) {
    float acc = 0.0f;
    if (c < Cout) {
      acc = bias != nullptr ? __ldg(bias + c) : 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < Cin) acc = fmaf(__ldg(Wm + c * Cin + k), v[k], acc);
    }
    o[c] = __float2half_rn(acc);
  }
}

// ---------------------------------------------------------------------------------------------- movement
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int NB, int H, int W, int VC) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
  const size_t total = static_cast<size_t>(NB) * (2 * H) * (2 * W) * VC;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int vc = static_cast<int>(i % VC);
    size_t r = i / VC;
    const int ox = static_cast<int>(r % (2 * W));
    r /= (2 * W);
    const int oy = static_cast<int>(r % (2 * H));
    const int n = static_cast<int>(r / (2 * H));
    y[i] = __ldg(x + ((static_cast<size_t>(n) * H + (oy >> 1)) * W + (ox >> 1)) * VC + vc);
  }
}


__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __half* __restrict__ y, int NB, int C, int HW, int Cpad) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
  const size_t total = static_cast<size_t>(NB) * HW * Cpad;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cpad);
    const size_t r = i / Cpad;
    const int p = static_cast<int>(r % HW);
    const int n = static_cast<int>(r / HW);
    y[i] = (c < C) ? __float2half_rn(x[(static_cast<size_t>(n) * C + c) * HW + p]) : __float2half(0.0f);
  }
}

__global__ void nhwc_to_nchw_kernel(const void* __restrict__ x, int is_f32, float* __restrict__ y, int NB, int C, int HW,
                                    int ld, float scale, float shift, int clamp01) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
  const size_t total = static_cast<size_t>(NB) * C * HW;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i % HW);
    const size_t r = i / HW;
    const int c = static_cast<int>(r % C);
    const int n = static_cast<int>(r / C);
    const size_t src = (static_cast<size_t>(n) * HW + p) * ld + c;
    float v = is_f32 ? reinterpret_cast<const float*>(x)[src] : __half2float(reinterpret_cast<const __half*>(x)[src]);
    v = v * scale + shift;
    if (clamp01) v = fminf(fmaxf(v, 0.0f), 1.0f);
    y[i] = v;
  }
}

inline int grid_for(size_t total, int threads) {
  size_t g = (total + threads - 1) / threads;
  const size_t cap = 148 * 32;
  return static_cast<int>(g < cap ? (g == 0 ? 1 : g) : cap);
}

// ---------------------------------------------------------------------------------------------- request front-end
// demo.py:52-62 on the device: uint8 HWC request image + uint8 HWC user mask (0 = keep, anything else = synthesise) ->
// image fp32 NCHW in [-1, 1], m = mean_c(mask == 0), masked = image * m, mask = 1 - m; source s = b % Bs is tiled over the
// batch (demo.py:78-80).  Same fp32 operations in the same order as the reference's torch expressions: bit-exact.
__global__ void request_pack_u8_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ msk, float* __restrict__ image,
                                       float* __restrict__ mask, float* __restrict__ masked, int B, int Bs, int HW, int MC) {
  griddep_launch();
  griddep_wait();
  const size_t total = static_cast<size_t>(B) * HW;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i % HW);
    const int b = static_cast<int>(i / HW);
    const size_t sp = static_cast<size_t>(b % Bs) * HW + p;
    float m = 0.0f;
    for (int c = 0; c < MC; ++c) m += (msk[sp * MC + c] == 0) ? 1.0f : 0.0f;
    m = __fdiv_rn(m, static_cast<float>(MC));
    mask[i] = 1.0f - m;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fsub_rn(__fdiv_rn(static_cast<float>(img[sp * 3 + c]), 127.5f), 1.0f);
      const size_t o = (static_cast<size_t>(b) * 3 + c) * HW + p;
      image[o] = v;
      masked[o] = __fmul_rn(v, m);
    }
  }
}

// demo.py:100-101 / test.py:94: fp32 NCHW in [0, 1] -> uint8 NHWC, (x * 255) truncated like numpy's astype(uint8)
__global__ void images_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, int NB, int C, int HW) {
  griddep_launch();
  griddep_wait();
  const size_t total = static_cast<size_t>(NB) * HW * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t r = i / C;
    const int p = static_cast<int>(r % HW);
    const int n = static_cast<int>(r / HW);
    const float v = __fmul_rn(x[(static_cast<size_t>(n) * C + c) * HW + p], 255.0f);
    y[i] = static_cast<uint8_t>(static_cast<int>(fminf(fmaxf(v, 0.0f), 255.0f)));
  }
}

// ---------------------------------------------------------------------------------------------- K12
// Noise-search score of one t_attn layer (loss.py:192-235 get_min_local_loss): one CTA per UNet sample b.  For every
// valid token l the head-mean attention map is gathered into shared memory, blurred with the ks x ks Gaussian (zero
// padding, cross-correlation like F.conv2d), multiplied by the nearest-sampled inpainting mask and max-reduced over the
// pixels; score[b] += -min_l (max + 1 - seg[l]).  max / min are order independent and the head sum runs in a fixed
// order, so the result is deterministic; the layers are accumulated in launch order on the stream.
constexpr int kLsThreads = 512;
constexpr int kLsMaxK = 7;

__global__ void __launch_bounds__(kLsThreads) attn_local_score_kernel(const float* __restrict__ probs,
                                                                      const float* __restrict__ mask,
                                                                      const float* __restrict__ seg,
                                                                      const float* __restrict__ gk, float* __restrict__ score,
                                                                      int Bm, int heads, int size, int L, int seg_l, int H, int W,
                                                                      int ks) {
  griddep_launch();
  griddep_wait();
  extern __shared__ float ls_sm[];
  __shared__ float red[kLsThreads / 32];
  __shared__ float s_gk[kLsMaxK * kLsMaxK];
  const int N = size * size;
  float* hm = ls_sm;        // [N] head-mean attention of the current token
  float* mk = ls_sm + N;    // [N] mask sampled at the map resolution
  const int b = blockIdx.x;
  const int bm = b % Bm;    // CFG-doubled UNet batch [uc; c]: both halves score against the same per-image mask
  const int pad = ks / 2;
  const float sy = static_cast<float>(H) / static_cast<float>(size), sx = static_cast<float>(W) / static_cast<float>(size);
  if (threadIdx.x < ks * ks) s_gk[threadIdx.x] = gk[threadIdx.x];
  for (int n = threadIdx.x; n < N; n += kLsThreads) {
    const int y = n / size, x = n - y * size;
    const int yy = min(static_cast<int>(floorf(y * sy)), H - 1), xx = min(static_cast<int>(floorf(x * sx)), W - 1);
    mk[n] = mask[(static_cast<size_t>(bm) * H + yy) * W + xx];
  }
  const float inv_heads = 1.0f / static_cast<float>(heads);
  float best = INFINITY;
  for (int l = 0; l < seg_l; ++l) {
    __syncthreads();   // previous token's hm fully consumed (and mk / s_gk written, first time round)
    for (int n = threadIdx.x; n < N; n += kLsThreads) {
      float acc = 0.0f;
      for (int h = 0; h < heads; ++h) acc += probs[((static_cast<size_t>(b) * heads + h) * N + n) * L + l];
      hm[n] = acc * inv_heads;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int n = threadIdx.x; n < N; n += kLsThreads) {
      const int y = n / size, x = n - y * size;
      float v = 0.0f;
      for (int i = 0; i < ks; ++i) {
        const int yy = y + i - pad;
        if (yy < 0 || yy >= size) continue;
        for (int j = 0; j < ks; ++j) {
          const int xx = x + j - pad;
          if (xx >= 0 && xx < size) v = fmaf(s_gk[i * ks + j], hm[yy * size + xx], v);
        }
      }
      mx = fmaxf(mx, mk[n] * v);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < kLsThreads / 32; ++i) mx = fmaxf(mx, red[i]);
      best = fminf(best, mx + (1.0f - seg[bm * seg_l + l]));
    }
  }
  if (threadIdx.x == 0) score[b] += -best;
}

}  // namespace

using udt_host::check_launch;
using udt_host::fail;
using udt_host::require_sm100;

extern "C" int udt_xattn_small_l(const void* q, const void* kc, const void* vc, void* o, float* probs, int32_t B,
                                 int32_t N, int32_t L, int32_t heads, int32_t ldq, int32_t ldkv, int32_t ldo,
                                 float scale, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (L < 1 || L > kXaMaxL) return fail(UDT_ERR_SHAPE, "udt_xattn_small_l: L=%d (1..%d)", L, kXaMaxL);
  if (ldq % 8 || ldo % 8) return fail(UDT_ERR_ALIGN, "udt_xattn_small_l: ldq/ldo must be multiples of 8");
  dim3 grid((N + kXaThreads - 1) / kXaThreads, heads, B);
  udt_host::launch_pdl(xattn_small_l_kernel, dim3(grid), dim3(kXaThreads), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __half*>(q), reinterpret_cast<const __half*>(kc), reinterpret_cast<const __half*>(vc),
      reinterpret_cast<__half*>(o), probs, N, L, heads, ldq, ldkv, ldo, scale);
  return check_launch("udt_xattn_small_l");
}

extern "C" int udt_label_embed(const int32_t* idx, const float* emb, const float* pe, void* out, float* out_f32, void* out_lo,
                               int32_t rows, int32_t L, int32_t D, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (rows < 1 || L < 1 || D < 1) return fail(UDT_ERR_SHAPE, "udt_label_embed: rows=%d L=%d D=%d", rows, L, D);
  const size_t total = static_cast<size_t>(rows) * D;
  udt_host::launch_pdl(label_embed_kernel, dim3(grid_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
      idx, emb, pe, reinterpret_cast<__half*>(out), out_f32, reinterpret_cast<__half*>(out_lo), rows, L, D);
  return check_launch("udt_label_embed");
}

extern "C" int udt_rowsum_norm_split(const float* in0, const float* in1, const float* in2, const float* res, int32_t rows,
                                     int32_t C, const float* gamma, const float* beta, float eps, int32_t relu, float* out_f32,
                                     void* out_hi, void* out_lo, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (rows < 1 || C < 1 || C > kRsThreads * kRsMaxPer || in0 == nullptr || (gamma != nullptr) != (beta != nullptr) ||
      (out_lo != nullptr && out_hi == nullptr))
    return fail(UDT_ERR_SHAPE, "udt_rowsum_norm_split: rows=%d C=%d (<= %d)", rows, C, kRsThreads * kRsMaxPer);
  udt_host::launch_pdl(rowsum_norm_split_kernel, dim3(rows), dim3(kRsThreads), 0, reinterpret_cast<cudaStream_t>(stream), in0, in1,
                       in2, res, C, gamma, beta, eps, relu, out_f32, reinterpret_cast<__half*>(out_hi),
                       reinterpret_cast<__half*>(out_lo));
  return check_launch("udt_rowsum_norm_split");
}

extern "C" int udt_mha_small_f32(const float* qkv, void* o_hi, void* o_lo, int32_t B, int32_t L, int32_t heads, int32_t dh,
                                 int32_t ld, int32_t ldo, float scale, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (L < 1 || L > kMhaMaxL || dh < 1 || dh > kMhaMaxD || B < 1 || heads < 1 || ld < 3 * heads * dh || o_hi == nullptr ||
      o_lo == nullptr)
    return fail(UDT_ERR_SHAPE, "udt_mha_small_f32: L=%d (<=%d) dh=%d (<=%d) ld=%d", L, kMhaMaxL, dh, kMhaMaxD, ld);
  const size_t smem = (static_cast<size_t>(3) * L * dh + static_cast<size_t>(L) * L) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(mha_small_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr_set = true;
  }
  udt_host::launch_pdl(mha_small_f32_kernel, dim3(heads, B), dim3(128), smem, reinterpret_cast<cudaStream_t>(stream), qkv,
                       reinterpret_cast<__half*>(o_hi), reinterpret_cast<__half*>(o_lo), L, heads, dh, ld, ldo, scale);
  return check_launch("udt_mha_small_f32");
}

extern "C" int udt_mha_small(const void* qkv, void* o, int32_t B, int32_t L, int32_t heads, int32_t dh, int32_t ld,
                             int32_t ldo, float scale, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (L < 1 || L > kMhaMaxL || dh < 1 || dh > kMhaMaxD || B < 1 || heads < 1 || ld < 3 * heads * dh)
    return fail(UDT_ERR_SHAPE, "udt_mha_small: L=%d (<=%d) dh=%d (<=%d) ld=%d", L, kMhaMaxL, dh, kMhaMaxD, ld);
  dim3 grid(heads, B);
  udt_host::launch_pdl(mha_small_kernel, dim3(grid), dim3(kMhaThreads), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __half*>(qkv), reinterpret_cast<__half*>(o), L, heads, dh, ld, ldo, scale);
  return check_launch("udt_mha_small");
}

extern "C" int udt_mha_masked(const void* q, const void* k, const void* v, void* o, int32_t B, int32_t Lq, int32_t Lk,
                              int32_t heads, int32_t dh, int32_t ldq, int32_t ldk, int32_t ldv, int32_t ldo, float scale,
                              const float* mask, int32_t ldm, const uint8_t* key_padding_mask, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (B < 1 || Lq < 1 || Lk < 1 || Lk > kMmMaxLk || heads < 1 || dh < 1 || dh > kMmMaxDh || (mask != nullptr && ldm < Lk))
    return fail(UDT_ERR_SHAPE, "udt_mha_masked: Lq=%d Lk=%d (<=%d) dh=%d (<=%d)", Lq, Lk, kMmMaxLk, dh, kMmMaxDh);
  const size_t smem = (static_cast<size_t>(2) * Lk * (dh + 1) + (kMmThreads / 32) * (dh + Lk)) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(mha_masked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_set = true;
  }
  udt_host::launch_pdl(mha_masked_kernel, dim3(heads, B), dim3(kMmThreads), smem, reinterpret_cast<cudaStream_t>(stream),
                       reinterpret_cast<const __half*>(q), reinterpret_cast<const __half*>(k), reinterpret_cast<const __half*>(v),
                       reinterpret_cast<__half*>(o), Lq, Lk, dh, ldq, ldk, ldv, ldo, scale, mask, ldm, key_padding_mask);
  return check_launch("udt_mha_masked");
}

extern "C" int udt_softmax_rows(void* x, int32_t rows, int32_t cols, int32_t ld, float scale, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (cols % 8 || ld % 8 || cols > kSmThreads * kSmMaxVec * 8 || cols < 8)
    return fail(UDT_ERR_SHAPE, "udt_softmax_rows: cols=%d ld=%d unsupported", cols, ld);
  udt_host::launch_pdl(softmax_rows_kernel, dim3(rows), dim3(kSmThreads), 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<__half*>(x), cols,
                                                                                       ld, scale);
  return check_launch("udt_softmax_rows");
}

extern "C" int udt_xattn_fold(const void* kc, const void* vc, int32_t ldkv, const void* wq, int32_t ldwq, const void* wo,
                              int32_t ldwo, void* w1, void* w2, int32_t B, int32_t L, int32_t heads, int32_t Npad, float scale,
                              void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (B < 1 || L < 1 || L > 16 || heads < 1 || Npad < heads * L || Npad % 8) return fail(UDT_ERR_SHAPE, "udt_xattn_fold: bad shape");
  const long long total = 2LL * B * Npad * heads * 64;
  udt_host::launch_pdl(xattn_fold_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0,
                       reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __half*>(kc),
                       reinterpret_cast<const __half*>(vc), ldkv, reinterpret_cast<const __half*>(wq), ldwq,
                       reinterpret_cast<const __half*>(wo), ldwo, reinterpret_cast<__half*>(w1), reinterpret_cast<__half*>(w2), B, L,
                       heads, Npad, scale);
  return check_launch("udt_xattn_fold");
}

extern "C" int udt_softmax_groups(const void* in, void* out, int32_t rows, int32_t cols, int32_t ld, int32_t groups, int32_t L,
                                  float* probs, int32_t N, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (rows < 1 || groups < 1 || L < 1 || L > 16 || groups * L > cols || cols > ld || (probs != nullptr && (N < 1 || rows % N)))
    return fail(UDT_ERR_SHAPE, "udt_softmax_groups: bad shape");
  const long long total = static_cast<long long>(rows) * (groups + 1);
  udt_host::launch_pdl(softmax_groups_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0,
                       reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __half*>(in), reinterpret_cast<__half*>(out),
                       rows, cols, ld, groups, L, probs, N);
  return check_launch("udt_softmax_groups");
}

extern "C" int udt_cfg_pack(const float* x, const float* concat_uc, const float* concat_c, void* unet_in, int32_t B,
                            int32_t HW, const float* c_in_dev, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (B < 1 || HW < 1 || c_in_dev == nullptr) return fail(UDT_ERR_SHAPE, "udt_cfg_pack: B=%d HW=%d", B, HW);
  const int total = 2 * B * HW;
  udt_host::launch_pdl(cfg_pack_kernel, dim3((total + 255) / 256), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, concat_uc, concat_c, reinterpret_cast<__half*>(unet_in), B, HW, c_in_dev);
  return check_launch("udt_cfg_pack");
}

extern "C" int udt_cfg_euler_step(float* x, const float* eps2b, int32_t B, int32_t HW, float cfg_scale,
                                  const float* dsigma_dev, const float* cfg_scale_dev, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (B < 1 || HW < 1 || dsigma_dev == nullptr) return fail(UDT_ERR_SHAPE, "udt_cfg_euler_step: B=%d HW=%d", B, HW);
  const int total = B * HW;
  udt_host::launch_pdl(cfg_euler_kernel, dim3((total + 255) / 256), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), x, eps2b, B, HW, cfg_scale,
                                                                                          dsigma_dev, cfg_scale_dev);
  return check_launch("udt_cfg_euler_step");
}

extern "C" int udt_vae_sample_pack(const float* moments, int32_t ld_moments, const float* noise_c, const float* noise_uc,
                                   const float* mask, float* concat_c, float* concat_uc, int32_t B, int32_t h, int32_t w,
                                   float scale_factor, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (B < 1 || h < 1 || w < 1 || ld_moments < 8) return fail(UDT_ERR_SHAPE, "udt_vae_sample_pack: B=%d h=%d w=%d ld=%d", B, h, w, ld_moments);
  const int total = B * h * w;
  udt_host::launch_pdl(vae_sample_pack_kernel, dim3((total + 255) / 256), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      moments, ld_moments, noise_c, noise_uc, mask, concat_c, concat_uc, B, h, w, scale_factor);
  return check_launch("udt_vae_sample_pack");
}

extern "C" int udt_pointwise_affine(const float* x, const float* Wm, const float* bias, void* out, int32_t B, int32_t HW,
                                    int32_t Cin, int32_t Cout, int32_t Cpad, float in_scale, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (B < 1 || HW < 1 || Cin < 1 || Cin > 8 || Cout < 1 || Cout > Cpad)
    return fail(UDT_ERR_SHAPE, "udt_pointwise_affine: Cin=%d (1..8) Cout=%d Cpad=%d", Cin, Cout, Cpad);
  const int total = B * HW;
  udt_host::launch_pdl(pointwise_affine_kernel, dim3((total + 255) / 256), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, Wm, bias, reinterpret_cast<__half*>(out), B, HW, Cin, Cout, Cpad, in_scale);
  return check_launch("udt_pointwise_affine");
}

extern "C" int udt_upsample2x_nhwc(const void* x, void* y, int32_t NB, int32_t H, int32_t W, int32_t C, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (C % 8) return fail(UDT_ERR_SHAPE, "udt_upsample2x_nhwc: C=%d not a multiple of 8", C);
  const size_t total = static_cast<size_t>(NB) * 4 * H * W * (C / 8);
  udt_host::launch_pdl(upsample2x_kernel, dim3(grid_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), NB, H, W, C / 8);
  return check_launch("udt_upsample2x_nhwc");
}

extern "C" int udt_nchw_f32_to_nhwc_f16(const float* x, void* y, int32_t NB, int32_t C, int32_t HW, int32_t Cpad,
                                        void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  const size_t total = static_cast<size_t>(NB) * HW * Cpad;
  udt_host::launch_pdl(nchw_to_nhwc_kernel, dim3(grid_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, reinterpret_cast<__half*>(y), NB, C, HW, Cpad);
  return check_launch("udt_nchw_f32_to_nhwc_f16");
}

extern "C" int udt_nhwc_to_nchw_f32(const void* x, int32_t x_is_fp32, float* y, int32_t NB, int32_t C, int32_t HW,
                                    int32_t ld, float scale, float shift, int32_t clamp01, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  const size_t total = static_cast<size_t>(NB) * C * HW;
  udt_host::launch_pdl(nhwc_to_nchw_kernel, dim3(grid_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, x_is_fp32, y, NB, C, HW, ld, scale, shift, clamp01);
  return check_launch("udt_nhwc_to_nchw_f32");
}

extern "C" int udt_attn_local_score(const float* probs, const float* mask, const float* seg, const float* gk, float* score,
                                    int32_t B, int32_t Bm, int32_t heads, int32_t size, int32_t L, int32_t seg_l, int32_t H,
                                    int32_t W, int32_t ks, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (B < 1 || Bm < 1 || B % Bm || heads < 1 || size < 1 || L < 1 || seg_l < 1 || seg_l > L || H < 1 || W < 1 || ks < 1 ||
      ks > kLsMaxK || (ks & 1) == 0)
    return fail(UDT_ERR_SHAPE, "udt_attn_local_score: B=%d Bm=%d heads=%d size=%d L=%d seg_l=%d ks=%d", B, Bm, heads, size, L,
                seg_l, ks);
  const size_t smem = static_cast<size_t>(2) * size * size * sizeof(float);
  if (smem > 200 * 1024) return fail(UDT_ERR_SHAPE, "udt_attn_local_score: map %dx%d does not fit shared memory", size, size);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_local_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail(UDT_ERR_LAUNCH, "cudaFuncSetAttribute(attn score smem): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  udt_host::launch_pdl(attn_local_score_kernel, dim3(B), dim3(kLsThreads), smem, reinterpret_cast<cudaStream_t>(stream), probs,
                       mask, seg, gk, score, Bm, heads, size, L, seg_l, H, W, ks);
  return check_launch("udt_attn_local_score");
}

extern "C" int udt_request_pack_u8(const uint8_t* image_hwc, const uint8_t* mask_hwc, float* image, float* mask, float* masked,
                                   int32_t B, int32_t Bs, int32_t H, int32_t W, int32_t MC, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (B < 1 || Bs < 1 || B % Bs || H < 1 || W < 1 || MC < 1 || MC > 4)
    return fail(UDT_ERR_SHAPE, "udt_request_pack_u8: B=%d Bs=%d H=%d W=%d MC=%d", B, Bs, H, W, MC);
  const size_t total = static_cast<size_t>(B) * H * W;
  udt_host::launch_pdl(request_pack_u8_kernel, dim3(grid_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                       image_hwc, mask_hwc, image, mask, masked, B, Bs, H * W, MC);
  return check_launch("udt_request_pack_u8");
}

extern "C" int udt_images_to_u8(const float* x, uint8_t* y, int32_t NB, int32_t C, int32_t HW, void* stream) {
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (NB < 1 || C < 1 || HW < 1) return fail(UDT_ERR_SHAPE, "udt_images_to_u8: NB=%d C=%d HW=%d", NB, C, HW);
  const size_t total = static_cast<size_t>(NB) * C * HW;
  udt_host::launch_pdl(images_to_u8_kernel, dim3(grid_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), x, y,
                       NB, C, HW);
  return check_launch("udt_images_to_u8");
}
