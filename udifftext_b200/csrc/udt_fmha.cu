// udt_fmha.cu — K4: softmax(Q K^T * scale) V for head dim 64 on tcgen05 tensor cores (sm_100a).
//
// One CTA owns two 128-row query tiles of one (batch, head) and streams the keys/values in 128-row tiles:
//   warp 9      TMA producer : Q tiles once, then a 2-stage ring of K / V tiles (128B-swizzled boxes)
//   warp 8      MMA issuer   : S_t = Q_t K_j^T  (M128 N128 K64, K-major operands)      -> TMEM S_t
//                              PV_t = P_t V_j   (M128 N64 K128, V as MN-major operand) -> TMEM O_t
//   warps 0-3 / 4-7          : softmax warpgroup of tile 0 / tile 1; thread = query row.  Online softmax in
//                              fp32 (exp2 domain), P_t written as fp16 into 128B-swizzled smem for the PV MMA,
//                              running output kept in registers and rescaled per key tile.
// The two query tiles ping-pong: while warpgroup 0 runs softmax on S_0 the tensor core works on tile 1.
// Replaces xformers.ops.memory_efficient_attention at reference sgm/modules/attention.py:246-248.
#include "udt_common.cuh"
#include "udt_host.h"

namespace {

using namespace udt;

constexpr int kTile = 128;
constexpr int kD = 64;
constexpr int kTileBytes = kTile * kD * 2;  // 16 KB: Q tile, K tile, V tile
constexpr int kPBytes = kTile * kTile * 2;  // 32 KB per query tile
constexpr int kThreads = 320;
constexpr int kTmemCols = 512;
constexpr int kColS = 0;    // S_t at columns [t*128, t*128+128)
constexpr int kColO = 256;  // O_t at columns [256 + t*64, ...+64)

struct FmhaParams {
  CUtensorMap mapQ, mapK, mapV;
  __half* o;
  int32_t Nq, Nkv, heads, ldo;
  float scale_log2;
};

// smem layout (offsets from the 1024-aligned base)
constexpr int kOffCtrl = 0;
constexpr int kOffQ = 1024;
constexpr int kOffK = kOffQ + 2 * kTileBytes;
constexpr int kOffV = kOffK + 2 * kTileBytes;
constexpr int kOffP = kOffV + 2 * kTileBytes;
constexpr int kSmemBytes = kOffP + 2 * kPBytes + 1024;

__global__ void __launch_bounds__(kThreads, 1) udt_fmha_kernel(const __grid_constant__ FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_addr - raw_addr);

  uint64_t* q_full = reinterpret_cast<uint64_t*>(base + kOffCtrl);
  uint64_t* kv_full = q_full + 1;   // [2]
  uint64_t* kv_empty = kv_full + 2; // [2]
  uint64_t* s_full = kv_empty + 2;  // [2]
  uint64_t* p_full = s_full + 2;    // [2]
  uint64_t* o_full = p_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.z;
  const int h = blockIdx.y;
  const int q0 = blockIdx.x * 2 * kTile;                 // first query row (within the batch) of this CTA
  const int ntiles = (p.Nq - q0 > kTile) ? 2 : 1;        // second query tile present?
  const int nkv = (p.Nkv + kTile - 1) / kTile;

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&p.mapQ);
    tma_prefetch_desc(&p.mapK);
    tma_prefetch_desc(&p.mapV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int col = h * kD;
      mbar_expect_tx(q_full, static_cast<uint32_t>(ntiles * kTileBytes));
      for (int t = 0; t < ntiles; ++t)
        tma_load_2d(&p.mapQ, q_full, base + kOffQ + t * kTileBytes, col, b * p.Nq + q0 + t * kTile);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1u);
        mbar_expect_tx(&kv_full[s], 2u * kTileBytes);
        tma_load_2d(&p.mapK, &kv_full[s], base + kOffK + s * kTileBytes, col, b * p.Nkv + j * kTile);
        tma_load_2d(&p.mapV, &kv_full[s], base + kOffV + s * kTileBytes, col, b * p.Nkv + j * kTile);
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(128, 128, false, false);
      const uint32_t idesc_o = umma_idesc_f16(128, 64, false, true);  // B = V is MN-major
      mbar_wait(q_full, 0);
      tc_fence_after();
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = base_addr + kOffK + s * kTileBytes;
        const uint32_t v_addr = base_addr + kOffV + s * kTileBytes;
        for (int t = 0; t < ntiles; ++t) {
          // S_t is free: p_full[t] of iteration j-1 was observed below before we get here.
          const uint64_t dq = umma_desc_kmajor_sw128(base_addr + kOffQ + t * kTileBytes);
          const uint64_t dk = umma_desc_kmajor_sw128(k_addr);
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk)
            umma_f16_ss(tmem_base + kColS + t * 128, dq + static_cast<uint64_t>(kk * 2), dk + static_cast<uint64_t>(kk * 2),
                        idesc_s, kk != 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
        }
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(&p_full[t], j & 1);
          tc_fence_after();
          const uint32_t p_addr = base_addr + kOffP + t * kPBytes;
#pragma unroll
          for (int kk = 0; kk < kTile / 16; ++kk) {
            const uint64_t dp = umma_desc_kmajor_sw128(p_addr + (kk >> 2) * kTileBytes + (kk & 3) * 32);
            const uint64_t dv = umma_desc_mnmajor_sw128(v_addr + kk * 16 * 128, 8192);
            umma_f16_ss(tmem_base + kColO + t * 64, dp, dv, idesc_o, kk != 0 ? 1u : 0u);
          }
          umma_commit(&o_full[t]);
        }
        umma_commit(&kv_empty[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int t = warp >> 2;  // query tile handled by this warpgroup
    if (t < ntiles) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t s_addr = tmem_base + lane_base + kColS + t * 128;
      const uint32_t o_addr = tmem_base + lane_base + kColO + t * 64;
      uint8_t* sP = base + kOffP + t * kPBytes;
      float m = -INFINITY, l = 0.0f;
      float acc[kD];
#pragma unroll
      for (int d = 0; d < kD; ++d) acc[d] = 0.0f;

      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        const int key_lim = p.Nkv - j * kTile;  // keys >= key_lim of this tile are padding
        float mx = m;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(s_addr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float sv = (c * 32 + i < key_lim) ? __uint_as_float(v[i]) * p.scale_log2 : -INFINITY;
            mx = fmaxf(mx, sv);
          }
        }
        const float alpha = ex2_approx(m - mx);  // m = -inf on the first tile -> 0
        if (j > 0) {
          mbar_wait(&o_full[t], (j - 1) & 1);  // PV_{j-1} done: O_t readable, sP reusable
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(o_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[c * 32 + i] += __uint_as_float(v[i]);
          }
        }
#pragma unroll
        for (int d = 0; d < kD; ++d) acc[d] *= alpha;
        l *= alpha;
        float rowsum = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(s_addr + c * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (c * 32 + i < key_lim) ? ex2_approx(__uint_as_float(v[i]) * p.scale_log2 - mx) : 0.0f;
            const float p1 = (c * 32 + i + 1 < key_lim) ? ex2_approx(__uint_as_float(v[i + 1]) * p.scale_log2 - mx) : 0.0f;
            rowsum += p0 + p1;
            pk[i >> 1] = pack_half2(p0, p1);
          }
          // keys [c*32, c*32+32): K-chunk (c>>1) of 64 keys, 16-byte groups ((c&1)*4 .. +3), 128B swizzle
          uint8_t* prow = sP + (c >> 1) * kTileBytes + row * 128;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int c16 = (c & 1) * 4 + g;
            uint4 val = make_uint4(pk[g * 4 + 0], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
            *reinterpret_cast<uint4*>(prow + ((c16 ^ (row & 7)) << 4)) = val;
          }
        }
        l += rowsum;
        m = mx;
        fence_proxy_async_smem();  // P visible to the tensor core (async proxy)
        tc_fence_before();         // our TMEM reads of S_t / O_t are ordered before the arrive
        mbar_arrive(&p_full[t]);
      }
      mbar_wait(&o_full[t], (nkv - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(o_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c * 32 + i] += __uint_as_float(v[i]);
      }
      const int qrow = q0 + t * kTile + row;
      if (qrow < p.Nq) {
        const float inv = 1.0f / l;
        uint4* o4 = reinterpret_cast<uint4*>(p.o + (static_cast<size_t>(b) * p.Nq + qrow) * p.ldo + h * kD);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          uint4 ov;
          ov.x = pack_half2(acc[v * 8 + 0] * inv, acc[v * 8 + 1] * inv);
          ov.y = pack_half2(acc[v * 8 + 2] * inv, acc[v * 8 + 3] * inv);
          ov.z = pack_half2(acc[v * 8 + 4] * inv, acc[v * 8 + 5] * inv);
          ov.w = pack_half2(acc[v * 8 + 6] * inv, acc[v * 8 + 7] * inv);
          o4[v] = ov;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace

extern "C" int udt_fmha_fwd(const void* q, const void* k, const void* v, void* o, int32_t B, int32_t Nq, int32_t Nkv,
                            int32_t heads, int32_t ldq, int32_t ldk, int32_t ldv, int32_t ldo, float scale,
                            void* stream) {
  using namespace udt_host;
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (B < 1 || Nq < 1 || Nkv < 1 || heads < 1) return fail(UDT_ERR_SHAPE, "udt_fmha_fwd: bad shape");
  if (ldo % 8 || (reinterpret_cast<uintptr_t>(o) & 15)) return fail(UDT_ERR_ALIGN, "udt_fmha_fwd: o / ldo alignment");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(udt_fmha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return fail(UDT_ERR_LAUNCH, "cudaFuncSetAttribute(fmha smem): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  FmhaParams p;
  rc = make_tmap_2d(&p.mapQ, q, static_cast<uint64_t>(heads) * kD, static_cast<uint64_t>(B) * Nq, ldq, kD, kTile);
  if (rc != UDT_OK) return rc;
  rc = make_tmap_2d(&p.mapK, k, static_cast<uint64_t>(heads) * kD, static_cast<uint64_t>(B) * Nkv, ldk, kD, kTile);
  if (rc != UDT_OK) return rc;
  rc = make_tmap_2d(&p.mapV, v, static_cast<uint64_t>(heads) * kD, static_cast<uint64_t>(B) * Nkv, ldv, kD, kTile);
  if (rc != UDT_OK) return rc;
  p.o = reinterpret_cast<__half*>(o);
  p.Nq = Nq;
  p.Nkv = Nkv;
  p.heads = heads;
  p.ldo = ldo;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((Nq + 2 * kTile - 1) / (2 * kTile), heads, B);
  udt_fmha_kernel<<<grid, kThreads, kSmemBytes, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("udt_fmha_fwd");
}
