// udt_fmha.cu — K4: softmax(Q K^T * scale) V for head dim 64 on tcgen05 tensor cores (sm_100a).
//
// One CTA owns two 128-row query tiles of one (batch, head) and streams the keys/values in 128-row tiles:
//   warp 9      TMA producer : Q tiles once, then a 3-stage ring of K / V tiles (128B-swizzled boxes)
//   warp 8      MMA issuer   : S_t = Q_t K_j^T   (M128 N128 K64, K-major operands)      -> TMEM S_t
//                              O_t += P_t V_j    (M128 N64 K128, V as MN-major operand) -> TMEM O_t (accumulating)
//                              issue order per key tile j and query tile t:  S_t(j+1) before P_t V_j, so the next
//                              scores are ready while the softmax warps still work on the other query tile.
//   warps 0-3 / 4-7          : softmax warpgroup of tile 0 / tile 1; thread = query row.  Online softmax in fp32
//                              (exp2 domain) with LAZY rescaling: the running output stays in TMEM and is only
//                              rescaled (tcgen05.ld / st round trip) when the row maximum grows by more than 2^8
//                              over the reference maximum — the probabilities are then bounded by 256, safe in
//                              fp16, and the final division by the row sum uses the same reference.
//                              P_t is written as fp16 into 128B-swizzled shared memory for the PV MMA.
// Replaces xformers.ops.memory_efficient_attention at reference sgm/modules/attention.py:246-248.
#include "udt_common.cuh"
#include "udt_host.h"

namespace {

using namespace udt;

constexpr int kTile = 128;
constexpr int kD = 64;
constexpr int kTileBytes = kTile * kD * 2;  // 16 KB: Q tile, K tile, V tile
constexpr int kPBytes = kTile * kTile * 2;  // 32 KB per query tile
constexpr int kKvStages = 3;
constexpr int kThreads = 320;
constexpr int kTmemCols = 512;
constexpr int kColS = 0;    // S_t at columns [t*128, t*128+128)
constexpr int kColO = 256;  // O_t at columns [256 + t*64, ...+64)
constexpr float kLazyThreshold = 8.0f;  // log2 units

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
constexpr int kOffV = kOffK + kKvStages * kTileBytes;
constexpr int kOffP = kOffV + kKvStages * kTileBytes;
constexpr int kSmemBytes = kOffP + 2 * kPBytes + 1024;

__global__ void __launch_bounds__(kThreads, 1) udt_fmha_kernel(const __grid_constant__ FmhaParams p) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  griddep_wait();     // PDL: wait for the producers of our inputs
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_addr - raw_addr);

  uint64_t* q_full = reinterpret_cast<uint64_t*>(base + kOffCtrl);
  uint64_t* kv_full = q_full + 1;            // [kKvStages]
  uint64_t* kv_empty = kv_full + kKvStages;  // [kKvStages]
  uint64_t* s_full = kv_empty + kKvStages;   // [2]
  uint64_t* p_full = s_full + 2;             // [2]
  uint64_t* o_full = p_full + 2;             // [2]
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
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
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
        const int s = j % kKvStages;
        mbar_wait(&kv_empty[s], ((j / kKvStages) & 1) ^ 1u);
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
      auto issue_s = [&](int t, int stage) {
        const uint64_t dq = umma_desc_kmajor_sw128(base_addr + kOffQ + t * kTileBytes);
        const uint64_t dk = umma_desc_kmajor_sw128(base_addr + kOffK + stage * kTileBytes);
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk)
          umma_f16_ss(tmem_base + kColS + t * 128, dq + static_cast<uint64_t>(kk * 2), dk + static_cast<uint64_t>(kk * 2),
                      idesc_s, kk != 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < ntiles; ++t) issue_s(t, 0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % kKvStages;
        const uint32_t v_addr = base_addr + kOffV + s * kTileBytes;
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(&p_full[t], j & 1);   // softmax t is done with S_t(j); P_t(j) is in smem; O_t is consistent
          tc_fence_after();
          if (j + 1 < nkv) {
            const int sn = (j + 1) % kKvStages;
            if (t == 0) {
              mbar_wait(&kv_full[sn], ((j + 1) / kKvStages) & 1);
              tc_fence_after();
            }
            issue_s(t, sn);               // next scores first: they are what the softmax warps wait for
          }
          const uint32_t p_addr = base_addr + kOffP + t * kPBytes;
#pragma unroll
          for (int kk = 0; kk < kTile / 16; ++kk) {
            const uint64_t dp = umma_desc_kmajor_sw128(p_addr + (kk >> 2) * kTileBytes + (kk & 3) * 32);
            const uint64_t dv = umma_desc_mnmajor_sw128(v_addr + kk * 16 * 128, 8192);
            umma_f16_ss(tmem_base + kColO + t * 64, dp, dv, idesc_o, (j | kk) != 0 ? 1u : 0u);
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
      const float sl2 = p.scale_log2;
      float m_ref = -INFINITY, l = 0.0f;

      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        const int key_lim = p.Nkv - j * kTile;  // keys >= key_lim of this tile are padding (only on the last tile)
        const bool partial = key_lim < kTile;
        // ---- pass 1: row maximum of the raw scores
        float mx = -INFINITY;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t v0[32], v1[32];
          tmem_ld32(s_addr + hf * 64, v0);
          tmem_ld32(s_addr + hf * 64 + 32, v1);
          tmem_ld_wait();
          if (!partial) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (hf * 64 + i < key_lim) mx = fmaxf(mx, __uint_as_float(v0[i]));
              if (hf * 64 + 32 + i < key_lim) mx = fmaxf(mx, __uint_as_float(v1[i]));
            }
          }
        }
        const float m_tile = mx * sl2;
        if (j > 0) {
          mbar_wait(&o_full[t], (j - 1) & 1);  // P_t V_{j-1} done: P buffer reusable, O_t stable
          tc_fence_after();
        }
        // ---- lazy rescale: only when this row's maximum outgrew the reference by more than 2^8
        const bool need = m_tile > m_ref + kLazyThreshold;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? m_tile : m_ref;
          const float alpha = need ? ex2_approx(m_ref - m_new) : 1.0f;  // m_ref = -inf on the first tile -> 0
          l *= alpha;
          if (j > 0) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t v[32];
              tmem_ld32(o_addr + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st32(o_addr + c * 32, v);
            }
            tmem_st_wait();
          }
          m_ref = m_new;
        }
        // ---- pass 2: probabilities (bounded by 2^8), row sum, P -> swizzled smem
        float rowsum = 0.0f;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t v0[32], v1[32];
          tmem_ld32(s_addr + hf * 64, v0);
          tmem_ld32(s_addr + hf * 64 + 32, v1);
          tmem_ld_wait();
          uint8_t* prow = sP + hf * kTileBytes + row * 128;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float s0 = __uint_as_float(c == 0 ? v0[i] : v1[i]);
              const float s1 = __uint_as_float(c == 0 ? v0[i + 1] : v1[i + 1]);
              float p0 = ex2_approx(fmaf(s0, sl2, -m_ref));
              float p1 = ex2_approx(fmaf(s1, sl2, -m_ref));
              if (partial) {
                const int k0 = hf * 64 + c * 32 + i;
                if (k0 >= key_lim) p0 = 0.0f;
                if (k0 + 1 >= key_lim) p1 = 0.0f;
              }
              rowsum += p0 + p1;
              pk[i >> 1] = pack_half2(p0, p1);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int c16 = c * 4 + g;  // 16-byte group within the 128-byte row of this 64-key half
              *reinterpret_cast<uint4*>(prow + ((c16 ^ (row & 7)) << 4)) =
                  make_uint4(pk[g * 4 + 0], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
            }
          }
        }
        l += rowsum;
        fence_proxy_async_smem();  // P visible to the tensor core (async proxy)
        tc_fence_before();         // TMEM reads of S_t / writes of O_t ordered before the arrive
        mbar_arrive(&p_full[t]);
      }
      mbar_wait(&o_full[t], (nkv - 1) & 1);
      tc_fence_after();
      const int qrow = q0 + t * kTile + row;
      const float inv = 1.0f / l;
      uint4* o4 = reinterpret_cast<uint4*>(p.o + (static_cast<size_t>(b) * p.Nq + min(qrow, p.Nq - 1)) * p.ldo + h * kD);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(o_addr + c * 32, v);
        tmem_ld_wait();
        if (qrow < p.Nq) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 ov;
            ov.x = pack_half2(__uint_as_float(v[g * 8 + 0]) * inv, __uint_as_float(v[g * 8 + 1]) * inv);
            ov.y = pack_half2(__uint_as_float(v[g * 8 + 2]) * inv, __uint_as_float(v[g * 8 + 3]) * inv);
            ov.z = pack_half2(__uint_as_float(v[g * 8 + 4]) * inv, __uint_as_float(v[g * 8 + 5]) * inv);
            ov.w = pack_half2(__uint_as_float(v[g * 8 + 6]) * inv, __uint_as_float(v[g * 8 + 7]) * inv);
            o4[c * 4 + g] = ov;
          }
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
  udt_host::launch_pdl(udt_fmha_kernel, dim3(grid), dim3(kThreads), kSmemBytes, reinterpret_cast<cudaStream_t>(stream), p);
  return check_launch("udt_fmha_fwd");
}
