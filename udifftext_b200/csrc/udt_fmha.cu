// udt_fmha.cu — K4: softmax(Q K^T * scale) V for head dim 64 on tcgen05 tensor cores (sm_100a).
//
// One CTA owns two 128-row query tiles of one (batch, head) and streams the keys/values in 128-row tiles:
//   warp 8      PV issuer    : O_t += P_t V_j    (M128 N64 K128, P from TMEM, V MN-major) -> TMEM O_t (accumulating);
//                              commits per score buffer (pv_done) and per query tile (o_full)
//   warp 9      score issuer : S = Q_t K_j^T     (M128 N128 K64, K-major operands)       -> one of THREE S buffers in TMEM,
//                + TMA ring    issued once the PV that last used the buffer has COMPLETED; the same event frees the K / V stage
//                              of a finished key tile, so this warp also loads Q once and keeps the 4-stage K / V ring full
//                              (128B-swizzled boxes).  Two issuing warps because ONE thread issuing the 12 MMAs + 3 commits of
//                              a computation was the kernel's pace-setter (see the comment in the kernel and DESIGN.md 4.1).
//                              The score tiles of both query tiles rotate through three TMEM buffers, so the scores a
//                              softmax warpgroup needs next are computed while it still works on its current tile, and the
//                              rotation keeps the two warpgroups in anti-phase on the exp pipe (a private score buffer per
//                              warpgroup ran them in lockstep: 270 vs 228 us).
//   warps 0-3 / 4-7          : softmax warpgroup of tile 0 / tile 1; thread = query row.  Online softmax in fp32
//                              (exp2 domain) with LAZY rescaling: the running output stays in TMEM and is only
//                              rescaled (tcgen05.ld / st round trip) when the row maximum grows by more than 2^8
//                              over the reference maximum — the probabilities are then bounded by 256, safe in
//                              fp16, and the final division by the row sum uses the same reference.  After the first
//                              key tile the scores are read from TMEM ONCE: exponentials are taken against the current
//                              reference while the row sum is tracked, and only a (rare) violation of the 2^8
//                              bound replays the tile.  The tcgen05.ld of the next 32 columns flies during the math.
//                              P_t goes back into TENSOR MEMORY as fp16 pairs over the first 64 columns of its own
//                              (fully consumed) score buffer and feeds the PV MMA as a TMEM A operand (TS form):
//                              tcgen05.st is 4x faster than tcgen05.ld, no shared-memory round trip, and the shared
//                              memory a P buffer would need buys the 4-stage K / V ring.
// Floors per 128x128 score tile (measured): tcgen05.ld 64 B/clk/SM -> 1024 clk; MUFU.EX2 16/clk/SM -> 1024 clk
// (`ex2.approx.f16x2` is NOT a way around it: ptxas emits two MUFU.EX2.F16 + a PRMT for it on sm_100a); the MMAs need 512 clk.
// Replaces xformers.ops.memory_efficient_attention at reference sgm/modules/attention.py:246-248.
#include "udt_common.cuh"
#include "udt_host.h"

namespace {

using namespace udt;

constexpr int kTile = 128;
constexpr int kD = 64;
constexpr int kTileBytes = kTile * kD * 2;  // 16 KB: Q tile, K tile, V tile
constexpr int kKvStages = 4;   // K / V ring depth
constexpr int kSBufs = 3;
constexpr int kThreads = 320;
constexpr int kTmemCols = 512;
constexpr int kColS = 0;    // S buffer b at columns [b*128, b*128+128)
constexpr int kColO = 384;  // O_t at columns [384 + t*64, ...+64)
constexpr float kLazyThreshold = 8.0f;  // log2 units

#ifdef UDT_IGEMM_TRACE
#define UDT_FDBG(bit) (p.debug & (bit))    // tuning builds (UDT_TRACE=1): UDT_FMHA_DEBUG experiment switches
#else
#define UDT_FDBG(bit) (false)
#endif
#ifdef UDT_FMHA_STAMPS
// tuning builds with UDT_STAMPS=1, UDT_FMHA_DEBUG & 64: lane 0 of every softmax warp of CTA 0 stamps clock() at fixed points of
// key tiles 8..11 and prints them when the CTA ends (scripts/fmha_timeline.py).  The stamps cost registers (spills) — timings of
// such a build are not comparable; stamps in the MMA-issuing warps would lengthen the issue chain they measure.
#define UDT_FSTAMP_DECL uint32_t stamps[4 * 8]; const bool stamp_cta = (p.debug & 64) && blockIdx.x == 0 && lane == 0
#define UDT_FSTAMP(j, slot) do { if (stamp_cta && (j) >= 8 && (j) < 12) stamps[((j) - 8) * 8 + (slot)] = static_cast<uint32_t>(clock()); } while (0)
#define UDT_FSTAMP_DUMP(nj) do { if (stamp_cta && (nj) >= 12) for (int jj = 0; jj < 4; ++jj) \
  printf("FSTAMP w%d j%d %u %u %u %u %u %u %u %u\n", warp, jj + 8, stamps[jj * 8], stamps[jj * 8 + 1], stamps[jj * 8 + 2], \
         stamps[jj * 8 + 3], stamps[jj * 8 + 4], stamps[jj * 8 + 5], stamps[jj * 8 + 6], stamps[jj * 8 + 7]); } while (0)
#else
#define UDT_FSTAMP_DECL
#define UDT_FSTAMP(j, slot)
#define UDT_FSTAMP_DUMP(nj)
#endif

struct FmhaParams {
  CUtensorMap mapQ, mapK, mapV;
  __half* o;
  int32_t Nq, Nkv, heads, ldo;
  float scale_log2;
  int32_t debug;   // UDT_FMHA_DEBUG experiment switches (tuning only; 0 in production)
  int32_t pairs_per_bh;   // CTAs [0, pairs_full) own a PAIR of query tiles (256 rows) of one (batch, head), numbered
  int32_t pairs_full;     // (b * heads + h) * pairs_per_bh + pair; the CTAs after them own ONE tile of the remaining pairs
  int32_t num_items;      // pairs_full + single-tile items (= the grid of the one-item-per-CTA kernel)
};

// 2^x on the FMA / ALU pipes (the MUFU pipe, 16 ex2/clk/SM, is the busiest pipe of this kernel: 66 % under ncu): round-to-
// nearest split x = j + f by the 1.5 * 2^23 trick, degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error
// 7.5e-5, well below the fp16 rounding of P, 4.9e-4), j added into the exponent field.  x is clamped to [-126, 126]: below,
// the result is ~2^-126 ~ 0; above, it is huge and trips the row-sum bound exactly like the MUFU result (+inf) would.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fminf(fmaxf(x, -126.0f), 126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  const float p = fmaf(fmaf(fmaf(0.055171654f, f, 0.24261113f), f, 0.69326097f), f, 0.99992806f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

// Work of this CTA.  The grid is linear: whole waves of tile pairs first, then — when the last, partial wave would leave
// more than half of the SMs idle — its pairs as twice as many single-tile CTAs (a single-tile CTA has the exp pipe and
// the TMEM read port to itself and finishes in ~0.55 of a pair's time, so the tail wave shrinks accordingly).
__device__ __forceinline__ bool fmha_item(const FmhaParams& p, int L, int& b, int& h, int& q0, int& ntiles) {
  int pair = L, half = 0;
  const bool single = L >= p.pairs_full;
  if (single) {
    const int s = L - p.pairs_full;
    pair = p.pairs_full + (s >> 1);
    half = s & 1;
  }
  const int bh = pair / p.pairs_per_bh;
  b = bh / p.heads;
  h = bh - b * p.heads;
  q0 = (pair - bh * p.pairs_per_bh) * 2 * kTile + half * kTile;   // first query row (within the batch) of this CTA
  ntiles = (!single && p.Nq - q0 > kTile) ? 2 : 1;                // second query tile present?
  return q0 < p.Nq;
}

__device__ __forceinline__ bool fmha_work(const FmhaParams& p, int& b, int& h, int& q0, int& ntiles) {
  return fmha_item(p, static_cast<int>(blockIdx.x), b, h, q0, ntiles);
}

// smem layout (offsets from the 1024-aligned base): control words, 2 Q tiles, the K ring, the V ring (P lives in TMEM)
constexpr int kOffCtrl = 0;
constexpr int kOffQ = 1024;
constexpr int kTsKvStages = kKvStages;
constexpr int kTsOffK = kOffQ + 2 * kTileBytes;
constexpr int kTsOffV = kTsOffK + kTsKvStages * kTileBytes;
constexpr int kTsSmemBytes = kTsOffV + kTsKvStages * kTileBytes + 1024;

// cursor over the score computations c = j * ntiles + t (key tile j, query tile t) without integer divisions
struct CompCursor {
  int c, j, t, stage, kvphase, b, u;   // stage = j % kKvStages, kvphase = (j / kKvStages) & 1, b = c % kSBufs, u = c / kSBufs
  __device__ __forceinline__ void init() { c = j = t = stage = kvphase = b = u = 0; }
  __device__ __forceinline__ void advance(int ntiles, int kvstages = kKvStages) {
    ++c;
    if (++b == kSBufs) { b = 0; ++u; }
    if (++t == ntiles) {
      t = 0;
      ++j;
      if (++stage == kvstages) { stage = 0; kvphase ^= 1; }
    }
  }
};

// kPoly > 0: every kPoly-th exponential of the single-pass path is evaluated by ex2_poly instead of MUFU.EX2
template <int kPoly>
__global__ void __launch_bounds__(kThreads, 1) udt_fmha_ts_kernel(const __grid_constant__ FmhaParams p) {
  griddep_launch();   // PDL: let the next kernel's prologue start
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_addr - raw_addr);

  uint64_t* q_full = reinterpret_cast<uint64_t*>(base + kOffCtrl);
  uint64_t* kv_full = q_full + 1;            // [kTsKvStages]
  uint64_t* s_full = kv_full + kTsKvStages;  // [kSBufs]  scores of a computation are in TMEM
  uint64_t* p_full = s_full + kSBufs;        // [2]       P_t(j) is in TMEM (over its score buffer)
  uint64_t* o_full = p_full + 2;             // [2]       P_t(j) V_j has been accumulated into O_t (softmax warpgroup t waits)
  uint64_t* pv_done = o_full + 2;            // [kSBufs]  same event per score buffer (the score issuer waits: buffer / K,V stage free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + kSBufs);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int b, h, q0, ntiles;
  if (!fmha_work(p, b, h, q0, ntiles)) return;           // odd tile count: the second half of the last pair is empty
  const int nkv = (p.Nkv + kTile - 1) / kTile;
  const int ncomp = nkv * ntiles;

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&p.mapQ);
    tma_prefetch_desc(&p.mapK);
    tma_prefetch_desc(&p.mapV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kTsKvStages; ++i) mbar_init(&kv_full[i], 1);
    for (int i = 0; i < kSBufs; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
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
  griddep_wait();     // PDL: q / k / v are produced by the previous kernel

  if (warp == 9) {
    // ------------------------------------------------------------------ score issuer + TMA producer
    // S(c) = Q_t K_j^T goes into score buffer c % 3 once P(c-3) V — the previous user of that buffer, issued by warp 8 — has
    // COMPLETED (pv_done; the two issuing threads are ordered through that barrier, tcgen05.commit -> wait ->
    // tcgen05.fence::after_thread_sync).  The same event frees the K / V stage of a finished key tile, so this warp also
    // refills the ring: no producer warp, no kv_empty barriers, no commit for them.
    // Why two issuing warps: one thread issuing all twelve MMAs + three commits of a computation needs ~900 clk per
    // computation (~40 clk per tcgen05.mma, ~100 per commit; scripts/fmha_timeline.py), ~2000 clk per key tile against
    // ~1750 clk of softmax work per warpgroup — the issue chain, not the exp pipe, set the pace; and while a tcgen05.mma waits
    // for room in the tensor-pipe queue it sits at the head of its sub-partition's MIO queue, where the MUFU / tcgen05.ld
    // instructions of that sub-partition's two softmax warps wait behind it (they needed ~2450 clk per key tile and moved
    // with the issuing warp when it was placed on another sub-partition).  Split, each thread has at most 2/3 of that work,
    // the chains overlap, and the MIO cost is shared by two sub-partitions.
    const bool issuer = elect_one();
    const uint32_t idesc_s = umma_idesc_f16(128, 128, false, false);
    const int col = h * kD;
    auto load_kv = [&](int j, int stage) {
      mbar_expect_tx(&kv_full[stage], 2u * kTileBytes);
      tma_load_2d(&p.mapK, &kv_full[stage], base + kTsOffK + stage * kTileBytes, col, b * p.Nkv + j * kTile);
      tma_load_2d(&p.mapV, &kv_full[stage], base + kTsOffV + stage * kTileBytes, col, b * p.Nkv + j * kTile);
    };
    if (issuer) {
      mbar_expect_tx(q_full, static_cast<uint32_t>(ntiles * kTileBytes));
      for (int t = 0; t < ntiles; ++t)
        tma_load_2d(&p.mapQ, q_full, base + kOffQ + t * kTileBytes, col, b * p.Nq + q0 + t * kTile);
      for (int j = 0; j < kTsKvStages && j < nkv; ++j) load_kv(j, j);
    }
    __syncwarp();
    mbar_wait(q_full, 0);
    CompCursor ks, kd;   // kd trails ks by kSBufs computations: the previous user of the score buffer ks is about to overwrite
    ks.init();
    kd.init();
    for (; ks.c < ncomp; ks.advance(ntiles, kTsKvStages)) {
      if (ks.t == 0) mbar_wait(&kv_full[ks.stage], static_cast<uint32_t>(ks.kvphase));   // K_j: long landed, checked off the critical path
      if (ks.c >= kSBufs) {
        mbar_wait(&pv_done[kd.b], static_cast<uint32_t>(kd.u & 1));
        if (kd.t == ntiles - 1 && kd.j + kTsKvStages < nkv && issuer)   // key tile kd.j is finished: its stage takes tile j + stages
          load_kv(kd.j + kTsKvStages, kd.stage);
        kd.advance(ntiles, kTsKvStages);
      }
      tc_fence_after();
      if (issuer) {
        const uint64_t dq = umma_desc_kmajor_sw128(base_addr + kOffQ + ks.t * kTileBytes);
        const uint64_t dk = umma_desc_kmajor_sw128(base_addr + kTsOffK + ks.stage * kTileBytes);
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk)
          umma_f16_ss(tmem_base + kColS + ks.b * 128, dq + static_cast<uint64_t>(kk * 2), dk + static_cast<uint64_t>(kk * 2),
                      idesc_s, kk != 0 ? 1u : 0u);
        umma_commit(&s_full[ks.b]);
      }
      __syncwarp();
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ PV issuer (converged warp, one elected lane issues)
    const bool issuer = elect_one();
    const uint32_t idesc_o = umma_idesc_f16(128, 64, false, true);  // B = V is MN-major
    CompCursor kp;
    kp.init();
    for (; kp.c < ncomp; kp.advance(ntiles, kTsKvStages)) {
      mbar_wait(&p_full[kp.t], static_cast<uint32_t>(kp.j & 1));   // P_t(j) is in TMEM over the first 64 columns of its score buffer
      if (kp.t == 0) mbar_wait(&kv_full[kp.stage], static_cast<uint32_t>(kp.kvphase));   // V_j (lands with K_j; cannot be refilled before this PV)
      tc_fence_after();
      if (issuer) {
        const uint32_t v_addr = base_addr + kTsOffV + kp.stage * kTileBytes;
        const uint32_t p_tmem = tmem_base + kColS + kp.b * 128;
#pragma unroll
        for (int kk = 0; kk < kTile / 16; ++kk) {
          const uint64_t dv = umma_desc_mnmajor_sw128(v_addr + kk * 16 * 128, 8192);
          umma_f16_ts(tmem_base + kColO + kp.t * 64, p_tmem + kk * 8, dv, idesc_o, (kp.j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&pv_done[kp.b]);   // first: the score issuer's wait is on the critical path, the warpgroup's is not
        umma_commit(&o_full[kp.t]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int t = warp >> 2;  // query tile handled by this warpgroup
    UDT_FSTAMP_DECL;
    if (t < ntiles) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t o_addr = tmem_base + lane_base + kColO + t * 64;
      const float sl2 = p.scale_log2;
      float m_ref = -INFINITY, l = 0.0f;
      int sb = t, su = 0;   // S buffer / use count of this warpgroup's next computation (c = j * ntiles + t)
#ifdef UDT_IGEMM_TRACE
      int cnt_replay = 0, cnt_rescale = 0;   // UDT_FMHA_DEBUG & 32: replays of the single pass / O rescales seen by this warp
#endif

      // rescale the running output (and row sum) when the reference maximum moves; O_t must be stable
      auto rescale = [&](bool need, float m_tile, int j) {
        const float m_new = need ? m_tile : m_ref;
        const float alpha = need ? ex2_approx(m_ref - m_new) : 1.0f;  // m_ref = -inf on the first tile -> 0
        l *= alpha;
        if (j > 0) {
          mbar_wait(&o_full[t], static_cast<uint32_t>((j - 1) & 1));
          tc_fence_after();
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
      };

      for (int j = 0; j < nkv; ++j) {
        UDT_FSTAMP(j, 0);
        mbar_wait(&s_full[sb], static_cast<uint32_t>(su & 1));
        tc_fence_after();
        UDT_FSTAMP(j, 1);
        const uint32_t s_addr = tmem_base + lane_base + kColS + sb * 128;
        const int key_lim = p.Nkv - j * kTile;  // keys >= key_lim of this tile are padding (only on the last tile)
        const bool partial = key_lim < kTile;
        float rowsum = 0.0f;
        bool replay = (j == 0) || partial;      // first / ragged tile: maximum first, then the exponentials
        // 32 probabilities (keys [32*ch, 32*ch+32) of this tile) -> fp16 -> this row's swizzled slots of the P buffer;
        // the stores interleave with the exponentials of the following chunk
        uint32_t pall[64];                      // the tile's probabilities, fp16 pairs (key 2i | key 2i+1)
        auto store_chunk = [&](const uint32_t (&pk)[16], int ch) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pall[ch * 16 + i] = pk[i];
        };
        if (!replay) {
          // ---- single pass: exponentials against the current reference, written to the P buffer right away.  No
          // maximum is tracked: every probability is bounded by 2^8 unless the tile's row sum exceeds 2^8, so a row sum
          // above that bound (rare: the reference would have to be stale by almost the whole lazy margin) sends the tile
          // through the two-pass path, which overwrites the optimistic P (nobody reads it before p_full).
          uint32_t va[32], vb[32];
          auto exp_chunk = [&](const uint32_t (&vv)[32], int ch) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float p0 = fmaf(__uint_as_float(vv[i]), sl2, -m_ref);
              float p1 = fmaf(__uint_as_float(vv[i + 1]), sl2, -m_ref);
              if (!UDT_FDBG(1)) {
                p0 = (kPoly > 0 && (i % kPoly) == kPoly - 1) ? ex2_poly(p0) : ex2_approx(p0);
                p1 = (kPoly > 0 && ((i + 1) % kPoly) == kPoly - 1) ? ex2_poly(p1) : ex2_approx(p1);
              }
              rowsum += p0 + p1;
              pk[i >> 1] = pack_half2(p0, p1);
            }
            if (!UDT_FDBG(2)) store_chunk(pk, ch);
          };
          tmem_ld32(s_addr, va);
          tmem_ld_wait_dep(va);
          UDT_FSTAMP(j, 2);
          tmem_ld32(s_addr + 32, vb);      // the next 32 columns fly during the math
          exp_chunk(va, 0);
          UDT_FSTAMP(j, 3);
          tmem_ld_wait_dep(vb);
          tmem_ld32(s_addr + 64, va);
          exp_chunk(vb, 1);
          tmem_ld_wait_dep(va);
          tmem_ld32(s_addr + 96, vb);
          exp_chunk(va, 2);
          tmem_ld_wait_dep(vb);
          UDT_FSTAMP(j, 4);
          exp_chunk(vb, 3);
          UDT_FSTAMP(j, 5);
          replay = __any_sync(0xffffffffu, !(rowsum <= 256.0f)) && !UDT_FDBG(3);   // also catches inf / nan
#ifdef UDT_IGEMM_TRACE
          cnt_replay += replay ? 1 : 0;
#endif
        }
        if (replay) {
          // ---- two passes: row maximum of the raw scores, reference update (+ O rescale), exponentials
          uint32_t v[32];
          float mx = -INFINITY;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            tmem_ld32(s_addr + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!partial || ch * 32 + i < key_lim) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
          const float m_tile = mx * sl2;
          const bool need = m_tile > m_ref + kLazyThreshold;
          if (__any_sync(0xffffffffu, need)) rescale(need, m_tile, j);
#ifdef UDT_IGEMM_TRACE
          cnt_rescale += (j > 0 && __any_sync(0xffffffffu, need)) ? 1 : 0;
#endif
          rowsum = 0.0f;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            tmem_ld32(s_addr + ch * 32, v);
            tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), sl2, -m_ref));
              float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sl2, -m_ref));
              if (partial) {
                const int k0 = ch * 32 + i;
                if (k0 >= key_lim) p0 = 0.0f;
                if (k0 + 1 >= key_lim) p1 = 0.0f;
              }
              rowsum += p0 + p1;
              pk[i >> 1] = pack_half2(p0, p1);
            }
            store_chunk(pk, ch);
          }
        }
        l += rowsum;
        // ---- P -> TMEM, over the first 64 columns of the (fully consumed) score buffer: the PV MMA reads it as its A operand
        {
          uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&pall[0]);
          uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&pall[32]);
          tmem_st32(s_addr, lo);
          tmem_st32(s_addr + 32, hi);
          tmem_st_wait();
        }
        UDT_FSTAMP(j, 6);
        // observe every o_full phase (the wait is almost always already satisfied: P_t V_{j-1} ran during this tile's
        // exponentials); a parity wait that skipped phases would be ambiguous in rescale() and at the end
        if (j > 0) mbar_wait(&o_full[t], static_cast<uint32_t>((j - 1) & 1));
        tc_fence_before();         // TMEM reads of S / writes of P, O_t ordered before the arrive
        mbar_arrive(&p_full[t]);
        UDT_FSTAMP(j, 7);
        sb += ntiles;
        if (sb >= kSBufs) { sb -= kSBufs; ++su; }
      }
      mbar_wait(&o_full[t], static_cast<uint32_t>((nkv - 1) & 1));
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
      UDT_FSTAMP_DUMP(nkv);
#ifdef UDT_IGEMM_TRACE
      if ((p.debug & 32) && lane == 0 && (blockIdx.x % 37) == 0)
        printf("FCNT nkv %d heads %d cta %d warp %d replays %d rescales %d\n", nkv, p.heads, static_cast<int>(blockIdx.x), warp, cnt_replay, cnt_rescale);
#endif
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}



#ifdef UDT_TUNING
// ------------------------------------------------------------------------------------------------------------------
// EXPERIMENT (tuning builds only, UDT_FMHA_PERSIST=1) — measured on B200: correct on the first run (all accuracy cases incl.
// ragged lengths, the 12 parity tests) and NOT faster: 4096 tokens 230.7 vs 227.7 us, 1024 tokens 45.4 vs 42.7, 256 tokens
// 15.4 vs 12.6, batch-32 shape (35 items per CTA) 1745 vs 1746 us.  The per-item start-up it was built to hide is therefore not
// launch / allocation / first-load latency between CTAs (that would have shown at 35 items per CTA); what remains per item is
// inherent — the two-pass first tile, the output epilogue, the refill of the score pipeline.
// PERSISTENT variant of the kernel above (same roles, same softmax code): one CTA per SM walks the work items
// blockIdx.x, blockIdx.x + gridDim.x, ... and every barrier phase, the score-buffer rotation and the K / V ring simply continue
// across items, so the next item's Q / K / V loads and first score MMAs run while the warpgroups still finish the current
// item, and the ~8 k clk a CTA of the kernel above spends on launch, TMEM allocation, the first loads and its drain are paid
// once per SM instead of once per item.  Items with the same number of query tiles form an EPOCH; at an epoch boundary (the
// single-tile items of the tail wave, ragged sequence lengths) the CTA drains and re-initialises its barriers.
// Extra state against the kernel above: a second Q buffer (Q of item e+1 is loaded when every score MMA of item e-1 has
// completed), o_free[t] (warpgroup t has read O_t of the previous item: the first P V of the next may overwrite it).
constexpr int kPkOffQ = 1024;                               // two Q buffers of two tiles
constexpr int kPkOffK = kPkOffQ + 4 * kTileBytes;
constexpr int kPkOffV = kPkOffK + kKvStages * kTileBytes;
constexpr int kPkSmemBytes = kPkOffV + kKvStages * kTileBytes + 1024;

// cursor over the computations of an epoch: item e, key tile j, query tile t; g = e * nkv + j (key tiles since the epoch began)
struct PkCursor {
  int c, e, j, t, g, stage, kvphase, b, u;
  __device__ __forceinline__ void init() { c = e = j = t = g = stage = kvphase = b = u = 0; }
  __device__ __forceinline__ void advance(int ntiles, int nkv) {
    ++c;
    if (++b == kSBufs) { b = 0; ++u; }
    if (++t == ntiles) {
      t = 0;
      ++g;
      if (++stage == kKvStages) { stage = 0; kvphase ^= 1; }
      if (++j == nkv) { j = 0; ++e; }
    }
  }
};

__global__ void __launch_bounds__(kThreads, 1) udt_fmha_pk_kernel(const __grid_constant__ FmhaParams p) {
  constexpr int kPoly = 0;
  griddep_launch();   // PDL: let the next kernel's prologue start
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_addr - raw_addr);

  uint64_t* q_full = reinterpret_cast<uint64_t*>(base + kOffCtrl);   // [2]
  uint64_t* kv_full = q_full + 2;            // [kKvStages]
  uint64_t* s_full = kv_full + kKvStages;    // [kSBufs]
  uint64_t* pv_done = s_full + kSBufs;       // [kSBufs]
  uint64_t* p_full = pv_done + kSBufs;       // [2]
  uint64_t* o_full = p_full + 2;             // [2]
  uint64_t* o_free = o_full + 2;             // [2]   warpgroup t has read O_t of the item it just finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkv = (p.Nkv + kTile - 1) / kTile;
  const int G = static_cast<int>(gridDim.x);

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&p.mapQ);
    tma_prefetch_desc(&p.mapK);
    tma_prefetch_desc(&p.mapV);
  }
  if (warp == 8) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();     // PDL: q / k / v are produced by the previous kernel

  bool first_epoch = true;
  int L = static_cast<int>(blockIdx.x);
  while (L < p.num_items) {
    int b0, h0, q00, ntiles;
    if (!fmha_item(p, L, b0, h0, q00, ntiles)) {   // odd tile count: the second half of a split pair does not exist
      L += G;
      continue;
    }
    int E = 1;                                     // items of this epoch: L, L + G, ... with the same number of query tiles
    for (;;) {
      const int Ln = L + E * G;
      int bb, hh, qq, nt;
      if (Ln >= p.num_items || !fmha_item(p, Ln, bb, hh, qq, nt) || nt != ntiles) break;
      ++E;
    }
    // ---- (re)initialise the barriers: everybody has left the previous epoch, none of its phases is pending
    if (!first_epoch) {
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
    if (threadIdx.x == 0) {
      uint64_t* bars = q_full;
      if (!first_epoch)
        for (int i = 0; i < 2 + kKvStages + 2 * kSBufs + 6; ++i) mbar_inval(&bars[i]);
      for (int i = 0; i < 2; ++i) mbar_init(&q_full[i], 1);
      for (int i = 0; i < kKvStages; ++i) mbar_init(&kv_full[i], 1);
      for (int i = 0; i < kSBufs; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&pv_done[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&p_full[i], 128);
        mbar_init(&o_full[i], 1);
        mbar_init(&o_free[i], 128);
      }
      fence_mbar_init();
    }
    __syncthreads();
    first_epoch = false;
    const int ncomp = E * nkv * ntiles;            // computations of the epoch

    if (warp == 9) {
      // ---------------------------------------------------------------- score issuer + TMA producer
      const bool issuer = elect_one();
      const uint32_t idesc_s = umma_idesc_f16(128, 128, false, false);
      const int T = E * nkv;                       // key tiles of the epoch
      int le = 0, lj = 0, ls = 0, lb = b0, lh = h0, loaded = 0;   // (elected lane) next key tile to load: item, tile, stage, its batch / head
      auto load_q = [&](int e) {
        int bb, hh, qq, nt;
        fmha_item(p, L + e * G, bb, hh, qq, nt);
        uint64_t* bar = &q_full[e & 1];
        mbar_expect_tx(bar, static_cast<uint32_t>(ntiles * kTileBytes));
        for (int t = 0; t < ntiles; ++t)
          tma_load_2d(&p.mapQ, bar, base + kPkOffQ + ((e & 1) * 2 + t) * kTileBytes, hh * kD, bb * p.Nq + qq + t * kTile);
      };
      auto load_next_kv = [&]() {
        mbar_expect_tx(&kv_full[ls], 2u * kTileBytes);
        tma_load_2d(&p.mapK, &kv_full[ls], base + kPkOffK + ls * kTileBytes, lh * kD, lb * p.Nkv + lj * kTile);
        tma_load_2d(&p.mapV, &kv_full[ls], base + kPkOffV + ls * kTileBytes, lh * kD, lb * p.Nkv + lj * kTile);
        ++loaded;
        if (++ls == kKvStages) ls = 0;
        if (++lj == nkv) {
          lj = 0;
          if (++le < E) {
            int qq, nt;
            fmha_item(p, L + le * G, lb, lh, qq, nt);
          }
        }
      };
      if (issuer) {
        load_q(0);
        if (E > 1) load_q(1);
        while (loaded < kKvStages && loaded < T) load_next_kv();
      }
      __syncwarp();
      PkCursor ks, kd;   // kd trails ks by kSBufs computations: the previous user of the score buffer ks is about to overwrite
      ks.init();
      kd.init();
      for (; ks.c < ncomp; ks.advance(ntiles, nkv)) {
        if (ks.t == 0) {
          mbar_wait(&kv_full[ks.stage], static_cast<uint32_t>(ks.kvphase));                       // K_j
          if (ks.j == 0) mbar_wait(&q_full[ks.e & 1], static_cast<uint32_t>((ks.e >> 1) & 1));    // Q of a new item
        }
        if (ks.c >= kSBufs) {
          // kd enters item e >= 1: every P V — hence every score MMA — of item e-1 has completed, its Q buffer takes item e+1
          if (kd.j == 0 && kd.t == 0 && kd.e >= 1 && kd.e + 1 < E && issuer) load_q(kd.e + 1);
          mbar_wait(&pv_done[kd.b], static_cast<uint32_t>(kd.u & 1));
          if (kd.t == ntiles - 1 && issuer && loaded < T) load_next_kv();   // a key tile is finished: its stage takes the next one
          kd.advance(ntiles, nkv);
        }
        tc_fence_after();
        if (issuer) {
          const uint64_t dq = umma_desc_kmajor_sw128(base_addr + kPkOffQ + ((ks.e & 1) * 2 + ks.t) * kTileBytes);
          const uint64_t dk = umma_desc_kmajor_sw128(base_addr + kPkOffK + ks.stage * kTileBytes);
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk)
            umma_f16_ss(tmem_base + kColS + ks.b * 128, dq + static_cast<uint64_t>(kk * 2), dk + static_cast<uint64_t>(kk * 2),
                        idesc_s, kk != 0 ? 1u : 0u);
          umma_commit(&s_full[ks.b]);
        }
        __syncwarp();
      }
    } else if (warp == 8) {
      // ---------------------------------------------------------------- PV issuer
      const bool issuer = elect_one();
      const uint32_t idesc_o = umma_idesc_f16(128, 64, false, true);  // B = V is MN-major
      PkCursor kp;
      kp.init();
      for (; kp.c < ncomp; kp.advance(ntiles, nkv)) {
        if (kp.j == 0 && kp.e > 0) mbar_wait(&o_free[kp.t], static_cast<uint32_t>((kp.e - 1) & 1));   // O_t of the previous item has been read out
        mbar_wait(&p_full[kp.t], static_cast<uint32_t>(kp.g & 1));
        if (kp.t == 0) mbar_wait(&kv_full[kp.stage], static_cast<uint32_t>(kp.kvphase));
        tc_fence_after();
        if (issuer) {
          const uint32_t v_addr = base_addr + kPkOffV + kp.stage * kTileBytes;
          const uint32_t p_tmem = tmem_base + kColS + kp.b * 128;
#pragma unroll
          for (int kk = 0; kk < kTile / 16; ++kk) {
            const uint64_t dv = umma_desc_mnmajor_sw128(v_addr + kk * 16 * 128, 8192);
            umma_f16_ts(tmem_base + kColO + kp.t * 64, p_tmem + kk * 8, dv, idesc_o, (kp.j | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&pv_done[kp.b]);
          umma_commit(&o_full[kp.t]);
        }
        __syncwarp();
      }
    } else {
      // ---------------------------------------------------------------- softmax warpgroups (code of the kernel above)
      const int t = warp >> 2;
      UDT_FSTAMP_DECL;
      if (t < ntiles) {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t o_addr = tmem_base + lane_base + kColO + t * 64;
        const float sl2 = p.scale_log2;
        int sb = t, su = 0;   // S buffer / use count of this warpgroup's next computation: continue across the items of the epoch
        for (int e = 0; e < E; ++e) {
          int b, h, q0, nt_unused;
          fmha_item(p, L + e * G, b, h, q0, nt_unused);
          const int g0 = e * nkv;   // key tiles before this item: p_full / o_full phases are counted per epoch
          float m_ref = -INFINITY, l = 0.0f;
          // rescale the running output (and row sum) when the reference maximum moves; O_t must be stable
          auto rescale = [&](bool need, float m_tile, int j) {
            const float m_new = need ? m_tile : m_ref;
            const float alpha = need ? ex2_approx(m_ref - m_new) : 1.0f;  // m_ref = -inf on the first tile -> 0
            l *= alpha;
            if (j > 0) {
              mbar_wait(&o_full[t], static_cast<uint32_t>((g0 + j - 1) & 1));
              tc_fence_after();
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
          };

          for (int j = 0; j < nkv; ++j) {
            UDT_FSTAMP(j, 0);
            mbar_wait(&s_full[sb], static_cast<uint32_t>(su & 1));
            tc_fence_after();
            UDT_FSTAMP(j, 1);
            const uint32_t s_addr = tmem_base + lane_base + kColS + sb * 128;
            const int key_lim = p.Nkv - j * kTile;  // keys >= key_lim of this tile are padding (only on the last tile)
            const bool partial = key_lim < kTile;
            float rowsum = 0.0f;
            bool replay = (j == 0) || partial;      // first / ragged tile: maximum first, then the exponentials
            // 32 probabilities (keys [32*ch, 32*ch+32) of this tile) -> fp16 -> this row's swizzled slots of the P buffer;
            // the stores interleave with the exponentials of the following chunk
            uint32_t pall[64];                      // the tile's probabilities, fp16 pairs (key 2i | key 2i+1)
            auto store_chunk = [&](const uint32_t (&pk)[16], int ch) {
    #pragma unroll
              for (int i = 0; i < 16; ++i) pall[ch * 16 + i] = pk[i];
            };
            if (!replay) {
              // ---- single pass: exponentials against the current reference, written to the P buffer right away.  No
              // maximum is tracked: every probability is bounded by 2^8 unless the tile's row sum exceeds 2^8, so a row sum
              // above that bound (rare: the reference would have to be stale by almost the whole lazy margin) sends the tile
              // through the two-pass path, which overwrites the optimistic P (nobody reads it before p_full).
              uint32_t va[32], vb[32];
              auto exp_chunk = [&](const uint32_t (&vv)[32], int ch) {
                uint32_t pk[16];
    #pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  float p0 = fmaf(__uint_as_float(vv[i]), sl2, -m_ref);
                  float p1 = fmaf(__uint_as_float(vv[i + 1]), sl2, -m_ref);
                  if (!UDT_FDBG(1)) {
                    p0 = (kPoly > 0 && (i % kPoly) == kPoly - 1) ? ex2_poly(p0) : ex2_approx(p0);
                    p1 = (kPoly > 0 && ((i + 1) % kPoly) == kPoly - 1) ? ex2_poly(p1) : ex2_approx(p1);
                  }
                  rowsum += p0 + p1;
                  pk[i >> 1] = pack_half2(p0, p1);
                }
                if (!UDT_FDBG(2)) store_chunk(pk, ch);
              };
              tmem_ld32(s_addr, va);
              tmem_ld_wait_dep(va);
              UDT_FSTAMP(j, 2);
              tmem_ld32(s_addr + 32, vb);      // the next 32 columns fly during the math
              exp_chunk(va, 0);
              UDT_FSTAMP(j, 3);
              tmem_ld_wait_dep(vb);
              tmem_ld32(s_addr + 64, va);
              exp_chunk(vb, 1);
              tmem_ld_wait_dep(va);
              tmem_ld32(s_addr + 96, vb);
              exp_chunk(va, 2);
              tmem_ld_wait_dep(vb);
              UDT_FSTAMP(j, 4);
              exp_chunk(vb, 3);
              UDT_FSTAMP(j, 5);
              replay = __any_sync(0xffffffffu, !(rowsum <= 256.0f)) && !UDT_FDBG(3);   // also catches inf / nan
            }
            if (replay) {
              // ---- two passes: row maximum of the raw scores, reference update (+ O rescale), exponentials
              uint32_t v[32];
              float mx = -INFINITY;
    #pragma unroll
              for (int ch = 0; ch < 4; ++ch) {
                tmem_ld32(s_addr + ch * 32, v);
                tmem_ld_wait();
    #pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (!partial || ch * 32 + i < key_lim) mx = fmaxf(mx, __uint_as_float(v[i]));
              }
              const float m_tile = mx * sl2;
              const bool need = m_tile > m_ref + kLazyThreshold;
              if (__any_sync(0xffffffffu, need)) rescale(need, m_tile, j);
              rowsum = 0.0f;
    #pragma unroll
              for (int ch = 0; ch < 4; ++ch) {
                tmem_ld32(s_addr + ch * 32, v);
                tmem_ld_wait();
                uint32_t pk[16];
    #pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), sl2, -m_ref));
                  float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sl2, -m_ref));
                  if (partial) {
                    const int k0 = ch * 32 + i;
                    if (k0 >= key_lim) p0 = 0.0f;
                    if (k0 + 1 >= key_lim) p1 = 0.0f;
                  }
                  rowsum += p0 + p1;
                  pk[i >> 1] = pack_half2(p0, p1);
                }
                store_chunk(pk, ch);
              }
            }
            l += rowsum;
            // ---- P -> TMEM, over the first 64 columns of the (fully consumed) score buffer: the PV MMA reads it as its A operand
            {
              uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&pall[0]);
              uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&pall[32]);
              tmem_st32(s_addr, lo);
              tmem_st32(s_addr + 32, hi);
              tmem_st_wait();
            }
            UDT_FSTAMP(j, 6);
            // observe every o_full phase (the wait is almost always already satisfied: P_t V_{j-1} ran during this tile's
            // exponentials); a parity wait that skipped phases would be ambiguous in rescale() and at the end
            if (j > 0) mbar_wait(&o_full[t], static_cast<uint32_t>((g0 + j - 1) & 1));
            tc_fence_before();         // TMEM reads of S / writes of P, O_t ordered before the arrive
            mbar_arrive(&p_full[t]);
            UDT_FSTAMP(j, 7);
            sb += ntiles;
            if (sb >= kSBufs) { sb -= kSBufs; ++su; }
          }
          mbar_wait(&o_full[t], static_cast<uint32_t>((g0 + nkv - 1) & 1));
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
          tc_fence_before();         // the tcgen05.ld of O_t ordered before the arrive
          mbar_arrive(&o_free[t]);
        }
      }
    }
    L += E * G;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}
#endif  // UDT_TUNING (persistent variant)

#ifdef UDT_TUNING
// ------------------------------------------------------------------------------------------------------------------
// EXPERIMENT (tuning builds only, UDT_FMHA_HT=1) — measured on B200: correct (incl. replay-heavy and ragged cases) and 22 %
// SLOWER than the production kernel (4096 tokens 244 -> 299 us, batch 32 1808 -> 2213 us): decoupling the two warpgroups
// does not pay for twice the barrier round trips and tcgen05.ld / st waits per key.
// Half-tile schedule: the score tiles are 128 queries x 64 keys and every softmax warpgroup owns THREE score buffers of
// its own.  Why: source-level warp sampling of udt_fmha_ts_kernel shows 14 % of all samples on the `s_full` spin — in that
// kernel the three 128-column score buffers rotate between the two warpgroups, S(c+3) can only be issued after P(c) V (P(c)
// lives in the buffer S(c+3) overwrites), and c / c+3 belong to DIFFERENT warpgroups, so each warpgroup's next scores wait
// for the other one's progress.  Tensor memory is full (384 score + 128 output columns), a fourth 128-column buffer does not
// fit — but six 64-column buffers do: warpgroup t owns columns [t*192, t*192+192) as buffers b = 0..2, the scores of its
// half-tile hc+3 reuse the buffer of ITS OWN half-tile hc, and the MMA warp serves whichever warpgroup has its probabilities
// ready (non-blocking test of both p_full barriers).  The two warpgroups only share the K / V ring and the tensor pipe.
// Same thread-per-row softmax as above (exp2 domain, lazy rescale, single pass with replay), 64 keys per step.
constexpr int kHBufs = 3;

__global__ void __launch_bounds__(kThreads, 1) udt_fmha_h_kernel(const __grid_constant__ FmhaParams p) {
  griddep_launch();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_addr - raw_addr);

  uint64_t* q_full = reinterpret_cast<uint64_t*>(base + kOffCtrl);
  uint64_t* kv_full = q_full + 1;                  // [kKvStages]
  uint64_t* kv_empty = kv_full + kKvStages;        // [kKvStages]  one arrival per warpgroup (its P V of the tile's second half)
  uint64_t* s_full = kv_empty + kKvStages;         // [2][kHBufs]
  uint64_t* p_full = s_full + 2 * kHBufs;          // [2]
  uint64_t* o_full = p_full + 2;                   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int b, h, q0, ntiles;
  if (!fmha_work(p, b, h, q0, ntiles)) return;
  const int nkv = (p.Nkv + kTile - 1) / kTile;
  const int H = 2 * nkv;                           // half-tiles per warpgroup

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&p.mapQ);
    tma_prefetch_desc(&p.mapK);
    tma_prefetch_desc(&p.mapV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], static_cast<uint32_t>(ntiles));
    }
    for (int i = 0; i < 2 * kHBufs; ++i) mbar_init(&s_full[i], 1);
    for (int i = 0; i < 2; ++i) {
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
  griddep_wait();

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int col = h * kD;
      mbar_expect_tx(q_full, static_cast<uint32_t>(ntiles * kTileBytes));
      for (int t = 0; t < ntiles; ++t)
        tma_load_2d(&p.mapQ, q_full, base + kOffQ + t * kTileBytes, col, b * p.Nq + q0 + t * kTile);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&kv_empty[s], ph ^ 1u);
        mbar_expect_tx(&kv_full[s], 2u * kTileBytes);
        tma_load_2d(&p.mapK, &kv_full[s], base + kTsOffK + s * kTileBytes, col, b * p.Nkv + j * kTile);
        tma_load_2d(&p.mapV, &kv_full[s], base + kTsOffV + s * kTileBytes, col, b * p.Nkv + j * kTile);
        if (++s == kKvStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    const bool issuer = elect_one();
    const uint32_t idesc_s = umma_idesc_f16(128, 64, false, false);
    const uint32_t idesc_o = umma_idesc_f16(128, 64, false, true);  // B = V is MN-major
    int kv_seen = -1;                              // highest key tile whose K / V are known to have landed
    int hs[2] = {0, 0}, hp[2] = {0, 0};            // next half-tile whose scores / whose P V is to be issued, per warpgroup
    // S(t, hs[t]) = Q_t K_j[half]^T -> buffer hs[t] % 3 of warpgroup t, if that buffer is free (its previous P V has been
    // issued) and key tile j has landed.  Never blocks: a warpgroup that runs ahead of the K / V ring must not keep the MMA
    // warp from serving the other one (whose progress frees the ring stage it is waiting for).
    auto try_issue_s = [&](int t) -> bool {
      if (hs[t] >= H || hs[t] >= hp[t] + kHBufs) return false;
      const int hc = hs[t], j = hc >> 1, half = hc & 1, bf = hc % kHBufs;
      while (kv_seen < j) {
        const int nx = kv_seen + 1;
        if (!__all_sync(0xffffffffu, mbar_test_wait(&kv_full[nx % kKvStages], static_cast<uint32_t>((nx / kKvStages) & 1))))
          return false;
        kv_seen = nx;
      }
      tc_fence_after();
      if (issuer) {
        const uint64_t dq = umma_desc_kmajor_sw128(base_addr + kOffQ + t * kTileBytes);
        const uint64_t dk = umma_desc_kmajor_sw128(base_addr + kTsOffK + (j % kKvStages) * kTileBytes + half * (64 * 128));
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk)
          umma_f16_ss(tmem_base + t * 192 + bf * 64, dq + static_cast<uint64_t>(kk * 2), dk + static_cast<uint64_t>(kk * 2),
                      idesc_s, kk != 0 ? 1u : 0u);
        umma_commit(&s_full[t * kHBufs + bf]);
      }
      __syncwarp();
      ++hs[t];
      return true;
    };
    mbar_wait(q_full, 0);
    int remaining = ntiles * H;
    uint32_t spins = 0;
    while (remaining > 0) {
      bool progressed = false;
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (t < ntiles && try_issue_s(t)) progressed = true;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t >= ntiles || hp[t] >= H) continue;
        const bool ready = __all_sync(0xffffffffu, mbar_test_wait(&p_full[t], static_cast<uint32_t>(hp[t] & 1)));
        if (!ready) continue;
        const int hc = hp[t], j = hc >> 1, half = hc & 1, bf = hc % kHBufs;
        tc_fence_after();
        if (issuer) {
          const uint32_t v_addr = base_addr + kTsOffV + (j % kKvStages) * kTileBytes;
          const uint32_t p_tmem = tmem_base + t * 192 + bf * 64;   // P (fp16 pairs) over the first 32 columns of its score buffer
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t dv = umma_desc_mnmajor_sw128(v_addr + (half * 4 + kk) * 16 * 128, 8192);
            umma_f16_ts(tmem_base + kColO + t * 64, p_tmem + kk * 8, dv, idesc_o, (hc | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&o_full[t]);
          if (half == 1) umma_commit(&kv_empty[j % kKvStages]);   // this warpgroup is done with key tile j
        }
        __syncwarp();
        ++hp[t];
        --remaining;
        progressed = true;
      }
      if (progressed) spins = 0;
      else if (++spins > UDT_SPIN_LIMIT) asm volatile("trap;");
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups (thread = query row)
    const int t = warp >> 2;
    if (t < ntiles) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t o_addr = tmem_base + lane_base + kColO + t * 64;
      const float sl2 = p.scale_log2;
      float m_ref = -INFINITY, l = 0.0f;

      auto rescale = [&](bool need, float m_tile, int hc) {
        const float m_new = need ? m_tile : m_ref;
        const float alpha = need ? ex2_approx(m_ref - m_new) : 1.0f;
        l *= alpha;
        if (hc > 0) {
          mbar_wait(&o_full[t], static_cast<uint32_t>((hc - 1) & 1));
          tc_fence_after();
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
      };

      int bf = 0, use = 0;     // bf = hc % 3, use = hc / 3
      for (int hc = 0; hc < H; ++hc) {
        mbar_wait(&s_full[t * kHBufs + bf], static_cast<uint32_t>(use & 1));
        tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_base + t * 192 + bf * 64;
        const int key_lim = p.Nkv - (hc >> 1) * kTile - (hc & 1) * 64;   // my keys >= key_lim are padding (last tile only)
        const bool partial = key_lim < 64;
        float rowsum = 0.0f;
        bool replay = (hc == 0) || partial;
        uint32_t pall[32];
        if (!replay) {
          uint32_t va[32], vb[32];
          tmem_ld32(s_addr, va);
          tmem_ld_wait_dep(va);
          tmem_ld32(s_addr + 32, vb);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(va[i]), sl2, -m_ref));
            const float p1 = ex2_approx(fmaf(__uint_as_float(va[i + 1]), sl2, -m_ref));
            rowsum += p0 + p1;
            pall[i >> 1] = pack_half2(p0, p1);
          }
          tmem_ld_wait_dep(vb);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(vb[i]), sl2, -m_ref));
            const float p1 = ex2_approx(fmaf(__uint_as_float(vb[i + 1]), sl2, -m_ref));
            rowsum += p0 + p1;
            pall[16 + (i >> 1)] = pack_half2(p0, p1);
          }
          replay = __any_sync(0xffffffffu, !(rowsum <= 256.0f));   // stale reference (or inf / nan): two-pass path
        }
        if (replay) {
          uint32_t v[32];
          float mx = -INFINITY;
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            tmem_ld32(s_addr + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!partial || ch * 32 + i < key_lim) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
          const float m_tile = mx * sl2;
          const bool need = m_tile > m_ref + kLazyThreshold;
          if (__any_sync(0xffffffffu, need)) rescale(need, m_tile, hc);
          rowsum = 0.0f;
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            tmem_ld32(s_addr + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), sl2, -m_ref));
              float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sl2, -m_ref));
              if (partial) {
                const int k0 = ch * 32 + i;
                if (k0 >= key_lim) p0 = 0.0f;
                if (k0 + 1 >= key_lim) p1 = 0.0f;
              }
              rowsum += p0 + p1;
              pall[ch * 16 + (i >> 1)] = pack_half2(p0, p1);
            }
          }
        }
        l += rowsum;
        tmem_st32(s_addr, pall);                 // P over the first 32 columns of my own (consumed) score buffer
        tmem_st_wait();
        if (hc > 0) mbar_wait(&o_full[t], static_cast<uint32_t>((hc - 1) & 1));   // observe every o_full phase
        tc_fence_before();
        mbar_arrive(&p_full[t]);
        if (++bf == kHBufs) { bf = 0; ++use; }
      }
      mbar_wait(&o_full[t], static_cast<uint32_t>((H - 1) & 1));
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

// ------------------------------------------------------------------------------------------------------------------
// EXPERIMENT (tuning builds only, UDT_FMHA_W4=1) — measured on B200: correct (same errors to 4 digits incl. replay and ragged
// tiles) and 7 % SLOWER than the two-warpgroup kernel (4096 tokens 244 -> 260 us, 1024 tokens 45.2 -> 49.1 us): the kernel is
// not limited by the number of warps a scheduler can pick from.
// Four softmax warpgroups: every 128x128 score tile is shared by TWO warpgroups, each thread owns one query row and HALF of
// the tile's keys (64 columns).  Why: ncu shows the two-warpgroup kernel with 2.5 active warps per scheduler of which 0.57
// are eligible (58 % of the cycles no warp can issue) — the tcgen05.ld -> FMA -> MUFU -> pack chain of one row of 128 keys
// is latency-bound with two warps per sub-partition.  Twice the warps on the same tensor-memory lanes (warp w and w + 4 share
// lanes 32 * (w % 4)) halve the chain per thread and double the warps a scheduler can pick from; registers per thread drop
// from 168 to <= 112 (32 packed probabilities instead of 64).
//   warps 0-3 / 4-7    : tile 0, keys [0, 64) / [64, 128) of every key tile        warp 16 : MMA issuer
//   warps 8-11 / 12-15 : tile 1, keys [0, 64) / [64, 128)                          warp 17 : TMA producer
// The two halves of a row agree once per key tile through a 64-thread named barrier + two shared-memory words: whether the
// optimistic single pass violated the 2^8 bound (then both replay), the row maximum of a replayed / first tile, and at the
// end the two partial row sums.  Each half writes its probabilities over score columns it has itself consumed — keys
// [0, 64) -> columns [0, 32), keys [64, 128) -> columns [64, 96) — so no half ever overwrites scores the other still reads;
// the PV MMA reads its A operand from those two column ranges.
constexpr int kW4Threads = 576;
constexpr int kW4OffX = 1024;                       // exchange area: [2 parity][2 tiles][2 halves][128 rows] floats + flags
constexpr int kW4XBytes = 2 * 2 * 2 * 128 * 4 + 256;
constexpr int kW4OffQ = 6144;                       // 1024-aligned operand tiles start here
constexpr int kW4OffK = kW4OffQ + 2 * kTileBytes;
constexpr int kW4OffV = kW4OffK + kKvStages * kTileBytes;
constexpr int kW4SmemBytes = kW4OffV + kKvStages * kTileBytes + 1024;
static_assert(kW4OffX + kW4XBytes <= kW4OffQ, "exchange area overlaps the Q tiles");

__global__ void __launch_bounds__(kW4Threads, 1) udt_fmha_w4_kernel(const __grid_constant__ FmhaParams p) {
  griddep_launch();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_addr - raw_addr);

  uint64_t* q_full = reinterpret_cast<uint64_t*>(base + kOffCtrl);
  uint64_t* kv_full = q_full + 1;              // [kKvStages]
  uint64_t* kv_empty = kv_full + kKvStages;    // [kKvStages]
  uint64_t* s_full = kv_empty + kKvStages;     // [kSBufs]
  uint64_t* p_full = s_full + kSBufs;          // [2]  both halves of P_t(j) are in tensor memory (256 arrivals)
  uint64_t* o_full = p_full + 2;               // [2]  P_t(j) V_j has been accumulated into O_t
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);
  float* xval = reinterpret_cast<float*>(base + kW4OffX);                    // [parity][tile][half][row]
  int* xflag = reinterpret_cast<int*>(base + kW4OffX + 2 * 2 * 2 * 128 * 4);   // [parity][tile][half][quarter]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int b, h, q0, ntiles;
  if (!fmha_work(p, b, h, q0, ntiles)) return;
  const int nkv = (p.Nkv + kTile - 1) / kTile;
  const int ncomp = nkv * ntiles;

  if (warp == 17 && lane == 0) {
    tma_prefetch_desc(&p.mapQ);
    tma_prefetch_desc(&p.mapK);
    tma_prefetch_desc(&p.mapV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < kSBufs; ++i) mbar_init(&s_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 256);
      mbar_init(&o_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();

  if (warp == 17) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int col = h * kD;
      mbar_expect_tx(q_full, static_cast<uint32_t>(ntiles * kTileBytes));
      for (int t = 0; t < ntiles; ++t)
        tma_load_2d(&p.mapQ, q_full, base + kW4OffQ + t * kTileBytes, col, b * p.Nq + q0 + t * kTile);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&kv_empty[s], ph ^ 1u);
        mbar_expect_tx(&kv_full[s], 2u * kTileBytes);
        tma_load_2d(&p.mapK, &kv_full[s], base + kW4OffK + s * kTileBytes, col, b * p.Nkv + j * kTile);
        tma_load_2d(&p.mapV, &kv_full[s], base + kW4OffV + s * kTileBytes, col, b * p.Nkv + j * kTile);
        if (++s == kKvStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 16) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    const bool issuer = elect_one();
    const uint32_t idesc_s = umma_idesc_f16(128, 128, false, false);
    const uint32_t idesc_o = umma_idesc_f16(128, 64, false, true);  // B = V is MN-major
    auto issue_s = [&](const CompCursor& k) {
      if (k.t == 0) mbar_wait(&kv_full[k.stage], static_cast<uint32_t>(k.kvphase));
      tc_fence_after();
      if (issuer) {
        const uint64_t dq = umma_desc_kmajor_sw128(base_addr + kW4OffQ + k.t * kTileBytes);
        const uint64_t dk = umma_desc_kmajor_sw128(base_addr + kW4OffK + k.stage * kTileBytes);
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk)
          umma_f16_ss(tmem_base + kColS + k.b * 128, dq + static_cast<uint64_t>(kk * 2), dk + static_cast<uint64_t>(kk * 2),
                      idesc_s, kk != 0 ? 1u : 0u);
        umma_commit(&s_full[k.b]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    CompCursor ks, kp;
    ks.init();
    kp.init();
    for (int i = 0; i < kSBufs && ks.c < ncomp; ++i) {
      issue_s(ks);
      ks.advance(ntiles);
    }
    for (; kp.c < ncomp; kp.advance(ntiles)) {
      mbar_wait(&p_full[kp.t], static_cast<uint32_t>(kp.j & 1));
      tc_fence_after();
      if (issuer) {
        const uint32_t v_addr = base_addr + kW4OffV + kp.stage * kTileBytes;
        const uint32_t p_tmem = tmem_base + kColS + kp.b * 128;
#pragma unroll
        for (int kk = 0; kk < kTile / 16; ++kk) {
          const uint64_t dv = umma_desc_mnmajor_sw128(v_addr + kk * 16 * 128, 8192);
          // keys [0, 64) of P live in columns [0, 32), keys [64, 128) in columns [64, 96) (8 columns per 16 keys)
          const uint32_t pa = p_tmem + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8);
          umma_f16_ts(tmem_base + kColO + kp.t * 64, pa, dv, idesc_o, (kp.j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&o_full[kp.t]);
        if (kp.t == ntiles - 1) umma_commit(&kv_empty[kp.stage]);
      }
      __syncwarp();
      if (ks.c < ncomp) {
        issue_s(ks);
        ks.advance(ntiles);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups (thread = row x half of the keys)
    const int t = warp >> 3;           // query tile
    const int half = (warp >> 2) & 1;  // key half of every tile
    if (t < ntiles) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t o_addr = tmem_base + lane_base + kColO + t * 64 + half * 32;   // this half rescales / writes 32 O columns
      const uint32_t bar_id = 1u + static_cast<uint32_t>(t * 4 + quarter);           // named barrier of this warp and its partner
      const float sl2 = p.scale_log2;
      float m_ref = -INFINITY, l = 0.0f;
      int sb = t, su = 0;
      // exchange slots, double buffered by tile parity (a warp is never two tiles ahead of its partner)
      auto xv = [&](int par, int hf) -> float& { return xval[((par * 2 + t) * 2 + hf) * 128 + row]; };
      auto xf = [&](int par, int hf) -> int& { return xflag[((par * 2 + t) * 2 + hf) * 4 + quarter]; };

      auto rescale = [&](bool need, float m_tile, int j) {
        const float m_new = need ? m_tile : m_ref;
        const float alpha = need ? ex2_approx(m_ref - m_new) : 1.0f;
        l *= alpha;
        if (j > 0) {
          mbar_wait(&o_full[t], static_cast<uint32_t>((j - 1) & 1));
          tc_fence_after();
          uint32_t v[32];
          tmem_ld32(o_addr, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st32(o_addr, v);
          tmem_st_wait();
        }
        m_ref = m_new;
      };

      for (int j = 0; j < nkv; ++j) {
        const int par = j & 1;
        mbar_wait(&s_full[sb], static_cast<uint32_t>(su & 1));
        tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_base + kColS + sb * 128 + half * 64;   // this half's 64 score columns
        const int key_lim = p.Nkv - j * kTile - half * 64;   // my keys >= key_lim are padding (only on the last tile)
        const bool partial = (p.Nkv - j * kTile) < kTile;    // uniform over both halves
        float rowsum = 0.0f;
        bool replay = (j == 0) || partial;
        uint32_t pall[32];                                   // my 64 probabilities, fp16 pairs
        if (!replay) {
          uint32_t va[32], vb[32];
          tmem_ld32(s_addr, va);
          tmem_ld_wait_dep(va);
          tmem_ld32(s_addr + 32, vb);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(va[i]), sl2, -m_ref));
            const float p1 = ex2_approx(fmaf(__uint_as_float(va[i + 1]), sl2, -m_ref));
            rowsum += p0 + p1;
            pall[i >> 1] = pack_half2(p0, p1);
          }
          tmem_ld_wait_dep(vb);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(vb[i]), sl2, -m_ref));
            const float p1 = ex2_approx(fmaf(__uint_as_float(vb[i + 1]), sl2, -m_ref));
            rowsum += p0 + p1;
            pall[16 + (i >> 1)] = pack_half2(p0, p1);
          }
          // both halves of a row must agree: a violation in either sends both through the two-pass path
          const bool viol = __any_sync(0xffffffffu, !(rowsum <= 256.0f));
          if (lane == 0) xf(par, half) = viol ? 1 : 0;
          named_bar_sync(bar_id, 64);
          replay = (xf(par, 0) | xf(par, 1)) != 0;
        }
        if (replay) {
          uint32_t v[32];
          float mx = -INFINITY;
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            tmem_ld32(s_addr + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!partial || ch * 32 + i < key_lim) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
          xv(par, half) = mx;
          named_bar_sync(bar_id, 64);
          const float m_tile = fmaxf(xv(par, 0), xv(par, 1)) * sl2;     // full-row maximum, identical in both halves
          const bool need = m_tile > m_ref + kLazyThreshold;
          if (__any_sync(0xffffffffu, need)) rescale(need, m_tile, j);
          rowsum = 0.0f;
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            tmem_ld32(s_addr + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), sl2, -m_ref));
              float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sl2, -m_ref));
              if (partial) {
                const int k0 = ch * 32 + i;
                if (k0 >= key_lim) p0 = 0.0f;
                if (k0 + 1 >= key_lim) p1 = 0.0f;
              }
              rowsum += p0 + p1;
              pall[ch * 16 + (i >> 1)] = pack_half2(p0, p1);
            }
          }
        }
        l += rowsum;
        // my P half -> the first 32 columns of my own (consumed) 64 score columns
        tmem_st32(s_addr, pall);
        tmem_st_wait();
        if (j > 0) mbar_wait(&o_full[t], static_cast<uint32_t>((j - 1) & 1));   // observe every o_full phase
        tc_fence_before();
        mbar_arrive(&p_full[t]);
        sb += ntiles;
        if (sb >= kSBufs) { sb -= kSBufs; ++su; }
      }
      mbar_wait(&o_full[t], static_cast<uint32_t>((nkv - 1) & 1));
      tc_fence_after();
      // row sum = the two partial sums (both relative to the same reference maximum)
      const int parE = nkv & 1;
      xv(parE, half) = l;
      named_bar_sync(bar_id, 64);
      const float inv = 1.0f / (xv(parE, 0) + xv(parE, 1));
      const int qrow = q0 + t * kTile + row;
      uint4* o4 = reinterpret_cast<uint4*>(p.o + (static_cast<size_t>(b) * p.Nq + min(qrow, p.Nq - 1)) * p.ldo + h * kD + half * 32);
      uint32_t v[32];
      tmem_ld32(o_addr, v);
      tmem_ld_wait();
      if (qrow < p.Nq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 ov;
          ov.x = pack_half2(__uint_as_float(v[g * 8 + 0]) * inv, __uint_as_float(v[g * 8 + 1]) * inv);
          ov.y = pack_half2(__uint_as_float(v[g * 8 + 2]) * inv, __uint_as_float(v[g * 8 + 3]) * inv);
          ov.z = pack_half2(__uint_as_float(v[g * 8 + 4]) * inv, __uint_as_float(v[g * 8 + 5]) * inv);
          ov.w = pack_half2(__uint_as_float(v[g * 8 + 6]) * inv, __uint_as_float(v[g * 8 + 7]) * inv);
          o4[g] = ov;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

#endif  // UDT_TUNING

}  // namespace

extern "C" int udt_fmha_fwd(const void* q, const void* k, const void* v, void* o, int32_t B, int32_t Nq, int32_t Nkv,
                            int32_t heads, int32_t ldq, int32_t ldk, int32_t ldv, int32_t ldo, float scale,
                            void* stream) {
  using namespace udt_host;
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (B < 1 || Nq < 1 || Nkv < 1 || heads < 1) return fail(UDT_ERR_SHAPE, "udt_fmha_fwd: bad shape");
  if (ldo % 8 || (reinterpret_cast<uintptr_t>(o) & 15)) return fail(UDT_ERR_ALIGN, "udt_fmha_fwd: o / ldo alignment");
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
  static const int dbg = tune_int("UDT_FMHA_DEBUG", 0);
  p.debug = dbg;
  p.pairs_per_bh = (Nq + 2 * kTile - 1) / (2 * kTile);
  const long npairs = static_cast<long>(p.pairs_per_bh) * heads * B;
  if (npairs > (1l << 30)) return fail(UDT_ERR_SHAPE, "udt_fmha_fwd: too many query tiles");
  // tail wave: if its R pairs occupy at most half of the SMs, run them as 2R single-tile CTAs (tuning builds: UDT_FMHA_TAIL=0)
  static const int tail_split = tune_int("UDT_FMHA_TAIL", 1);
  const int nsm = num_sms();
  const int tail = static_cast<int>(npairs % nsm);
  p.pairs_full = static_cast<int32_t>((tail_split && tail > 0 && 2 * tail <= nsm) ? npairs - tail : npairs);
  dim3 grid(static_cast<unsigned>(p.pairs_full + 2 * (npairs - p.pairs_full)), 1, 1);
  // tuning builds: UDT_FMHA_POLY=4 moves every 4th exponential from MUFU to the FMA pipe.  Measured on B200: no gain
  // (4096 tokens 249.0 -> 250.4 us; every 2nd: 275 us) — the kernel is bound by the TMEM read path as much as by MUFU
#ifdef UDT_TUNING
  // experiment: half-tile schedule (every softmax warpgroup owns three 64-key score buffers), UDT_FMHA_HT=1 — 22 % slower
  static const int use_ht = tune_int("UDT_FMHA_HT", 0);
  if (use_ht) {
    static bool ht_attr = false;
    if (!ht_attr) {
      cudaError_t e = cudaFuncSetAttribute(udt_fmha_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmemBytes);
      if (e != cudaSuccess) return fail(UDT_ERR_LAUNCH, "cudaFuncSetAttribute(fmha ht smem): %s", cudaGetErrorString(e));
      ht_attr = true;
    }
    udt_host::launch_pdl(udt_fmha_h_kernel, dim3(grid), dim3(kThreads), kTsSmemBytes, reinterpret_cast<cudaStream_t>(stream), p);
    return check_launch("udt_fmha_fwd (half tiles)");
  }
#endif
#ifdef UDT_TUNING
  // experiment: four softmax warpgroups (two per score tile), UDT_FMHA_W4=1 — measured 7 % slower, see the kernel's header
  static const int use_w4 = tune_int("UDT_FMHA_W4", 0);
  if (use_w4) {
    static bool w4_attr = false;
    if (!w4_attr) {
      cudaError_t e = cudaFuncSetAttribute(udt_fmha_w4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kW4SmemBytes);
      if (e != cudaSuccess) return fail(UDT_ERR_LAUNCH, "cudaFuncSetAttribute(fmha w4 smem): %s", cudaGetErrorString(e));
      w4_attr = true;
    }
    udt_host::launch_pdl(udt_fmha_w4_kernel, dim3(grid), dim3(kW4Threads), kW4SmemBytes, reinterpret_cast<cudaStream_t>(stream), p);
    return check_launch("udt_fmha_fwd (w4)");
  }
#endif
  p.num_items = static_cast<int32_t>(grid.x);
#ifdef UDT_TUNING
  static const int persist = tune_int("UDT_FMHA_PERSIST", 0);
  if (persist) {   // persistent CTAs: one per SM, items round-robin
    static bool pk_attr = false;
    if (!pk_attr) {
      cudaError_t e = cudaFuncSetAttribute(udt_fmha_pk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPkSmemBytes);
      if (e != cudaSuccess) return fail(UDT_ERR_LAUNCH, "cudaFuncSetAttribute(fmha persistent smem): %s", cudaGetErrorString(e));
      pk_attr = true;
    }
    const unsigned g = grid.x < static_cast<unsigned>(nsm) ? grid.x : static_cast<unsigned>(nsm);
    udt_host::launch_pdl(udt_fmha_pk_kernel, dim3(g), dim3(kThreads), kPkSmemBytes, reinterpret_cast<cudaStream_t>(stream), p);
    return check_launch("udt_fmha_fwd (persistent)");
  }
#endif
  static const int poly = tune_int("UDT_FMHA_POLY", 0);
  void (*kern)(FmhaParams) = poly == 0 ? udt_fmha_ts_kernel<0> : udt_fmha_ts_kernel<4>;
  static bool ts_attr = false;
  if (!ts_attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmemBytes);
    if (e != cudaSuccess) return fail(UDT_ERR_LAUNCH, "cudaFuncSetAttribute(fmha smem): %s", cudaGetErrorString(e));
    ts_attr = true;
  }
  udt_host::launch_pdl(kern, dim3(grid), dim3(kThreads), kTsSmemBytes, reinterpret_cast<cudaStream_t>(stream), p);
  return check_launch("udt_fmha_fwd");
}
