// udt_igemm.cu — K2/K3: segmented implicit GEMM on tcgen05 tensor cores (sm_100a).
//
//   out[m, n] = act( sum_{seg,tap,c} A_seg[pixel(m) + tap, c] * Wt[n, k] + bias[n] + rowbias[img(m), n] ) + res[m, n]
//
// Two launch shapes of the same kernel:
//   pair   : clusters of two CTAs (one per SM of a TPC) drive tcgen05 cta_group::2 — one M = 256 tile per pair, each
//            CTA loads its own 128 rows of A and HALF of the weight tile, which halves the per-SM operand traffic
//            (the per-SM TMA ingest rate, ~73 B/clk measured, is what bounds these GEMMs); the leader CTA's MMA
//            thread issues for both SMs, commits are multicast to both CTAs' barriers.
//   single : one CTA per tile (tiny problems, fp32 / narrow outputs).
// One persistent CTA per SM, warp-specialised:
//   warp 0      A producer     : per K step one 4-D TMA box {64 ch, bw, bh, bn} of the NHWC activation (shifted by the
//                                3x3 tap, out-of-bounds zero filled = conv padding), 128B-swizzled.
//   warp 10     B producer     : per K step one 2-D TMA box {64, BN} of the K-major fp16 weights.  Two issuing threads
//                                because ONE thread sustains only one TMA instruction per ~250 clk (measured) — with a
//                                single producer every K step cost ~520 clk regardless of its size.
//   warp 1      MMA issuer     : tcgen05.mma.cta_group::1.kind::f16, M = 128, N = BN, K = 16, fp32 accumulators
//                                in TMEM, two accumulator buffers so the epilogue of tile i overlaps tile i+1.
//   warps 2..5  epilogue       : tcgen05.ld (thread = output row), bias / per-image bias / SiLU / ReLU / GEGLU /
//                                residual.  fp16 outputs are staged: 32-column chunks go registers -> 64B-swizzled
//                                shared memory -> TMA tile store (coalesced, clipped at the tensor edge), the
//                                residual chunks are TMA-prefetched two chunks ahead (across tile boundaries) and
//                                the bias row is staged in shared memory once per tile.  Tiny / fp32 outputs
//                                (N_out < 8, out_fp32, BN = 16) use direct global stores.
// Pipelines: smem full/empty mbarrier ring (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), residual full.
//
// Replaces the cuDNN / cuBLAS call sites listed in include/udt_api.h (udt_igemm).
#include "udt_common.cuh"
#include "udt_host.h"
#include <stdlib.h>

namespace {

using namespace udt;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                         // fp16 elements per K step = one 128B swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kMaxStages = 8;
constexpr int kThreads = 352;                        // A-TMA warp, MMA warp, 2 x 4 epilogue warps, B-TMA warp
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;                     // TMEM column offset of the second accumulator
constexpr int kSmemBudget = 227 * 1024;
constexpr int kCtrlBytes = 1024;
constexpr int kGegluTile = 256;                      // preferred GEGLU weight interleave (128 is accepted via bn_hint)
constexpr int kChunkCols = 32;                      // epilogue staging granularity: 32 fp16 columns = 64 B rows
constexpr int kChunkBytes = kBlockM * kChunkCols * 2;  // 8 KB per 128-row chunk = 4 warp slabs
constexpr int kSlabBytes = 32 * kChunkCols * 2;        // 2 KB: one warp's 32 rows x 32 fp16 columns
constexpr int kBiasImgs = 4;                        // per-image bias rows staged per tile (tiles spanning more images
                                                    // read rowbias from global memory)
constexpr int kBiasBytes = 2 * kBiasImgs * 256 * 4; // 8 KB: double buffered (the next tile's rows are fetched a tile ahead)
constexpr int kEpiBarrier = 1;                      // named barrier ids 1, 2: the two epilogue warpgroups
constexpr int kMaxResBufs = 4;

struct IGemmParams {
  CUtensorMap mapA[3];
  CUtensorMap mapB;
  int32_t seg_kc[3];
  int32_t seg_taps[3];
  int32_t seg_stride[3];
  int32_t seg_pad[3];
  int32_t nseg, ksteps;
  int32_t W, H, NB;
  int32_t bw, bh, bn;
  int32_t lbw, lbh;      // log2(bw), log2(bh): the spatial tile sides are powers of two
  int32_t tiles_w, tiles_h, tiles_nb, tiles_n, num_tiles;
  uint32_t mg_n, mg_w, mg_h;   // fast_div magic numbers of tiles_n, tiles_w, tiles_h
  // split-K (small-M, weight-streaming problems): work item = (tile, split); split s accumulates K steps
  // [s*kper, min(ksteps, (s+1)*kper)) and stores its fp32 partial tile at out + s*split_stride (reduced by a second kernel)
  int32_t ksplit, kper;
  uint32_t mg_s;
  long long split_stride;
  int32_t b_img_rows;     // per-image weights: image i reads weight rows [i * b_img_rows, ...) (0 = one shared weight)
  int32_t N_out, BN, stages;
  const float* bias;
  const float* rowbias;
  const __half* residual;
  int32_t ldr, ld_rowbias;
  void* out;
  int32_t ldo, out_fp32, act;
  int32_t debug;         // UDT_IGEMM_DEBUG experiment flags (0 in production)
  int32_t staged;        // 1: smem-staged epilogue with TMA store (fp16 out), 0: direct global stores
  int32_t egroups;       // staged mode: 1 or 2 epilogue warpgroups share a tile's chunks (2 for short-K, epilogue-bound GEMMs)
  int32_t out_bufs;      // staged mode: output chunk buffers per group (2 or 3)
  int32_t res_bufs;      // staged mode: residual chunk buffers per group (prefetch distance; 0 without residual)
  CUtensorMap mapOut;    // staged mode: {32, bw, bh, bn} boxes of the output tensor, 64B swizzle
  CUtensorMap mapRes;    // staged mode with residual: same boxes of the residual tensor
  unsigned long long* trace;  // udt_debug_set_trace(): per-CTA role timestamps (NULL in production)
};

// trace layout per CTA: [0] globaltimer at start, [1] at end, [2] clock64 at start, [3] at end, then per tile
// iteration kTraceEvents clock64 stamps (see scripts/igemm_trace.py)
constexpr int kTraceTiles = 24;
constexpr int kTraceEvents = 8;
constexpr int kTraceStride = 4 + kTraceTiles * kTraceEvents;

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#ifdef UDT_IGEMM_TRACE
constexpr bool kTraceBuild = true;    // role timestamps + UDT_IGEMM_DEBUG experiment switches compiled in (tuning builds)
#else
constexpr bool kTraceBuild = false;   // production: no trace / debug code in the hot loops
#endif
#define UDT_DBG(p, bit) (kTraceBuild && ((p).debug & (bit)))

__device__ __forceinline__ void trace_ev(const IGemmParams& p, int iter, int ev, bool fine = false) {
  if (kTraceBuild && p.trace != nullptr && iter < kTraceTiles && (((p.debug & 16) != 0) == fine || ev == 4))
    p.trace[static_cast<size_t>(blockIdx.x) * kTraceStride + 4 + iter * kTraceEvents + ev] =
        static_cast<unsigned long long>(clock64());
}

struct TileCoord {
  int w0, h0, n0, n_blk;
};

// n / d for n * d < 2^32 with magic = ceil(2^32 / d) (host: fast_div_magic)
__device__ __forceinline__ int fast_div(int n, uint32_t magic, int d) {
  return d == 1 ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), magic));
}

// `tile` counts (M tile or M-tile pair, N tile); `rank` = CTA rank within the pair (0 in single mode).  An odd tile
// count leaves a phantom M tile whose image coordinate is out of range: TMA zero-fills its loads and clips its stores.
template <bool kPair>
__device__ __forceinline__ TileCoord decode_tile(const IGemmParams& p, int tile, int rank) {
  TileCoord t;
  // exact division by multiply-high with host-computed magic numbers (a runtime integer division costs > 100 clk of
  // dependent latency, and this runs on the critical path of every role once per tile / residual prefetch)
  int m_blk = fast_div(tile, p.mg_n, p.tiles_n);
  t.n_blk = tile - m_blk * p.tiles_n;
  if (kPair) m_blk = 2 * m_blk + rank;
  const int r = fast_div(m_blk, p.mg_w, p.tiles_w);
  const int tw = m_blk - r * p.tiles_w;
  const int tn = fast_div(r, p.mg_h, p.tiles_h);
  const int th = r - tn * p.tiles_h;
  t.w0 = tw * p.bw;
  t.h0 = th * p.bh;
  t.n0 = tn * p.bn;
  return t;
}

struct WorkItem {
  int tile, split, k0, k1;
};
__device__ __forceinline__ WorkItem decode_work(const IGemmParams& p, int w) {
  WorkItem it;
  if (p.ksplit > 1) {
    it.tile = fast_div(w, p.mg_s, p.ksplit);
    it.split = w - it.tile * p.ksplit;
    it.k0 = it.split * p.kper;
    it.k1 = min(p.ksteps, it.k0 + p.kper);
  } else {
    it.tile = w;
    it.split = 0;
    it.k0 = 0;
    it.k1 = p.ksteps;
  }
  return it;
}

// Finish `cnt` (<= 32) consecutive output columns of one row: f[] holds the fp32 accumulators.
__device__ __forceinline__ void finish_columns(const IGemmParams& p, float (&f)[32], int cnt, size_t m, int img,
                                               int col, bool fast, int split) {
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < cnt) f[j] += __ldg(p.bias + col + j);
  }
  if (p.rowbias != nullptr) {
    const float* rb = p.rowbias + static_cast<size_t>(img) * p.ld_rowbias + col;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < cnt) f[j] += __ldg(rb + j);
  }
  if (p.act == UDT_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
  } else if (p.act == UDT_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
  } else if (p.act == UDT_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = gelu_erf_f(f[j]);
  }
  if (p.out_fp32) {
    float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(split) * p.split_stride + m * p.ldo + col;
    if (p.ksplit > 1 && cnt == 32 && (p.ldo & 3) == 0) {   // split-K partial tile: raw accumulators, 16-byte stores
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
      return;
    }
    if (p.residual != nullptr) {
      const __half* r = p.residual + m * p.ldr + col;
      for (int j = 0; j < cnt; ++j) f[j] += __half2float(r[j]);
    }
    for (int j = 0; j < cnt; ++j) o[j] = f[j];
    return;
  }
  __half* o = reinterpret_cast<__half*>(p.out) + m * p.ldo + col;
  if (fast && cnt == 32) {
    if (p.residual != nullptr) {
      const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + m * p.ldr + col);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint4 rv = r4[v];
        const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 rf = __half22float2(rh[j]);
          f[v * 8 + 2 * j] += rf.x;
          f[v * 8 + 2 * j + 1] += rf.y;
        }
      }
    }
    uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      uint4 ov;
      ov.x = pack_half2(f[v * 8 + 0], f[v * 8 + 1]);
      ov.y = pack_half2(f[v * 8 + 2], f[v * 8 + 3]);
      ov.z = pack_half2(f[v * 8 + 4], f[v * 8 + 5]);
      ov.w = pack_half2(f[v * 8 + 6], f[v * 8 + 7]);
      o4[v] = ov;
    }
    return;
  }
  if (p.residual != nullptr) {
    const __half* r = p.residual + m * p.ldr + col;
    for (int j = 0; j < cnt; ++j) f[j] += __half2float(r[j]);
  }
  for (int j = 0; j < cnt; ++j) o[j] = __float2half_rn(f[j]);
}

// kMode selects the epilogue that is compiled in (the hot loop carries no dead variants):
//   0 staged, bias / staged per-image bias / residual      1 staged GEGLU
//   2 staged, general (SiLU / ReLU / per-image bias of tiles spanning many images)      3 direct global stores
template <bool kPair, int kMode>
__global__ void __launch_bounds__(kThreads, 1) udt_igemm_kernel(const __grid_constant__ IGemmParams p) {
  constexpr bool kStaged = (kMode != 3);
  constexpr bool kGeglu = (kMode == 1);
  constexpr bool kGeneral = (kMode == 2);
  griddep_launch();   // PDL: the next kernel of the stream may start its prologue while this one runs
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_addr - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;           // [8 epilogue warps][kMaxResBufs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 8 * kMaxResBufs);

  // staged epilogue buffers (per epilogue group: out chunks, residual chunks, bias rows) sit between the control
  // block and the operand ring
  const bool has_res_stage = kStaged && !kGeglu && p.residual != nullptr;
  const int group_bytes = (p.out_bufs + p.res_bufs) * kChunkBytes + kBiasBytes;
  const int epi_bytes = kStaged ? p.egroups * group_bytes : (p.ksplit > 1 ? 8 * 4096 : 0);   // split-K: transpose scratch

  const int b_rows = kPair ? p.BN / 2 : p.BN;          // weight rows this CTA stages per K step
  const int b_bytes = b_rows * kBlockK * 2;
  const int rank = kPair ? static_cast<int>(cluster_ctarank()) : 0;
  const int tile0 = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  uint8_t* sA = base + kCtrlBytes + epi_bytes;
  uint8_t* sB = sA + p.stages * kABytes;
  const uint32_t sA_addr = base_addr + kCtrlBytes + epi_bytes;
  const uint32_t sB_addr = sA_addr + p.stages * kABytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.mapA[s]);
    tma_prefetch_desc(&p.mapB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 2);    // the A and the B producer each arrive once (expect_tx) per K step
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kPair ? 16 : 8);   // 8 epilogue warps per CTA; the leader's copy collects both CTAs
    }
    for (int s = 0; s < 8 * kMaxResBufs; ++s) mbar_init(&res_full[s], 1);
    if (kStaged) {
      tma_prefetch_desc(&p.mapOut);
      if (p.residual != nullptr) tma_prefetch_desc(&p.mapRes);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair) tmem_alloc_pair<kTmemCols>(tmem_slot); else tmem_alloc<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if (kPair) cluster_sync_relaxed(); else __syncthreads();   // peer barriers must be initialised (fence.mbarrier_init) before remote arrives
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();     // PDL: everything above overlapped the previous kernel; its results are needed from here on
  if (kTraceBuild && p.trace != nullptr && threadIdx.x == 0) {
    p.trace[static_cast<size_t>(blockIdx.x) * kTraceStride + 0] = globaltimer_ns();
    p.trace[static_cast<size_t>(blockIdx.x) * kTraceStride + 2] = static_cast<unsigned long long>(clock64());
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ A (activation) producer
    // whole warp runs the warp-uniform loop (coordinates stay in uniform registers), one elected lane issues
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      const uint32_t full0 = kPair ? mapa_u32(smem_u32(&full_bar[0]), 0) : 0u;   // the leader's full barriers
      for (int wk = tile0; wk < p.num_tiles; wk += tile_step, ++iter) {
        const WorkItem wi = decode_work(p, wk);
        const TileCoord tc = decode_tile<kPair>(p, wi.tile, rank);
        if (issuer) trace_ev(p, iter, 0);
        int kstep = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const CUtensorMap* ma = &p.mapA[s];
          const int taps = p.seg_taps[s];
          const int kc = p.seg_kc[s];
          const int sw0 = tc.w0 * p.seg_stride[s], sh0 = tc.h0 * p.seg_stride[s];
          for (int t = 0; t < taps; ++t) {
            if (kstep >= wi.k1) break;                       // split-K: past this work item's K range
            if (kstep + kc <= wi.k0) {                       // split-K: whole tap before the range
              kstep += kc;
              continue;
            }
            // 3x3 window: (ky, kx) - pad;  2x2 window (one phase of a fused nearest-2x upsample conv): pad = 2*pad_y + pad_x
            const int dy = (taps == 9) ? (t / 3 - p.seg_pad[s]) : (taps == 4 ? ((t >> 1) - (p.seg_pad[s] >> 1)) : 0);
            const int dx = (taps == 9) ? (t % 3 - p.seg_pad[s]) : (taps == 4 ? ((t & 1) - (p.seg_pad[s] & 1)) : 0);
            for (int c = 0; c < kc; ++c, ++kstep) {
              if (kstep < wi.k0 || kstep >= wi.k1) continue;   // another split's K range
              mbar_wait_backoff(&empty_bar[stage], phase ^ 1u);
              if (issuer) {
                if (kPair) {
                  // both CTAs' boxes complete on the leader's barrier, which expects the bytes of the whole pair
                  if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * static_cast<uint32_t>(kABytes));
                  tma_load_4d_pair(ma, full0 + static_cast<uint32_t>(stage) * 8u, sA + stage * kABytes, c * kBlockK,
                                   sw0 + dx, sh0 + dy, tc.n0);
                } else {
                  mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(kABytes));
                  tma_load_4d(ma, &full_bar[stage], sA + stage * kABytes, c * kBlockK, sw0 + dx, sh0 + dy, tc.n0);
                }
              }
              __syncwarp();
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
        if (issuer) trace_ev(p, iter, 1);
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------------ B (weight) producer
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full0 = kPair ? mapa_u32(smem_u32(&full_bar[0]), 0) : 0u;
      for (int wk = tile0; wk < p.num_tiles; wk += tile_step) {
        const WorkItem wi = decode_work(p, wk);
        int n_row = (wi.tile - fast_div(wi.tile, p.mg_n, p.tiles_n) * p.tiles_n) * p.BN + rank * b_rows;
        if (p.b_img_rows > 0) n_row += decode_tile<kPair>(p, wi.tile, rank).n0 * p.b_img_rows;   // this tile's image
        for (int k = wi.k0; k < wi.k1; ++k) {
          mbar_wait_backoff(&empty_bar[stage], phase ^ 1u);
          if (issuer) {
            if (kPair) {
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * static_cast<uint32_t>(b_bytes));
              tma_load_2d_pair(&p.mapB, full0 + static_cast<uint32_t>(stage) * 8u, sB + stage * b_bytes, k * kBlockK, n_row);
            } else {
              mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(b_bytes));
              tma_load_2d(&p.mapB, &full_bar[stage], sB + stage * b_bytes, k * kBlockK, n_row);
            }
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the (warp-uniform) loop and waits; one elected lane issues.  Keeping the control flow
    // converged lets ptxas build the descriptors in uniform registers instead of moving them lane -> uniform per MMA.
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_f16(kPair ? 2 * kBlockM : kBlockM, static_cast<uint32_t>(p.BN), false, false);
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      int iter = 0;
      for (int wk = tile0; wk < p.num_tiles; wk += tile_step, ++iter) {
        const WorkItem wi = decode_work(p, wk);
        mbar_wait_backoff(&tmem_empty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kAccStride);
        for (int k = wi.k0; k < wi.k1; ++k) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (k == wi.k0 && issuer) trace_ev(p, iter, 2);
          const uint64_t da = umma_desc_kmajor_sw128(sA_addr + stage * kABytes);
          const uint64_t db = umma_desc_kmajor_sw128(sB_addr + stage * b_bytes);
          if (issuer) {
#pragma unroll
            for (int kk = 0; kk < kBlockK / 16; ++kk) {
              // advance 16 fp16 = 32 B along K inside the swizzle atom: +2 in the (addr >> 4) field
              if (kPair)
                umma_f16_ss_pair(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc,
                                 (k > wi.k0 || kk != 0) ? 1u : 0u);
              else
                umma_f16_ss(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc,
                            (k > wi.k0 || kk != 0) ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
            if (kPair) umma_commit_pair(&empty_bar[stage], 3); else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (issuer) {
          // accumulator complete -> epilogue warps (of both CTAs)
          if (kPair) umma_commit_pair(&tmem_full[as], 3); else umma_commit(&tmem_full[as]);
          trace_ev(p, iter, 3);
        }
        __syncwarp();
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else if (kStaged) {
    // ------------------------------------------------------------------ staged epilogue (warps 2..9, two groups)
    // Every warp drains its own 32 accumulator rows on its own: tcgen05.ld (software pipelined, the next chunk's load
    // flies while this one is finished) -> bias / activation / residual -> 64B-swizzled 2 KB slab in shared memory ->
    // its own TMA slab store.  No CTA- or group-wide barrier per chunk (only the per-tile bias staging syncs a group).
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are accessible to this warp
    const int row = quarter * 32 + lane;
    const int eg = (warp - 2) >> 2;              // epilogue group 0 / 1
    const int et = threadIdx.x - 64 - eg * 128;  // 0..127 within the group
    const bool issuer = elect_one();             // this warp's TMA thread (stores, residual prefetches)
    const bool tracer = issuer && quarter == 0;
    const bool active = eg < p.egroups;          // with one group, warps 6..9 only keep the TMEM hand-shake going
    const uint32_t bar_id = kEpiBarrier + eg;
    constexpr bool geglu = kGeglu;
    const int cols_per_tile = geglu ? p.BN / 2 : p.BN;     // logical output columns per tile
    const int nchunks = cols_per_tile / kChunkCols;
    const int n_logical = geglu ? p.N_out / 2 : p.N_out;
    const int bwh = p.bw * p.bh;
    const int img_in_tile = row / bwh;           // which of the tile's images this row belongs to
    const bool bias_staged_rb = !kGeneral || (p.rowbias == nullptr) || (p.bn <= kBiasImgs);
    const uint32_t swz = static_cast<uint32_t>((lane >> 1) & 3);
    // this warp's slab of the tile: rows [32*quarter, +32) = a {sbw, sbh, sbn} pixel box at (wq, hq, nq) inside the tile
    const int r0 = quarter * 32;
    const int wq = r0 % p.bw, hq = (r0 / p.bw) % p.bh, nq = r0 / bwh;
    uint8_t* sGroup = base + kCtrlBytes + (active ? eg : 0) * group_bytes;
    uint8_t* sOut = sGroup + quarter * (p.out_bufs + p.res_bufs) * kSlabBytes;     // [out_bufs][32 rows][64 B]
    uint8_t* sRes = sOut + p.out_bufs * kSlabBytes;                                // [res_bufs][32 rows][64 B]
    float* sBias = reinterpret_cast<float*>(sGroup + 4 * (p.out_bufs + p.res_bufs) * kSlabBytes);
    uint64_t* my_res_full = res_full + (warp - 2) * kMaxResBufs;
    uint8_t* my_out_row0 = sOut + lane * 64;
    const uint8_t* my_res_row0 = sRes + lane * 64;
    const int estep = p.egroups;                 // this group handles every estep-th chunk of the CTA's chunk stream
    // accumulator hand-back: the MMA thread (leader CTA) waits on ITS tmem_empty barriers
    uint32_t te_addr[2];
    te_addr[0] = kPair ? mapa_u32(smem_u32(&tmem_empty[0]), 0) : smem_u32(&tmem_empty[0]);
    te_addr[1] = kPair ? mapa_u32(smem_u32(&tmem_empty[1]), 0) : smem_u32(&tmem_empty[1]);
    auto release_acc = [&](int a) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(te_addr[a]); else mbar_arrive(&tmem_empty[a]);
      }
    };

    // residual prefetch cursor (issuer lane): runs res_bufs chunks of this warp ahead of the consumer, across tiles
    int pf_tile = tile0, pf_chunk = 0, pf_seq = 0;   // pf_seq: position in the CTA-wide chunk stream
    int pf_b = 0;                                    // residual slab the next prefetch lands in
    auto prefetch_residual = [&]() {
      while (pf_tile < p.num_tiles && (pf_seq & (estep - 1)) != eg) {   // skip the other group's chunks (estep is 1 or 2)
        ++pf_seq;
        if (++pf_chunk == nchunks) { pf_chunk = 0; pf_tile += tile_step; }
      }
      if (pf_tile >= p.num_tiles) return;
      const TileCoord t = decode_tile<kPair>(p, pf_tile, rank);
      const int col = t.n_blk * cols_per_tile + pf_chunk * kChunkCols;
      const int b = pf_b;
      if (++pf_b == p.res_bufs) pf_b = 0;
      mbar_expect_tx(&my_res_full[b], kSlabBytes);
      tma_load_4d(&p.mapRes, &my_res_full[b], sRes + b * kSlabBytes, col, t.w0 + wq, t.h0 + hq, t.n0 + nq);
      ++pf_seq;
      if (++pf_chunk == nchunks) { pf_chunk = 0; pf_tile += tile_step; }
    };
    if (has_res_stage && issuer && active) {
      for (int i = 0; i < p.res_bufs; ++i) prefetch_residual();
    }

    // per-tile bias rows (+ per-image bias): fetched one tile ahead into registers, staged in a double-buffered smem row
    float bv[2][kBiasImgs];
    const int nimg = (p.rowbias != nullptr && bias_staged_rb) ? min(p.bn, kBiasImgs) : 1;
    auto load_bias = [&](const TileCoord& t) {
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int cc = et + h2 * 128;              // column within the (packed) tile, < BN <= 256
        const int gcol = t.n_blk * p.BN + cc;
        const bool ok = (cc < p.BN) && (gcol < p.N_out);
        const float b0 = (ok && p.bias != nullptr) ? __ldg(p.bias + gcol) : 0.0f;
#pragma unroll
        for (int im = 0; im < kBiasImgs; ++im) {
          float v = b0;
          if (ok && im < nimg && p.rowbias != nullptr && bias_staged_rb) {
            const int img = min(t.n0 + im, p.NB - 1);
            v += __ldg(p.rowbias + static_cast<size_t>(img) * p.ld_rowbias + gcol);
          }
          bv[h2][im] = v;
        }
      }
    };
    if (active && tile0 < p.num_tiles) load_bias(decode_tile<kPair>(p, tile0, rank));

    int as = 0;
    uint32_t aphase = 0;
    int ob = 0;                    // output slab the next chunk of this warp is staged in
    int rb = 0;                    // residual slab / phase the next chunk of this warp reads
    uint32_t rphase = 0;
    int seq = 0;       // position in the CTA-wide chunk stream
    int iter = 0;
    const int esh = estep - 1;     // estep is 1 or 2: x % estep == x & esh, x / estep == x >> esh
    for (int tile = tile0; tile < p.num_tiles; tile += tile_step, ++iter) {
      const TileCoord tc = decode_tile<kPair>(p, tile, rank);
      const int col_tile = tc.n_blk * cols_per_tile;
      // does this group own any chunk of this tile?  (uniform over the group)
      const int first_c = (eg - seq) & esh;
      const bool has_work = active && first_c < nchunks;
      int last_c = -1;
      if (has_work) last_c = first_c + (((nchunks - 1 - first_c) >> esh) << esh);
      const float* my_bias = sBias;
      if (active) {
        float* sb = sBias + (iter & 1) * (kBiasImgs * 256);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int cc = et + h2 * 128;
          if (cc < p.BN) {
#pragma unroll
            for (int im = 0; im < kBiasImgs; ++im)
              if (im < nimg) sb[im * 256 + cc] = bv[h2][im];
          }
        }
        // one barrier per tile: this buffer was last read two tiles ago, and every thread passed the previous tile's
        // barrier only after it had finished those reads
        named_bar_sync(bar_id, 128);
        if (tile + tile_step < p.num_tiles) load_bias(decode_tile<kPair>(p, tile + tile_step, rank));
        my_bias = sb + ((p.rowbias != nullptr && bias_staged_rb) ? min(img_in_tile, kBiasImgs - 1) * 256 : 0);
      }
      const int my_img = min(tc.n0 + img_in_tile, p.NB - 1);

      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      if (tracer && active) trace_ev(p, iter, 4 + 2 * eg);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * kAccStride);
      if (!has_work) release_acc(as);              // nothing to read: release our share of the accumulator buffer
      uint32_t v[32];
      if (has_work && !geglu && !UDT_DBG(p, 8)) tmem_ld32(taddr + first_c * kChunkCols, v);   // prologue of the load pipeline
      for (int c = first_c; has_work && c < nchunks; c += estep) {
        const int b = ob;
        if (++ob == p.out_bufs) ob = 0;
        const int c0 = c * kChunkCols;
        float f[32];
        if constexpr (geglu) {
          uint32_t vg[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld32(taddr + p.BN / 2 + c0, vg);
          tmem_ld_wait();
          const float* bgate = my_bias + p.BN / 2;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {      // bias rows as 16-byte broadcast loads (the loop is issue-bound)
            const float4 bx = *reinterpret_cast<const float4*>(my_bias + c0 + j);
            const float4 bg = *reinterpret_cast<const float4*>(bgate + c0 + j);
            f[j] = (__uint_as_float(v[j]) + bx.x) * gelu_erf_fast(__uint_as_float(vg[j]) + bg.x);
            f[j + 1] = (__uint_as_float(v[j + 1]) + bx.y) * gelu_erf_fast(__uint_as_float(vg[j + 1]) + bg.y);
            f[j + 2] = (__uint_as_float(v[j + 2]) + bx.z) * gelu_erf_fast(__uint_as_float(vg[j + 2]) + bg.z);
            f[j + 3] = (__uint_as_float(v[j + 3]) + bx.w) * gelu_erf_fast(__uint_as_float(vg[j + 3]) + bg.w);
          }
        } else {
          if (!UDT_DBG(p, 8)) tmem_ld_wait_dep(v);   // this chunk's accumulators have landed in v[]
          if (tracer && eg == 0 && c == first_c) trace_ev(p, iter, 0, true);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(my_bias + c0 + j);
            f[j] = __uint_as_float(v[j]) + b4.x;
            f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
            f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
            f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
          }
          if (c + estep < nchunks && !UDT_DBG(p, 8)) tmem_ld32(taddr + (c + estep) * kChunkCols, v);   // next chunk's load in flight
          if constexpr (kGeneral) {
            if (!bias_staged_rb) {
              const float* rb = p.rowbias + static_cast<size_t>(my_img) * p.ld_rowbias + col_tile + c0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col_tile + c0 + j < p.N_out) f[j] += __ldg(rb + j);
            }
            if (p.act == UDT_ACT_SILU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
            } else if (p.act == UDT_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
            } else if (p.act == UDT_ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = gelu_erf_fast(f[j]);
            }
          }
        }
        if (geglu && c == last_c) release_acc(as);   // (pipelined path: released after the loop, no load in flight)
        if (has_res_stage) {
          const int rb_i = rb;
          mbar_wait(&my_res_full[rb_i], rphase);
          if (++rb == p.res_bufs) {
            rb = 0;
            rphase ^= 1u;
          }
          const uint8_t* rrow = my_res_row0 + rb_i * kSlabBytes;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 rv = *reinterpret_cast<const uint4*>(rrow + ((static_cast<uint32_t>(q) ^ swz) << 4));
            const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 rf = __half22float2(rh[j]);
              f[q * 8 + 2 * j] += rf.x;
              f[q * 8 + 2 * j + 1] += rf.y;
            }
          }
        }
        if (tracer && eg == 0 && c == first_c) trace_ev(p, iter, 1, true);
        // the slab store that last used sOut[b] (out_bufs chunks ago) must have finished reading it
        if (issuer && !UDT_DBG(p, 32)) {
          if (p.out_bufs == 4) tma_store_wait_read<3>();
          else if (p.out_bufs == 3) tma_store_wait_read<2>();
          else tma_store_wait_read<1>();
        }
        __syncwarp();
        if (tracer && eg == 0 && c == first_c) trace_ev(p, iter, 2, true);
        uint8_t* orow = my_out_row0 + b * kSlabBytes;
        if (!UDT_DBG(p, 4))
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 ov;
          ov.x = pack_half2(f[q * 8 + 0], f[q * 8 + 1]);
          ov.y = pack_half2(f[q * 8 + 2], f[q * 8 + 3]);
          ov.z = pack_half2(f[q * 8 + 4], f[q * 8 + 5]);
          ov.w = pack_half2(f[q * 8 + 6], f[q * 8 + 7]);
          *reinterpret_cast<uint4*>(orow + ((static_cast<uint32_t>(q) ^ swz) << 4)) = ov;
        }
        if (!UDT_DBG(p, 2)) fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the TMA engine
        __syncwarp();                              // slab complete in sOut[b]; the residual slab is consumed
        if (tracer && eg == 0 && c == first_c) trace_ev(p, iter, 3, true);
        if (issuer) {
          const int col = col_tile + c0;
          if (col < n_logical && !UDT_DBG(p, 1))
            tma_store_4d(&p.mapOut, sOut + b * kSlabBytes, col, tc.w0 + wq, tc.h0 + hq, tc.n0 + nq);
          if (!UDT_DBG(p, 64)) tma_store_commit();
          if (has_res_stage) prefetch_residual();  // refill the residual slab just consumed
        }
        if (tracer && eg == 0 && c == first_c) trace_ev(p, iter, 6, true);
      }
      if (has_work && !geglu) release_acc(as);     // every load of this accumulator has completed
      if (tracer && active) trace_ev(p, iter, 5 + 2 * eg);
      if (tracer && eg == 0) trace_ev(p, iter, 7, true);
      seq += nchunks;
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
    if (issuer && active) tma_store_wait_all<0>();   // all output slabs written before the CTA retires
  } else if (p.ksplit > 1) {
    // ------------------------------------------------------------------ split-K partial tiles (warps 2..9)
    // raw fp32 accumulators -> workspace.  tcgen05.ld hands every thread one ROW (32 columns = 128 B); written as is, one
    // store instruction would touch 32 different rows.  Each warp therefore transposes its 32x32 block through a 4 KB
    // XOR-swizzled shared-memory scratch so that eight lanes write one row's 128 bytes (4 full lines per instruction).
    const int quarter = warp & 3;
    const int eg = (warp - 2) >> 2;
    float* scratch = reinterpret_cast<float*>(base + kCtrlBytes + (warp - 2) * 4096);
    const int nch = p.BN / kChunkCols;
    const int sub = lane >> 3, cj = lane & 7;
    int as = 0;
    uint32_t aphase = 0;
    for (int wk = tile0; wk < p.num_tiles; wk += tile_step) {
      const WorkItem wi = decode_work(p, wk);
      const TileCoord tc = decode_tile<kPair>(p, wi.tile, rank);
      long long moff[8];
      bool ok[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {             // the 8 rows of this warp's block that this lane stores
        const int r = quarter * 32 + 4 * i + sub;
        const int w = tc.w0 + (r & (p.bw - 1));
        const int h = tc.h0 + ((r >> p.lbw) & (p.bh - 1));
        const int n = tc.n0 + (r >> (p.lbw + p.lbh));
        ok[i] = (w < p.W) && (h < p.H) && (n < p.NB);
        moff[i] = ((static_cast<long long>(n) * p.H + h) * p.W + w) * p.ldo;
      }
      float* outp = reinterpret_cast<float*>(p.out) + static_cast<long long>(wi.split) * p.split_stride;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * kAccStride);
      for (int c = eg; c < nch; c += 2) {
        uint32_t v[32];
        tmem_ld32(taddr + c * kChunkCols, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(scratch + lane * 32 + ((j ^ (lane & 7)) << 2)) =
              make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int col = tc.n_blk * p.BN + c * kChunkCols + cj * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = 4 * i + sub;
          const uint4 t = *reinterpret_cast<const uint4*>(scratch + rl * 32 + ((cj ^ (rl & 7)) << 2));
          if (ok[i] && col + 4 <= p.N_out) *reinterpret_cast<uint4*>(outp + moff[i] + col) = t;
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0)); else mbar_arrive(&tmem_empty[as]);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  } else if (warp >= 6) {
    // ------------------------------------------------------------------ direct mode: warps 6..9 only hand-shake
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = tile0; tile < p.num_tiles; tile += tile_step) {
      mbar_wait(&tmem_full[as], aphase);
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0)); else mbar_arrive(&tmem_empty[as]);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  } else {
    // ------------------------------------------------------------------ direct epilogue (warps 2..5)
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are accessible to this warp
    const int row = quarter * 32 + lane;
    const int bwh = p.bw * p.bh;
    const int r_w = row % p.bw;
    const int r_h = (row / p.bw) % p.bh;
    const int r_n = row / bwh;
    const bool geglu = (p.act == UDT_ACT_GEGLU);
    const int n_logical = geglu ? p.N_out / 2 : p.N_out;
    const bool fast = (p.ldo % 8 == 0) && (p.residual == nullptr || p.ldr % 8 == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
    int as = 0;
    uint32_t aphase = 0;
    for (int wk = tile0; wk < p.num_tiles; wk += tile_step) {
      const WorkItem wi = decode_work(p, wk);
      const TileCoord tc = decode_tile<kPair>(p, wi.tile, rank);
      const int w = tc.w0 + r_w, h = tc.h0 + r_h, n = tc.n0 + r_n;
      const bool valid = (w < p.W) && (h < p.H) && (n < p.NB);
      const size_t m = (static_cast<size_t>(n) * p.H + h) * p.W + w;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * kAccStride);
      if (geglu) {
        const int half_bn = p.BN / 2;
        for (int c0 = 0; c0 < half_bn; c0 += 32) {
          uint32_t vx[32], vg[32];
          tmem_ld32(taddr + c0, vx);
          tmem_ld32(taddr + half_bn + c0, vg);
          tmem_ld_wait();
          const int col = tc.n_blk * half_bn + c0;      // logical output column
          const int bcol = tc.n_blk * p.BN + c0;        // packed weight row of the x half
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(vx[j]);
            float g = __uint_as_float(vg[j]);
            if (p.bias != nullptr) {
              x += __ldg(p.bias + bcol + j);
              g += __ldg(p.bias + bcol + half_bn + j);
            }
            f[j] = x * gelu_erf_f(g);
          }
          if (valid && col < n_logical) {
            __half* o = reinterpret_cast<__half*>(p.out) + m * p.ldo + col;
            if (fast) {
              uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                uint4 ov;
                ov.x = pack_half2(f[v * 8 + 0], f[v * 8 + 1]);
                ov.y = pack_half2(f[v * 8 + 2], f[v * 8 + 3]);
                ov.z = pack_half2(f[v * 8 + 4], f[v * 8 + 5]);
                ov.w = pack_half2(f[v * 8 + 6], f[v * 8 + 7]);
                o4[v] = ov;
              }
            } else {
              for (int j = 0; j < 32; ++j) o[j] = __float2half_rn(f[j]);
            }
          }
        }
      } else if (p.BN >= 32) {
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          const int col = tc.n_blk * p.BN + c0;
          if (valid && col < p.N_out) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            const int cnt = min(32, p.N_out - col);
            finish_columns(p, f, cnt, m, n, col, fast, wi.split);
          }
        }
      } else {  // BN == 16
        uint32_t v[16];
        tmem_ld16(taddr, v);
        tmem_ld_wait();
        const int col = tc.n_blk * p.BN;
        if (valid && col < p.N_out) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = (j < 16) ? __uint_as_float(v[j & 15]) : 0.0f;
          const int cnt = min(16, p.N_out - col);
          finish_columns(p, f, cnt, m, n, col, false, wi.split);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0)); else mbar_arrive(&tmem_empty[as]);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  if (kPair) cluster_sync_relaxed(); else __syncthreads();   // the peer may still read our operands / signal our barriers
  if (kTraceBuild && p.trace != nullptr && threadIdx.x == 0) {
    p.trace[static_cast<size_t>(blockIdx.x) * kTraceStride + 1] = globaltimer_ns();
    p.trace[static_cast<size_t>(blockIdx.x) * kTraceStride + 3] = static_cast<unsigned long long>(clock64());
  }
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair<kTmemCols>(tmem_base); else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// split-K second kernel: out[m, n] = sum_s ws[s][m][n] + bias[n] + rowbias[img(m)][n] + residual[m][n], 8 columns per thread
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int S, long long split_stride, int M,
                                                            int N, const float* __restrict__ bias,
                                                            const float* __restrict__ rowbias, int ld_rowbias, int rows_per_img,
                                                            const __half* __restrict__ residual, int ldr, __half* __restrict__ out,
                                                            int ldo) {
  griddep_launch();
  griddep_wait();
  const int nv = N >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(M) * nv) return;
  const int m = static_cast<int>(idx / nv);
  const int col = static_cast<int>(idx - static_cast<long long>(m) * nv) * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  const float* src = ws + static_cast<long long>(m) * N + col;
  for (int s = 0; s < S; ++s) {       // fixed order: deterministic
    const float4 a0 = *reinterpret_cast<const float4*>(src + s * split_stride);
    const float4 a1 = *reinterpret_cast<const float4*>(src + s * split_stride + 4);
    acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
    acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
  }
  if (bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __ldg(bias + col + j);
  }
  if (rowbias != nullptr) {
    const float* rb = rowbias + static_cast<long long>(m / rows_per_img) * ld_rowbias + col;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __ldg(rb + j);
  }
  if (residual != nullptr) {
    const uint4 rv = *reinterpret_cast<const uint4*>(residual + static_cast<long long>(m) * ldr + col);
    const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 rf = __half22float2(rh[j]);
      acc[2 * j] += rf.x;
      acc[2 * j + 1] += rf.y;
    }
  }
  uint4 ov;
  ov.x = pack_half2(acc[0], acc[1]);
  ov.y = pack_half2(acc[2], acc[3]);
  ov.z = pack_half2(acc[4], acc[5]);
  ov.w = pack_half2(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(out + static_cast<long long>(m) * ldo + col) = ov;
}

unsigned long long* g_trace_buf = nullptr;
long long g_trace_bytes = 0;

// Column-tile choice by a per-tile cost model calibrated on B200 (scripts/igemm_trace.py, scripts/micro/tma_rate.cu):
// a K step (64 deep) costs max(MMA floor = 2*BN clk, operand ingest = (16 KB + B bytes) / ~73 B/clk); a CTA pair halves the
// B bytes per CTA.  Tiles are processed in waves over the SMs (pairs: over SM pairs).
int pick_bn(int N_out, int tiles_m, int ksteps, bool pair, int sms) {
  if (N_out <= 16) return 16;
  static const int cands[] = {256, 224, 192, 160, 128, 96, 64, 32};
  const int units = pair ? sms / 2 : sms;
  const int tm = pair ? (tiles_m + 1) / 2 : tiles_m;
  int best = 128;
  double best_cost = 1e30;
  for (int bn : cands) {
    const int tn = (N_out + bn - 1) / bn;
    const long tiles = static_cast<long>(tn) * tm;
    const long waves = (tiles + units - 1) / units;
    const double ingest = 224.0 + (pair ? 0.877 : 1.754) * bn;
    const double kstep = (2.0 * bn > ingest) ? 2.0 * bn : ingest;
    const double mainloop = ksteps * kstep;
    const double epi = 400.0 + 14.0 * bn;
    const double tile = (mainloop > epi ? mainloop : epi) + 300.0;
    const double cost = static_cast<double>(waves) * tile + epi;
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

template <bool kPair, int kMode>
int launch_igemm(const IGemmParams& p, int grid, int smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(udt_igemm_kernel<kPair, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return udt_host::fail(UDT_ERR_LAUNCH, "cudaFuncSetAttribute(igemm smem): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (kPair) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = 2;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  na += udt_host::pdl_attr(&attrs[na]);
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, udt_igemm_kernel<kPair, kMode>, p);
  if (e != cudaSuccess) return udt_host::fail(UDT_ERR_LAUNCH, "udt_igemm launch: %s", cudaGetErrorString(e));
  return UDT_OK;
}

}  // namespace

extern "C" int udt_geglu_tile(void) { return kGegluTile; }

extern "C" int udt_debug_set_trace(void* buf, int64_t nbytes) {
  if (!kTraceBuild) return 0;   // production build: tracing is compiled out (rebuild with UDT_TRACE=1)
  g_trace_buf = reinterpret_cast<unsigned long long*>(buf);
  g_trace_bytes = nbytes;
  return kTraceStride;
}

extern "C" int udt_igemm(const udt_igemm_desc* d, void* stream) {
  using namespace udt_host;
  int rc = require_sm100();
  if (rc != UDT_OK) return rc;
  if (d == nullptr) return fail(UDT_ERR_SHAPE, "udt_igemm: null descriptor");
  const int nsrc = d->nsrc, NB = d->NB, H = d->H, W = d->W, N_out = d->N_out, act = d->act;
  if (nsrc < 1 || nsrc > 3) return fail(UDT_ERR_SHAPE, "udt_igemm: nsrc=%d (1..3)", nsrc);
  if (NB < 1 || H < 1 || W < 1 || N_out < 1) return fail(UDT_ERR_SHAPE, "udt_igemm: bad shape NB=%d H=%d W=%d N=%d", NB, H, W, N_out);
  const int geglu_tile = (act == UDT_ACT_GEGLU) ? (d->bn_hint > 0 ? d->bn_hint : kGegluTile) : 0;
  if (act == UDT_ACT_GEGLU && ((geglu_tile != 128 && geglu_tile != 256) || N_out % geglu_tile != 0 || d->out_fp32 || d->residual || d->rowbias))
    return fail(UDT_ERR_SHAPE, "udt_igemm: GEGLU needs a 128/256 column interleave (bn_hint), N_out %% tile == 0, fp16 out, no residual/rowbias");
  if (d->out == nullptr || d->weight == nullptr) return fail(UDT_ERR_SHAPE, "udt_igemm: null out / weight");

  int max_stride = 1;
  for (int s = 0; s < nsrc; ++s) {
    const int st = d->src[s].stride > 0 ? d->src[s].stride : 1;
    if (st != 1 && st != 2) return fail(UDT_ERR_SHAPE, "udt_igemm: segment %d stride=%d (1 or 2)", s, st);
    if (st > max_stride) max_stride = st;
  }

  IGemmParams p;
  memset(&p, 0, sizeof(p));
  // spatial tile: bw*bh*bn == 128, fewest tiles, prefer wide boxes (TMA box extent bw*stride <= 256)
  {
    long best_tiles = -1;
    for (int bw = 128; bw >= 1; bw >>= 1) {
      for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
        const int bn = 128 / (bw * bh);
        if (bw > 2 * W && bw > 1) continue;
        if (bw * max_stride > 256 || bh * max_stride > 256) continue;
        const long t = static_cast<long>((W + bw - 1) / bw) * ((H + bh - 1) / bh) * ((NB + bn - 1) / bn);
        if (best_tiles < 0 || t < best_tiles) {
          best_tiles = t;
          p.bw = bw;
          p.bh = bh;
          p.bn = bn;
        }
      }
    }
  }
  p.lbw = 0;
  while ((1 << p.lbw) < p.bw) ++p.lbw;
  p.lbh = 0;
  while ((1 << p.lbh) < p.bh) ++p.lbh;
  p.W = W;
  p.H = H;
  p.NB = NB;
  p.tiles_w = (W + p.bw - 1) / p.bw;
  p.tiles_h = (H + p.bh - 1) / p.bh;
  p.tiles_nb = (NB + p.bn - 1) / p.bn;
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_nb;
  const int sms = num_sms();
  int ksteps_est = 0;
  for (int s = 0; s < nsrc; ++s) ksteps_est += d->src[s].taps * ((d->src[s].C + 63) / 64);
  // CTA pairs (tcgen05 cta_group::2) whenever there are at least two M tiles and a real column tile
  static const int pair_env = udt_host::tune_int("UDT_IGEMM_PAIR", 1);
  bool pair = pair_env != 0 && tiles_m >= 2 && N_out > 16 && sms >= 2;
  int BN = geglu_tile > 0 ? geglu_tile : (d->bn_hint > 0 ? d->bn_hint : pick_bn(N_out, tiles_m, ksteps_est, pair, sms));
  if (BN != 16 && (BN % 32 != 0 || BN < 32 || BN > 256)) return fail(UDT_ERR_SHAPE, "udt_igemm: BN=%d unsupported", BN);
  if (BN == 16) pair = false;
  // split-K: a small-M problem (few tiles, long K) streams its weights with every SM by splitting the K range; the
  // fp32 partial tiles go to the caller's workspace and a second kernel reduces them and applies the epilogue
  int ksplit = 1, kper = ksteps_est;
  static const int split_env = udt_host::tune_int("UDT_IGEMM_SPLITK", 1);
  if (split_env != 0 && d->bn_hint == 0 && act == UDT_ACT_NONE && !d->out_fp32 && d->workspace != nullptr && N_out % 8 == 0 &&
      d->out_stride_w == 0 && d->weight_img_rows == 0 &&
      N_out >= 64 && d->ldo % 8 == 0 && (d->residual == nullptr || d->ldr % 8 == 0) &&
      ((reinterpret_cast<uintptr_t>(d->out) | reinterpret_cast<uintptr_t>(d->residual)) & 15) == 0) {
    const int units = pair ? sms / 2 : sms;
    const int tm = pair ? (tiles_m + 1) / 2 : tiles_m;
    // widest column tile that wastes < 10 % of the MMA work
    int bn_s = 256;
    for (int cand : {256, 224, 192, 160, 128, 96, 64}) {
      const int tn = (N_out + cand - 1) / cand;
      if (tn * cand * 10 <= N_out * 11) { bn_s = cand; break; }
      bn_s = cand;
    }
    const int tiles_s = tm * ((N_out + bn_s - 1) / bn_s);
    int want = units / tiles_s;
    if (want > ksteps_est / 8) want = ksteps_est / 8;
    if (want >= 2 && ksteps_est >= 64) {   // short K: the second kernel costs more than the idle SMs
      kper = (ksteps_est + want - 1) / want;
      ksplit = (ksteps_est + kper - 1) / kper;
      const long long need = static_cast<long long>(ksplit) * NB * H * W * N_out * 4;
      if (ksplit >= 2 && need <= d->workspace_bytes) BN = bn_s; else ksplit = 1;
    }
  }
  p.BN = BN;
  p.N_out = N_out;
  p.tiles_n = (N_out + BN - 1) / BN;
  p.ksplit = ksplit;
  p.kper = ksplit > 1 ? kper : ksteps_est;
  p.num_tiles = (pair ? (tiles_m + 1) / 2 : tiles_m) * p.tiles_n * ksplit;
  const int b_rows = pair ? BN / 2 : BN;
  {
    auto magic = [](int dv) { return dv <= 1 ? 0u : static_cast<uint32_t>(((1ull << 32) + dv - 1) / dv); };
    p.mg_n = magic(p.tiles_n);
    p.mg_w = magic(p.tiles_w);
    p.mg_h = magic(p.tiles_h);
    p.mg_s = magic(p.ksplit);
    const unsigned long long worst = static_cast<unsigned long long>(2 * p.num_tiles + 2) *
        static_cast<unsigned long long>(p.tiles_n > p.tiles_w ? (p.tiles_n > p.tiles_h ? p.tiles_n : p.tiles_h)
                                                              : (p.tiles_w > p.tiles_h ? p.tiles_w : p.tiles_h));
    if (worst >= (1ull << 32)) return fail(UDT_ERR_SHAPE, "udt_igemm: problem too large for the tile decoder");
  }

  int ktotal = 0;
  p.nseg = nsrc;
  for (int s = 0; s < nsrc; ++s) {
    const udt_gemm_src& a = d->src[s];
    const int st = a.stride > 0 ? a.stride : 1;
    const int Hs = a.H > 0 ? a.H : H * st, Ws = a.W > 0 ? a.W : W * st;
    if (a.C < 8 || a.C % 8 != 0) return fail(UDT_ERR_SHAPE, "udt_igemm: segment %d has C=%d (multiple of 8 required)", s, a.C);
    if (a.taps != 1 && a.taps != 9 && a.taps != 4) return fail(UDT_ERR_SHAPE, "udt_igemm: segment %d taps=%d (1, 4 or 9)", s, a.taps);
    if (a.taps == 4 && (a.pad < 0 || a.pad > 3 || st != 1)) return fail(UDT_ERR_SHAPE, "udt_igemm: segment %d: 2x2 window needs pad in 0..3 (2*pad_y + pad_x), stride 1", s);
    if (a.ld < a.C) return fail(UDT_ERR_SHAPE, "udt_igemm: segment %d ld=%d < C=%d", s, a.ld, a.C);
    if (a.taps == 9 && a.pad != 0 && a.pad != 1) return fail(UDT_ERR_SHAPE, "udt_igemm: segment %d pad=%d (0 or 1)", s, a.pad);
    if (a.taps == 1 && st != 1) return fail(UDT_ERR_SHAPE, "udt_igemm: segment %d: point-wise segments must have stride 1", s);
    rc = make_tmap_nhwc(&p.mapA[s], a.ptr, a.C, Ws, Hs, NB, a.ld, p.bw, p.bh, p.bn, st);
    if (rc != UDT_OK) return rc;
    p.seg_kc[s] = (a.C + 63) / 64;  // channels are zero-filled by TMA up to the next multiple of 64
    p.seg_taps[s] = a.taps;
    p.seg_stride[s] = st;
    p.seg_pad[s] = a.pad;
    ktotal += a.taps * p.seg_kc[s] * 64;
  }
  p.ksteps = ktotal / 64;
  // a plain GEMM (one point-wise segment) may pass an unpadded [N_out, C] weight: TMA zero-fills the K tail
  const int kdim_b = (nsrc == 1 && d->src[0].taps == 1) ? d->src[0].C : ktotal;
  const int ldw = d->ldw > 0 ? d->ldw : kdim_b;
  if (ldw < kdim_b) return fail(UDT_ERR_SHAPE, "udt_igemm: ldw=%d < K=%d (conv weights must be packed per 64-channel block)", ldw, kdim_b);
  p.b_img_rows = 0;
  uint64_t w_rows = static_cast<uint64_t>(N_out);
  if (d->weight_img_rows > 0) {
    // per-image weights (folded cross-attention): a tile must lie inside one image, and so must a CTA pair
    if (p.bn != 1 || d->weight_img_rows < N_out || (pair && ((p.tiles_w * p.tiles_h) & 1)))
      return fail(UDT_ERR_SHAPE, "udt_igemm: per-image weights need >= 128 (pairs: a multiple of 256) pixels per image");
    p.b_img_rows = d->weight_img_rows;
    w_rows = static_cast<uint64_t>(d->weight_img_rows) * NB;
  }
  rc = make_tmap_2d(&p.mapB, d->weight, static_cast<uint64_t>(kdim_b), w_rows,
                    static_cast<uint64_t>(ldw), 64, static_cast<uint32_t>(b_rows));
  if (rc != UDT_OK) return rc;

  // staged (TMA store) epilogue for fp16 outputs with a TMA-compatible layout; tiny / fp32 outputs store directly
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(d->out) & 15) == 0) && (d->ldo % 8 == 0) &&
                         (d->residual == nullptr || (((reinterpret_cast<uintptr_t>(d->residual) & 15) == 0) && d->ldr % 8 == 0));
  const int n_logical = act == UDT_ACT_GEGLU ? N_out / 2 : N_out;
  p.staged = (ksplit == 1 && !d->out_fp32 && BN >= 32 && aligned16 && n_logical >= 8) ? 1 : 0;
  if (d->out_stride_w > 0 && (!p.staged || d->residual != nullptr || d->out_stride_w % 8 || d->out_stride_h % 8 || d->out_stride_n % 8))
    return fail(UDT_ERR_SHAPE, "udt_igemm: a strided output view needs the TMA-store epilogue (fp16, >= 8 columns), no residual, strides %% 8 == 0");
  int epi_bytes = 0;
  if (p.staged) {
    // every epilogue warp stores its own 32-row slab of the tile: a {sbw, sbh, sbn} pixel box
    const int sbw = p.bw < 32 ? p.bw : 32;
    const int sbh = p.bh < 32 / sbw ? p.bh : 32 / sbw;
    const int sbn = 32 / (sbw * sbh);
    if (d->out_stride_w > 0)   // strided output view (one phase of the 2x-upsampled output)
      rc = make_tmap_nhwc_c32_strided(&p.mapOut, d->out, static_cast<uint64_t>(n_logical), W, H, NB, d->out_stride_w,
                                      d->out_stride_h, d->out_stride_n, sbw, sbh, sbn);
    else
      rc = make_tmap_nhwc_c32(&p.mapOut, d->out, static_cast<uint64_t>(n_logical), W, H, NB, d->ldo, sbw, sbh, sbn);
    if (rc != UDT_OK) return rc;
    // short-K GEMMs are epilogue-bound: two epilogue warpgroups and deeper chunk buffering; long-K tiles hide a
    // single group's epilogue behind the mainloop and keep the shared memory for operand stages instead
    const bool short_k = (ktotal / 64) <= 24;
    p.egroups = short_k ? 2 : 1;
    p.out_bufs = short_k ? 3 : 2;
    p.res_bufs = 0;
    if (d->residual != nullptr) {
      rc = make_tmap_nhwc_c32(&p.mapRes, d->residual, static_cast<uint64_t>(n_logical), W, H, NB, d->ldr, sbw, sbh, sbn);
      if (rc != UDT_OK) return rc;
      p.res_bufs = short_k ? 3 : 2;
    }
    epi_bytes = p.egroups * ((p.out_bufs + p.res_bufs) * kChunkBytes + kBiasBytes);
    const int stage_b = kABytes + b_rows * kBlockK * 2;
    if (short_k && (kSmemBudget - kCtrlBytes - 1024 - epi_bytes) / stage_b < 3) {
      // wide column tiles: keep at least 3 operand stages, fall back to the shallow single-group epilogue
      p.egroups = 1;
      p.out_bufs = 2;
      p.res_bufs = d->residual != nullptr ? 2 : 0;
      epi_bytes = (p.out_bufs + p.res_bufs) * kChunkBytes + kBiasBytes;
    }
  }
  if (ksplit > 1) epi_bytes = 8 * 4096;   // per-warp transpose scratch of the split-K partial-tile stores
  const int stage_bytes = kABytes + b_rows * kBlockK * 2;
  int stages = (kSmemBudget - kCtrlBytes - 1024 - epi_bytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return fail(UDT_ERR_SHAPE, "udt_igemm: tile does not fit shared memory");
  p.stages = stages;
  p.bias = d->bias;
  p.rowbias = d->rowbias;
  p.ld_rowbias = d->ld_rowbias;
  p.residual = reinterpret_cast<const __half*>(d->residual);
  p.ldr = d->ldr;
  p.out = d->out;
  p.ldo = d->ldo;
  p.out_fp32 = d->out_fp32;
  p.act = act;
  const long long M_rows = static_cast<long long>(NB) * H * W;
  if (ksplit > 1) {   // the GEMM kernel writes raw fp32 partial tiles; bias / per-image bias / residual move to the reduce kernel
    p.out = d->workspace;
    p.ldo = N_out;
    p.out_fp32 = 1;
    p.split_stride = M_rows * N_out;
    p.bias = nullptr;
    p.rowbias = nullptr;
    p.residual = nullptr;
  }

  static const int dbg = udt_host::tune_int("UDT_IGEMM_DEBUG", 0);
  p.debug = dbg;
  p.trace = nullptr;
  if (dbg & 0xF00) {
    const int force = (dbg >> 8) & 0xF;
    if (force >= 2 && force <= stages) stages = force, p.stages = force;
  }
  const int smem = kCtrlBytes + 1024 + epi_bytes + stages * stage_bytes;
  const int units = pair ? sms / 2 : sms;
  const int grid = (p.num_tiles < units ? p.num_tiles : units) * (pair ? 2 : 1);
  if (g_trace_buf != nullptr && static_cast<long long>(grid) * kTraceStride * 8 <= g_trace_bytes) p.trace = g_trace_buf;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int mode = 3;
  if (p.staged) {
    if (act == UDT_ACT_GEGLU) mode = 1;
    else if (act != UDT_ACT_NONE || (p.rowbias != nullptr && p.bn > kBiasImgs)) mode = 2;
    else mode = 0;
  }
  switch (mode * 2 + (pair ? 1 : 0)) {
    case 0: rc = launch_igemm<false, 0>(p, grid, smem, st); break;
    case 1: rc = launch_igemm<true, 0>(p, grid, smem, st); break;
    case 2: rc = launch_igemm<false, 1>(p, grid, smem, st); break;
    case 3: rc = launch_igemm<true, 1>(p, grid, smem, st); break;
    case 4: rc = launch_igemm<false, 2>(p, grid, smem, st); break;
    case 5: rc = launch_igemm<true, 2>(p, grid, smem, st); break;
    case 6: rc = launch_igemm<false, 3>(p, grid, smem, st); break;
    default: rc = launch_igemm<true, 3>(p, grid, smem, st); break;
  }
  if (rc != UDT_OK || ksplit == 1) return rc;
  const long long work = M_rows * (N_out / 8);
  launch_pdl(splitk_reduce_kernel, dim3(static_cast<unsigned>((work + 255) / 256)), dim3(256), 0, st,
             reinterpret_cast<const float*>(d->workspace), ksplit, p.split_stride, static_cast<int>(M_rows), N_out, d->bias,
             d->rowbias, d->ld_rowbias, H * W, reinterpret_cast<const __half*>(d->residual), d->ldr,
             reinterpret_cast<__half*>(d->out), d->ldo);
  return check_launch("udt_igemm (split-K reduce)");
}
