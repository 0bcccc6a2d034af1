/* udt_api.h — C-ABI of libudt_b200.so: the sm_100a kernels behind UDiffText's inference hot path.
 *
 * The reference (ZYM-PKU/UDiffText) is pure Python/PyTorch and has no FFI of its own; every entry point
 * below replaces one library-kernel call site of the reference (cited as file:line relative to the
 * reference root).  Conventions:
 *   - plain C, no torch types; device pointers are raw `void*` (16-byte aligned), fp16 activations are
 *     NHWC / row-major `[rows, channels]`, fp32 where stated;
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), never allocates or frees
 *     device memory, never synchronises, and is CUDA-graph capturable;
 *   - return value 0 = ok, negative = UDT_ERR_*; `udt_last_error()` describes the last failure of the
 *     calling thread.  There is no CPU / non-sm_100 fallback: on any other device the calls fail.
 */
#ifndef UDT_API_H
#define UDT_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UDT_OK 0
#define UDT_ERR_SHAPE (-1)
#define UDT_ERR_ALIGN (-2)
#define UDT_ERR_ARCH (-3)
#define UDT_ERR_LAUNCH (-4)
#define UDT_ERR_DRIVER (-5)

/* epilogue activation of udt_igemm */
#define UDT_ACT_NONE 0
#define UDT_ACT_SILU 1
#define UDT_ACT_GEGLU 2 /* out[:, j] = x_j * gelu_erf(gate_j); weight rows interleaved per column tile */

int udt_version(void);            /* ABI version (1) */
int udt_arch(void);               /* compute capability of the current device *10 (100 on B200); <0 on error */
const char* udt_last_error(void); /* thread-local message of the last failing call */
int udt_num_sms(void);

/* One K-segment of the implicit GEMM: an NHWC fp16 activation tensor read either point-wise (taps = 1:
 * nn.Linear / 1x1 conv) or through a 3x3, stride-1, zero-pad-1 window (taps = 9). */
typedef struct {
  const void* ptr; /* fp16 [NB, H, W, ld] */
  int32_t C;       /* channels consumed from this tensor (multiple of 64) */
  int32_t ld;      /* channel pitch in elements (>= C, multiple of 8) */
  int32_t taps;    /* 1 or 9 */
} udt_gemm_src;

/* K2/K3 — segmented implicit GEMM on tcgen05/TMEM fed by TMA:
 *     out[m, n] = act( sum_seg sum_tap sum_c A_seg[pixel(m)+tap, c] * Wt[n, k(seg,tap,c)]
 *                      + bias[n] + rowbias[image(m), n] ) + residual[m, n]
 * Replaces: nn.Conv2d 3x3 (openaimodel.py:186,223-229,85-87; model.py:108-117), 1x1 skip/nin_shortcut
 * (openaimodel.py:240; model.py:124-126) fused as an extra K segment, nn.Linear (attention.py:47,66,
 * 127-135,193-199,375,395; openaimodel.py:212-215,341-343), the `h + emb_out` add (openaimodel.py:266)
 * as `rowbias`, residual adds (openaimodel.py:268; attention.py:315-341,416) and GEGLU (attention.py:49-51).
 * `weight` is fp16 [N_out, K_total], K ordered (segment, tap = ky*3+kx, channel).  For UDT_ACT_GEGLU the
 * logical output has N_out/2 columns (see udt_geglu_tile()).  `residual` may alias `out`.
 * `rowbias` is fp32 [NB, ld_rowbias] (one row per image).  `bn_hint` = 0 lets the library pick the column tile. */
int udt_igemm(const udt_gemm_src* srcs, int32_t nsrc, int32_t NB, int32_t H, int32_t W, const void* weight,
              int32_t N_out, const float* bias, const float* rowbias, int32_t ld_rowbias, const void* residual,
              int32_t ldr, void* out, int32_t ldo, int32_t out_fp32, int32_t act, int32_t bn_hint, void* stream);
/* column tile (BN) the GEGLU weight interleave must be packed for (x half then gate half per tile) */
int udt_geglu_tile(void);

/* K1 — GroupNorm(32 groups)(+SiLU) over NHWC fp16, fp32 statistics; optionally normalises the channel
 * concatenation of two tensors (`x1` may be NULL) and writes one [NB, HW, C0+C1] tensor.
 * Replaces GroupNorm32+SiLU (diffusionmodules/util.py:273-275; openaimodel.py:185,220,538), Normalize
 * (attention.py:82-85; model.py:49-52) and the th.cat of skip connections (openaimodel.py:620).
 * Deterministic (no atomics): per-CTA fp64 partial sums are written to `stats_ws`, a caller-owned 8-byte
 * aligned workspace of udt_groupnorm_ws_bytes(NB, HW, C0+C1, groups) bytes, and folded in a fixed order. */
int64_t udt_groupnorm_ws_bytes(int32_t NB, int32_t HW, int32_t C, int32_t groups);
int udt_groupnorm_nhwc(const void* x0, int32_t C0, const void* x1, int32_t C1, void* y, int32_t NB, int32_t HW,
                       int32_t groups, const float* gamma, const float* beta, float eps, int32_t silu,
                       void* stats_ws, void* stream);

/* K6 — LayerNorm over the last dim of fp16 [rows, C] (attention.py:297,310-311). */
int udt_layernorm(const void* x, void* y, int32_t rows, int32_t C, const float* gamma, const float* beta, float eps,
                  void* stream);

/* K4 — softmax(Q K^T * scale) V, head dim 64, fp16 in/out, tcgen05 S/O tiles in TMEM.
 * q/k/v/o are row-major [B*N, ld*] with head h at columns [h*64, h*64+64) of each pointer.
 * Replaces xformers.ops.memory_efficient_attention (attention.py:246-248). */
int udt_fmha_fwd(const void* q, const void* k, const void* v, void* o, int32_t B, int32_t Nq, int32_t Nkv,
                 int32_t heads, int32_t ldq, int32_t ldk, int32_t ldv, int32_t ldo, float scale, void* stream);

/* K5 — textual cross-attention with a short context (L <= 16 tokens), head dim 64:
 * probs = softmax_L(q k^T * scale) (sigmoid if L == 1), o = probs v; optionally exports probs as fp32
 * [B*heads, N, L] (the reference's attn_map_cache, attention.py:147-174). kc/vc: fp16 [B, L, ldkv]. */
int udt_xattn_small_l(const void* q, const void* kc, const void* vc, void* o, float* probs, int32_t B, int32_t N,
                      int32_t L, int32_t heads, int32_t ldq, int32_t ldkv, int32_t ldo, float scale, void* stream);

/* row-wise softmax over fp16 [rows, cols] in place with a pre-scale (VAE single-head attention,
 * model.py:246-248, executed as GEMM -> softmax -> GEMM). */
int udt_softmax_rows(void* x, int32_t rows, int32_t cols, int32_t ld, float scale, void* stream);

/* K7 — sampler glue (guiders.py:25-40, denoiser.py:22-28, wrappers.py:27, sampling_utils.py:39-40,
 * sampling.py:85-86,349-351), all on NHWC: x fp32 [B,HW,4].
 *  pack:  unet_in[2B,HW,16] fp16 <- cat(x * c_in, concat_{uc|c}) zero padded to 16 channels
 *  step:  eps = eps_u + scale*(eps_c - eps_u); x += (sigma_next - sigma) * eps            */
int udt_cfg_pack(const float* x, const float* concat_uc, const float* concat_c, void* unet_in, int32_t B, int32_t HW,
                 float c_in, void* stream);
int udt_cfg_euler_step(float* x, const float* eps2b, int32_t B, int32_t HW, float cfg_scale, float dsigma,
                       void* stream);

/* data movement helpers */
/* nearest-neighbour 2x upsample of NHWC fp16 (openaimodel.py:99; model.py:65) */
int udt_upsample2x_nhwc(const void* x, void* y, int32_t NB, int32_t H, int32_t W, int32_t C, void* stream);
/* explicit im2col for the rare convs the TMA path does not cover (C_in not a multiple of 64, stride 2,
 * asymmetric VAE padding model.py:77-85): out fp16 [NB*Ho*Wo, Kpad], K order (tap, channel), zero padded. */
int udt_im2col3x3_nhwc(const void* x, void* out, int32_t NB, int32_t H, int32_t W, int32_t C, int32_t ld,
                       int32_t stride, int32_t pad_lo, int32_t Ho, int32_t Wo, int32_t Kpad, void* stream);
/* layout / dtype conversion at the API boundary: NCHW fp32 <-> NHWC fp16 (channel padded with zeros) */
int udt_nchw_f32_to_nhwc_f16(const float* x, void* y, int32_t NB, int32_t C, int32_t HW, int32_t Cpad, void* stream);
int udt_nhwc_to_nchw_f32(const void* x, int32_t x_is_fp32, float* y, int32_t NB, int32_t C, int32_t HW, int32_t ld,
                         float scale, float shift, int32_t clamp01, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UDT_API_H */
