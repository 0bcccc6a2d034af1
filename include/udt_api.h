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
#define UDT_ACT_RELU 3
#define UDT_ACT_GELU 4 /* exact (erf) GELU: PARSeq's ViT MLP and decoder FFN (strhub/models/parseq/modules.py:33-50) */

int udt_version(void);            /* ABI version (5: udt_cfg_euler_step takes cfg_scale_dev; fp32-stream LabelEncoder entry points) */
int udt_arch(void);               /* compute capability of the current device *10 (100 on B200); <0 on error */
const char* udt_last_error(void); /* thread-local message of the last failing call */
int udt_num_sms(void);
int udt_sizeof_igemm_desc(void);  /* sizeof(udt_igemm_desc) as compiled: lets a foreign-language binding check its struct mirror */

/* One K-segment of the implicit GEMM: an NHWC fp16 activation tensor read either point-wise (taps = 1:
 * nn.Linear / 1x1 conv) or through a 3x3 window (taps = 9) with zero padding, stride 1 or 2. */
typedef struct {
  const void* ptr; /* fp16 [NB, H, W, ld] */
  int32_t C;       /* channels consumed from this tensor (multiple of 8; each tap is zero-filled up to a multiple
                      of 64 by TMA, and the packed weight carries the same per-tap padding) */
  int32_t ld;      /* channel pitch in elements (>= C, multiple of 8) */
  int32_t taps;    /* 1 (point-wise), 9 (3x3 window) or 4 (2x2 window: one phase of a fused nearest-2x upsample + 3x3 conv) */
  int32_t H, W;    /* spatial size of this tensor; 0 = output size * stride */
  int32_t stride;  /* 1 or 2 (0 = 1): output pixel (y, x) reads input pixel (y*stride + ky - pad, ...) */
  int32_t pad;     /* low-side zero padding of the 3x3 window: 1 = symmetric pad 1 (openaimodel.py:132-139),
                      0 = the VAE's pad-(0,1,0,1) downsample (model.py:77-85); 2x2 window: 2*pad_y + pad_x (each 0 or 1) */
} udt_gemm_src;

typedef struct {
  udt_gemm_src src[3];
  int32_t nsrc;          /* 1..3 K-segments */
  int32_t NB, H, W;      /* output pixels: M = NB*H*W rows (a plain GEMM is NB = H = 1, W = M) */
  const void* weight;    /* fp16 [N_out, ldw], K ordered (segment, tap = ky*3+kx, channel padded to 64) */
  int32_t ldw;           /* weight row pitch in elements (0 = K_total) */
  int32_t N_out;
  const float* bias;     /* fp32 [N_out] or NULL */
  const float* rowbias;  /* fp32 [NB, ld_rowbias] (one row per image; ld_rowbias = 0 broadcasts one row) or NULL */
  int32_t ld_rowbias;
  const void* residual;  /* fp16 [M, ldr] or NULL; may alias out */
  int32_t ldr;
  void* out;             /* fp16 (or fp32) [M, ldo] */
  int32_t ldo;
  int32_t out_fp32;
  int32_t act;           /* UDT_ACT_* */
  int32_t bn_hint;       /* column tile, 0 = library picks */
  int64_t out_stride_w, out_stride_h, out_stride_n; /* 0 = dense NHWC output; otherwise the output pixel (w, h, n)
                            lives at out + w*stride_w + h*stride_h + n*stride_n elements (a strided view, e.g. one of the four
                            phases of a 2x-upsampled tensor); fp16 TMA-store epilogue only, no residual */
  int32_t weight_img_rows; /* 0 = one weight [N_out, ldw] for all images; > 0 = per-image weights: image i uses rows
                            [i * weight_img_rows, i * weight_img_rows + N_out) of `weight` (>= 128 pixels per image) */
  void* workspace;       /* optional scratch (device, 16-byte aligned) for split-K partial tiles; NULL disables split-K */
  int64_t workspace_bytes;
} udt_igemm_desc;

/* K2/K3 — segmented implicit GEMM on tcgen05/TMEM fed by TMA:
 *     out[m, n] = act( sum_seg sum_tap sum_c A_seg[pixel(m)*stride+tap-pad, c] * Wt[n, k(seg,tap,c)]
 *                      + bias[n] + rowbias[image(m), n] ) + residual[m, n]
 * Replaces: nn.Conv2d 3x3 stride 1/2 (openaimodel.py:186,223-229,85-87,132-139; model.py:108-117,77-85), 1x1
 * skip/nin_shortcut (openaimodel.py:240; model.py:124-126) fused as an extra K segment, nn.Linear
 * (attention.py:47,66,127-135,193-199,375,395; openaimodel.py:212-215,341-343), the `h + emb_out` add
 * (openaimodel.py:266) as `rowbias`, residual adds (openaimodel.py:268; attention.py:315-341,416) and GEGLU
 * (attention.py:49-51).  For UDT_ACT_GEGLU the logical output has N_out/2 columns (see udt_geglu_tile()).
 * nearest-2x upsample + 3x3 conv (openaimodel.py:99-102; model.py:55-68) runs as four 2x2-window launches, one per output
 * phase, with the 3x3 taps that fall on the same source pixel pre-summed in the weights (2.25x fewer FLOPs). */
int udt_igemm(const udt_igemm_desc* desc, void* stream);
/* column tile (BN) the GEGLU weight interleave must be packed for (x half then gate half per tile) */
int udt_geglu_tile(void);
/* tuning aid (scripts/igemm_trace.py): while `buf` (device memory, nbytes) is set, every udt_igemm launch whose
 * grid fits writes per-CTA role timestamps into it; buf = NULL switches tracing off.  Returns the number of
 * 8-byte slots per CTA. */
int udt_debug_set_trace(void* buf, int64_t nbytes);

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

/* Request front-end — the batch construction of demo.py:52-62,78-80 and the image export of
 * demo.py:100-101 / test.py:94 on the device, so PCIe carries uint8:
 * udt_request_pack_u8: image_hwc uint8 [Bs, H, W, 3], mask_hwc uint8 [Bs, H, W, MC] (0 = keep; MC <= 4) ->
 *   image fp32 [B, 3, H, W] = u8 / 127.5 - 1, m = mean_c(mask == 0), masked = image * m, mask fp32 [B, 1, H, W] = 1 - m;
 *   sample b reads source b % Bs (Bs = 1: one request tiled num_samples times).  Bit-exact vs the torch expressions.
 * udt_images_to_u8: x fp32 [NB, C, HW] (NCHW, in [0, 1]) -> y uint8 [NB, HW, C] = trunc(x * 255). */
int udt_request_pack_u8(const uint8_t* image_hwc, const uint8_t* mask_hwc, float* image, float* mask, float* masked,
                        int32_t B, int32_t Bs, int32_t H, int32_t W, int32_t MC, void* stream);
int udt_images_to_u8(const float* x, uint8_t* y, int32_t NB, int32_t C, int32_t HW, void* stream);

/* K12 — noise-search score of one t_attn layer (loss.py:192-235 FullLoss.get_min_local_loss, read by
 * sampling.py:340 get_init_noise):
 *   score[b] += -min_{l < seg_l} ( max_n mask_s[b % Bm, n] * blur(mean_h probs[b*heads + h, n, l]) + 1 - seg[b % Bm, l] )
 * probs fp32 [B*heads, N = size*size, L] (the exported attn_map), mask fp32 [Bm, H, W] sampled like
 * F.interpolate(mask, (size, size)) (nearest), seg fp32 [Bm, seg_l], gk fp32 [ks*ks] Gaussian (zero padding ks/2,
 * ks odd <= 7), score fp32 [B] accumulated over successive launches (zero it first; divide by the layer count).
 * B = UNet batch ([uc; c] when CFG-doubled: B = 2*Bm). */
int udt_attn_local_score(const float* probs, const float* mask, const float* seg, const float* gk, float* score,
                         int32_t B, int32_t Bm, int32_t heads, int32_t size, int32_t L, int32_t seg_l, int32_t H, int32_t W,
                         int32_t ks, void* stream);

/* LabelEncoder (encoders/modules.py:1088-1173): character embedding + sinusoid positional encoding,
 * out fp16 [rows = B*L, D] = emb[idx[row], :] + pe[row % L, :] (emb fp32 [95, D], pe fp32 [L, D], idx int32), and
 * the multi-head self-attention of its nn.TransformerEncoder layers over the fused in_proj output
 * qkv fp16 [B*L, ld >= 3*heads*dh] (q | k | v), L <= 16 tokens, head dim <= 256, no mask:
 * o fp16 [B*L, ldo] = softmax(q k^T * scale) v per head.  The projections / FFN run through udt_igemm. */
int udt_label_embed(const int32_t* idx, const float* emb, const float* pe, void* out, float* out_f32 /* NULL or fp32 [rows, D] */,
                    void* out_lo /* NULL or fp16 [rows, D]: rn(v - rn(v)), the low half of the fp16 pair */, int32_t rows,
                    int32_t L, int32_t D, void* stream);
/* fp32-stream evaluation of the LabelEncoder: its output conditions every sampler step of a request, so its rounding error
 * is systematic (it does not average out over the steps like the UNet's activation rounding).  The encoder is tiny
 * (7 GFLOP per string), so its residual stream stays fp32 and every GEMM operand is an fp16 PAIR hi + lo
 * (x ~= hi + lo to 2^-22): x W^T ~= hi Wh^T + lo Wh^T + hi Wl^T on the same tensor-core kernel (udt_igemm, fp32 out).
 *   udt_rowsum_norm_split: y[r, :] = LN( relu?( in0 + in1 + in2 ) + res )  (in1 / in2 / res / the LN (gamma = beta = NULL) are
 *     optional) -> out_f32 and / or the pair out_hi / out_lo; the post-LN `x = norm(x + sublayer(x))` of
 *     nn.TransformerEncoderLayer (encoders/modules.py:1103-1104) and the partial-product sums of the pair GEMMs.
 *   udt_mha_small_f32: udt_mha_small on fp32 q | k | v, output as the pair o_hi / o_lo. */
int udt_rowsum_norm_split(const float* in0, const float* in1, const float* in2, const float* res, int32_t rows, int32_t C,
                          const float* gamma, const float* beta, float eps, int32_t relu, float* out_f32, void* out_hi,
                          void* out_lo, void* stream);
int udt_mha_small_f32(const float* qkv, void* o_hi, void* o_lo, int32_t B, int32_t L, int32_t heads, int32_t dh, int32_t ld,
                      int32_t ldo, float scale, void* stream);
int udt_mha_small(const void* qkv, void* o, int32_t B, int32_t L, int32_t heads, int32_t dh, int32_t ld, int32_t ldo,
                  float scale, void* stream);

/* Multi-head attention over short sequences with masks — the self- / cross-attention of PARSeq's two-stream decoder layer
 * (src/parseq/strhub/models/parseq/modules.py:57-75, nn.MultiheadAttention batch_first; OCR scoring of test.py:58-91):
 * o[b, i, h*dh:(h+1)*dh] = softmax_j( q_i . k_j * scale + mask[i, j] ; -inf where key_padding_mask[b, j] ) v_j.
 * q fp16 [B*Lq, ldq], k / v fp16 [B*Lk, ldk / ldv] (head h at columns [h*dh, h*dh + dh)), o fp16 [B*Lq, ldo];
 * mask fp32 [Lq, ldm] additive (may hold -inf) or NULL; key_padding_mask uint8 [B, Lk] (non-zero = ignore) or NULL;
 * Lk <= 160, dh <= 64.  A fully masked row yields zeros. */
int udt_mha_masked(const void* q, const void* k, const void* v, void* o, int32_t B, int32_t Lq, int32_t Lk, int32_t heads,
                   int32_t dh, int32_t ldq, int32_t ldk, int32_t ldv, int32_t ldo, float scale, const float* mask, int32_t ldm,
                   const uint8_t* key_padding_mask, void* stream);

/* Folded textual cross-attention (attention.py:140-174 with the step-invariant K / V of the 12 context tokens folded into
 * the projections; exact in real arithmetic):
 *   scores[m, h*L + l] = LN(t)[m, :] . W1[s(m)][h*L + l, :],   W1[s][h*L + l, c] = scale * sum_d K[s, l, h*64 + d] * Wq[h*64 + d, c]
 *   out[m, :]          = P[m, :] . W2[s(m)]^T + bias,           W2[s][c, h*L + l]  = sum_d Wo[c, h*64 + d] * V[s, l, h*64 + d]
 * udt_xattn_fold builds W1 [B, Npad, C] and W2 [B, C, Npad] (fp16, Npad = heads*L rounded up to 64, pad rows / columns 0)
 * from kc / vc fp16 [B, L, ldkv] (head h at columns [h*64, h*64+64)), wq / wo fp16 [C, ldwq] / [C, ldwo] (C = heads*64);
 * the two GEMMs run through udt_igemm with per-image weights; udt_softmax_groups turns the scores into probabilities:
 * in / out fp16 [rows, ld] (may alias), `groups` groups of L consecutive columns per row, columns >= groups*L of `out` are
 * zeroed up to `cols`; optional fp32 export probs[(img*groups + g), n, l] with rows = imgs * N (the reference's attn_map). */
int udt_xattn_fold(const void* kc, const void* vc, int32_t ldkv, const void* wq, int32_t ldwq, const void* wo, int32_t ldwo,
                   void* w1, void* w2, int32_t B, int32_t L, int32_t heads, int32_t Npad, float scale, void* stream);
int udt_softmax_groups(const void* in, void* out, int32_t rows, int32_t cols, int32_t ld, int32_t groups, int32_t L,
                       float* probs, int32_t N, void* stream);

/* row-wise softmax over fp16 [rows, cols] in place with a pre-scale (VAE single-head attention,
 * model.py:246-248, executed as GEMM -> softmax -> GEMM). */
int udt_softmax_rows(void* x, int32_t rows, int32_t cols, int32_t ld, float scale, void* stream);

/* K7 — sampler glue (guiders.py:25-40, denoiser.py:22-28, wrappers.py:27, sampling_utils.py:39-40,
 * sampling.py:85-86,349-351).  The sampler state keeps the reference's layout: x fp32 NCHW [B,4,HW], concat
 * fp32 NCHW [B,5,HW]; the UNet side is NHWC.  The per-step scalars live in device memory (`c_in_dev`,
 * `dsigma_dev` point at one fp32 each) so that one captured CUDA graph serves every step.
 *  pack:  unet_in[2B,HW,16] fp16 <- cat(x * c_in, concat_{uc|c}) zero padded to 16 channels (uc half first)
 *  step:  eps = eps_u + scale*(eps_c - eps_u); x += (sigma_next - sigma) * eps;  eps2b fp32 NHWC [2B,HW,4]
 *         (`cfg_scale_dev` != NULL: the guidance scale is read from device memory too, so a captured graph serves every
 *          request whatever scale its guider carries)   */
int udt_cfg_pack(const float* x, const float* concat_uc, const float* concat_c, void* unet_in, int32_t B, int32_t HW,
                 const float* c_in_dev, void* stream);
int udt_cfg_euler_step(float* x, const float* eps2b, int32_t B, int32_t HW, float cfg_scale, const float* dsigma_dev,
                       const float* cfg_scale_dev /* NULL, or one fp32 on the device that overrides cfg_scale */, void* stream);

/* K10 — conditioner tail (encoders/modules.py:843-857,1011-1014,195-198; distributions.py:24-41):
 * concat_{c,uc} fp32 NCHW [B,5,h*w] = cat(bilinear_1/8(mask), scale_factor * (mean + exp(0.5*clamp(logvar,-30,20)) * noise_{c,uc}))
 * moments fp32 NHWC [B,h*w,ld_moments] (mean = channels 0..3, logvar = 4..7); noise fp32 NCHW [B,4,h*w];
 * mask fp32 [B,1,8h,8w]. */
int udt_vae_sample_pack(const float* moments, int32_t ld_moments, const float* noise_c, const float* noise_uc,
                        const float* mask, float* concat_c, float* concat_uc, int32_t B, int32_t h, int32_t w,
                        float scale_factor, void* stream);

/* per-pixel affine map on <= 8 channels: out fp16 NHWC [B,HW,Cpad] = Wm[Cout,Cin] * (x[B,Cin,HW] * in_scale) + bias,
 * channels >= Cout zero.  post_quant_conv with the 1/scale_factor of decode_first_stage folded in
 * (autoencoder.py:313-316; diffusion.py:124-129). */
int udt_pointwise_affine(const float* x, const float* Wm, const float* bias, void* out, int32_t B, int32_t HW,
                         int32_t Cin, int32_t Cout, int32_t Cpad, float in_scale, void* stream);

/* data movement helpers */
/* nearest-neighbour 2x upsample of NHWC fp16 (openaimodel.py:99; model.py:65) */
int udt_upsample2x_nhwc(const void* x, void* y, int32_t NB, int32_t H, int32_t W, int32_t C, void* stream);
/* layout / dtype conversion at the API boundary: NCHW fp32 <-> NHWC fp16 (channel padded with zeros) */
int udt_nchw_f32_to_nhwc_f16(const float* x, void* y, int32_t NB, int32_t C, int32_t HW, int32_t Cpad, void* stream);
int udt_nhwc_to_nchw_f32(const void* x, int32_t x_is_fp32, float* y, int32_t NB, int32_t C, int32_t HW, int32_t ld,
                         float scale, float shift, int32_t clamp01, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UDT_API_H */
