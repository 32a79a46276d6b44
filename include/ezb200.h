/* libezb200 -- C ABI of the B200-native EzAudio hot path (DiT denoiser step + Oobleck VAE decode).
 *
 * The reference (haidog-yaqub/EzAudio) is pure Python/PyTorch and has no FFI of its own; its drop-in boundary is the
 * Python call surface.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference root).  Python binds these with ctypes (ezaudio_b200/_lib.py); INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes, no torch types.  Every device pointer is owned by the caller (PyTorch) and
 * must stay valid until the stream-ordered call has executed.  All work is enqueued on the cudaStream_t passed as
 * `stream` (void*).  The per-step entry points (set_context, set_timesteps, forward, controlnet_forward, cfg_ddim_step, decode,
 * encode, energy_condition, t5_forward) never synchronise and are safe inside CUDA-graph capture; the load-time ones (create,
 * load_weight, finalize_weights) may synchronise the device.  A handle is not re-entrant.  Returns 0 on success,
 * a negative ezb_status otherwise; ezb_last_error() gives the message of the calling thread's last failure.
 */
#ifndef EZB200_H
#define EZB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  EZB_OK = 0, EZB_ERR_ARG = -1, EZB_ERR_SHAPE = -2, EZB_ERR_UNSUPPORTED = -3, EZB_ERR_CUDA = -4, EZB_ERR_STATE = -5,
  EZB_ERR_WEIGHT = -6
} ezb_status;

typedef struct ezb_dit ezb_dit; /* one MaskDiT/UDiT or DiTControlNet instance on one device */
typedef struct ezb_vae ezb_vae; /* one OobleckDecoder instance on one device */

/* Hyper-parameters of `MaskDiT(**params['model'])` (api/ezaudio.py:83; ckpts/ezaudio-xl.yml:5-37).  Only the shipped
 * switch combination is implemented (1d, ada_sola_bias, cross, rope shared, qk layernorm, geglu, skip+skip_norm). */
typedef struct {
  int32_t embed_dim, num_heads, depth, context_dim, inner_dim, ada_rank;
  float ada_scaling;          /* ada_sola_alpha / ada_sola_rank (src/models/blocks.py:25) */
  int32_t latent_chans;       /* 128: x / gt channels; in_chans = 2*latent_chans + 1 */
  int32_t is_controlnet;      /* 1: DiTControlNet (src/models/controlnet.py:87) -- first half + stem + zero linears */
  int32_t cond_c0, cond_c1;   /* controlnet stem widths (cond_blocks, ckpts/controlnet/energy_l.yml:40) */
  int32_t max_batch, max_len, max_ctx_len, max_timesteps; /* workspace bounds (effective batch incl. CFG doubling) */
  int32_t precision;          /* 0: bf16 operands / fp32 accumulate; 1: bf16x3 split operands (fp32-grade parity mode) */
} ezb_dit_desc;

int ezb_version(void);
const char* ezb_last_error(void);

/* --- model lifetime / weights: replaces MaskDiT(...).load_state_dict(torch.load(ckpt)['model']) (api/ezaudio.py:83-85) */
int ezb_dit_create(ezb_dit** out, const ezb_dit_desc* desc, int device);
int ezb_dit_destroy(ezb_dit* h);
/* One call per state-dict entry, reference key names and layouts; `data` is a DEVICE fp32 pointer.
 * The library repacks into its own layouts (QKV concat, GEGLU interleave, bf16 / split-bf16 cast) and keeps no
 * reference to `data`. */
int ezb_dit_load_weight(ezb_dit* h, const char* ref_key, const float* data, const int64_t* shape, int ndim, void* stream);
int ezb_dit_finalize_weights(ezb_dit* h, void* stream); /* fails listing the first missing key */

/* --- step-invariant precompute.
 * context path: udit.py:94-97,295 (context_embed) + blocks.py:150 (norm_context) + attention.py:128-129,142 (to_k,to_v,norm_k)
 * for every block; ctx (Be,Lc,context_dim) fp32, ctx_mask (Be,Lc) uint8 (1 = keep; attention.py:30-37). */
int ezb_dit_set_context(ezb_dit* h, const float* ctx, const uint8_t* ctx_mask, int Be, int Lc, void* stream);
/* time path: modules.py:19-61 (TimestepEmbedder), udit.py:313-316 (time_act, time_ada, time_ada_final), blocks.py:39-45
 * (AdaLN) evaluated for n distinct timestep values (HOST array); forward calls then refer to them by index. */
int ezb_dit_set_timesteps(ezb_dit* h, const int64_t* timesteps_host, int n, void* stream);

/* --- one denoiser forward: MaskDiT.forward (conditioners.py:156-183) + UDiT.forward (udit.py:281-362).
 * x (Be,C,L) fp32; gt (Be,C,L) fp32 or NULL (mask_embed everywhere, conditioners.py:174-175); gt_mask (Be,L) uint8 or
 * NULL: 1 = position is regenerated (gt replaced by mask_embed there, mask channel = 1; conditioners.py:150-153,176);
 * t_index_host[Be]: index into the table of ezb_dit_set_timesteps per sample (NULL = all use `t_index_all`);
 * controlnet_skips: NULL or depth/2 device pointers (Be,L,D) fp32 in in-block order (udit.py:345-348);
 * out (Be,C,L) fp32. */
int ezb_dit_forward(ezb_dit* h, const float* x, const float* gt, const uint8_t* gt_mask, const int32_t* t_index_host,
                    int t_index_all, const float* const* controlnet_skips, float* out, int Be, int L, void* stream);
/* DiTControlNet.forward (controlnet.py:252-315): condition (Be,1,2L) fp32; writes depth/2 skips (Be,L,D) fp32, already
 * multiplied by conditioning_scale, into skips_out[i]. */
int ezb_controlnet_forward(ezb_dit* h, const float* x, const float* gt, const uint8_t* gt_mask, const int32_t* t_index_host,
                           int t_index_all, const float* condition, float conditioning_scale, float* const* skips_out, int Be,
                           int L, void* stream);

/* --- fused classifier-free guidance + rescale + DDIM update (src/inference.py:12-23,88-100; diffusers DDIMScheduler.step
 * restated).  model_out holds B text rows followed by B uncond rows when guidance_scale != 0, else B
 * rows.  coef = {sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma}; noise (B,C,L) may be NULL when
 * sigma == 0.  latents updated in place.  `device`: the CUDA device the pointers live on (the call makes it current). */
int ezb_cfg_ddim_step(int device, const float* model_out, float* latents, const float* noise, int B, int C, int L, float guidance_scale,
                      float guidance_rescale, const float* coef5_host, void* stream);

/* --- VAE decoder: OobleckDecoder.forward (stable_vae/models/autoencoders.py:149-190) behind
 * Autoencoder(embedding=z) (src/modules/autoencoder_wrapper.py:74-77). */
typedef struct {
  int32_t latent_dim, channels, out_channels;
  int32_t n_stages;
  int32_t c_mults[8];  /* config c_mults (without the leading 1) */
  int32_t strides[8];
  int32_t max_batch, max_latent_len;
  int32_t precision;
  int32_t with_encoder;    /* 1: also hold OobleckEncoder + VAE bottleneck (editing_audio, api/ezaudio.py:175) */
  int32_t in_channels;     /* 1 */
  int32_t enc_latent_dim;  /* 2 * latent_dim (mean | scale) */
} ezb_vae_desc;
int ezb_vae_create(ezb_vae** out, const ezb_vae_desc* desc, int device);
int ezb_vae_destroy(ezb_vae* h);
/* keys: "decoder.layers...." with weight_g / weight_v / bias / alpha / beta (stable_vae/__init__.py:25-31 after prefix strip) */
int ezb_vae_load_weight(ezb_vae* h, const char* ref_key, const float* data, const int64_t* shape, int ndim, void* stream);
int ezb_vae_finalize_weights(ezb_vae* h, void* stream);
int ezb_vae_decode(ezb_vae* h, const float* z /*(B,latent,L)*/, float* wav /*(B,out_channels,L*prod(strides))*/, int B, int L,
                   void* stream);
/* Autoencoder(audio=x) (src/modules/autoencoder_wrapper.py:69-73): OobleckEncoder (stable_vae/models/autoencoders.py:115-146) +
 * VAEBottleneck.encode (bottleneck.py:66-87): z = mean + (softplus(scale) + 1e-4) * noise.  audio (B,1,T) fp32, T a multiple of the
 * hop (480); noise (B,latent,L) fp32 drawn by the caller (the reference uses torch.randn_like), NULL -> z = mean. */
int ezb_vae_encode(ezb_vae* h, const float* audio, const float* noise, float* z, int B, int T, void* stream);

/* EnergyExtractor.forward (src/models/conditions/energy.py:19-56) as wrapped by Conditioner (condition_wrapper.py:26-42): audio (B,T) fp32
 * -> (B, T/hop) fp32 frame energies in dB, normalised per clip when norm != 0; quantize_levels <= 1 disables quantisation. Only the shipped
 * padding mode ('reflect') exists. */
int ezb_energy_condition(int device, const float* audio, float* out, int B, int T, int hop_size, int window_size, float min_db, int norm,
                         int quantize_levels, void* stream);

/* --- waveform pre / post-processing around the path, device pointers, fp32 mono clips.
 * ezb_wave_prepare: per clip  x <- x / (max|x| + 1e-9) when normalize != 0 (api/ezaudio.py:147, api/controlnet.py:119), samples with
 *   |x| <= gate zeroed when gate > 0 (`surpass_noise`, api/controlnet.py:121-124), then zero-padded or cropped from T_in to T_out samples
 *   (api/controlnet.py:131-136).  in (B,T_in), out (B,T_out).
 * ezb_wave_splice: dst[start : start+n] = src[0 : n] -- the paste of the regenerated chunk into the original clip (api/ezaudio.py:198-203).
 * ezb_wave_to_pcm16: round(x * 32768) saturated to int16 -- the sample format soundfile.write(path, audio, sr) produces by default for
 *   WAV (t2a_demo.py:13,20). */
int ezb_wave_prepare(int device, const float* in, float* out, int B, int T_in, int T_out, int normalize, float gate, void* stream);
int ezb_wave_splice(int device, float* dst, long long dst_len, const float* src, long long start, long long n, void* stream);
int ezb_wave_to_pcm16(int device, const float* in, int16_t* out, long long n, void* stream);

/* --- T5 (v1.1 / flan-T5, gated-GELU) text encoder: `text_encoder(input_ids=, attention_mask=).last_hidden_state`, src/inference.py:38-50;
 * model class transformers.T5EncoderModel loaded at api/ezaudio.py:78-79 (the step before the denoiser path). */
typedef struct ezb_t5 ezb_t5;
typedef struct {
  int32_t vocab_size, d_model, d_kv, num_heads, d_ff, num_layers;
  int32_t num_buckets;   /* relative_attention_num_buckets (32) */
  int32_t max_distance;  /* relative_attention_max_distance (128) */
  float eps;             /* layer_norm_epsilon (1e-6) */
  int32_t max_batch, max_len;
  int32_t precision;     /* 0 = bf16 operands, 1 = bf16x3 (parity mode), as in ezb_dit_desc */
} ezb_t5_desc;
int ezb_t5_create(ezb_t5** out, const ezb_t5_desc* desc, int device);
int ezb_t5_destroy(ezb_t5* h);
/* keys of T5EncoderModel.state_dict(): shared.weight, encoder.block.{i}.layer.0.SelfAttention.{q,k,v,o}.weight, ...relative_attention_bias.weight
 * (block 0), encoder.block.{i}.layer.{0,1}.layer_norm.weight, ...DenseReluDense.{wi_0,wi_1,wo}.weight, encoder.final_layer_norm.weight */
int ezb_t5_load_weight(ezb_t5* h, const char* ref_key, const float* data, const int64_t* shape, int ndim, void* stream);
int ezb_t5_finalize_weights(ezb_t5* h, void* stream);
/* ids (B, L) int32, attention mask (B, L) uint8 (1 = token), out (B, L, d_model) fp32, all device pointers.  `buckets` (L, L) int32 device
 * pointer = T5Attention._relative_position_bucket(key - query) as computed by the caller with the reference's own torch ops, or NULL to let
 * the library compute it on the host in float32. */
int ezb_t5_forward(ezb_t5* h, const int32_t* ids, const uint8_t* mask, const int32_t* buckets, float* out, int B, int L, void* stream);

/* --- kernel-level hooks used by tests/ and profiling only (not part of the drop-in surface). */
typedef struct {
  const float* bias; int32_t bias_mod;
  const float* resid; int32_t ldr;
  const float* gate; int32_t gate_bstride; int32_t rows_per_batch;
  float* out_f32; int32_t ld32;
  void* out_bf16; int32_t ld16; int32_t split_stride;
  int32_t act; const float* act_a; const float* act_b;
} ezb_test_epilogue;
/* C = A[M,K] W[N,K]^T through the tcgen05 GEMM; epi_kind 0 = linear epilogue, 1 = GEGLU (packed W). conv_* = 0 for plain. */
int ezb_test_gemm(int device, const void* A_bf16, int lda, const void* W_bf16, int ldw, int M, int N, int K, int bn, int epi_kind,
                  const ezb_test_epilogue* e, int conv_taps, int conv_center, int conv_dil, int conv_cin_pad, int conv_T,
                  int conv_B, void* stream);
/* impl 0: fp32 CUDA-core kernel (q, k, v fp32 [B,H,L,dh]); 1: the tcgen05 kernel the product uses (generation 6 unless the option "attn6" says
   otherwise); 4 / 6: generation 4 / 6 forced; +100: q / k rows of 80 elements for dh = 72 (the product's layout) instead of 128 */
int ezb_test_attention(int device, const void* q, const void* k, const void* vt, const uint8_t* key_mask, void* out_bf16,
                       int B, int H, int Lq, int Lk, int dh, int impl, void* stream);

/* runtime switches for A/B measurements and profiling (csrc/host.cuh, csrc/ezb.cu list them with their measured verdicts): e.g. "pair_gemm" (1 =
   cta_group::2 256-row tiles, default), "attn6" (attention kernel generation / mode, default 5), "ksub2", "ln_variant", "skip" (profiling: kernel
   classes not launched).  Defaults are the measured-best settings; products never need to call this. */
int ezb_set_option(const char* name, int value);
/* incremented by every ezb_set_option call: hosts that cache captured CUDA graphs key them on it (options change kernel selection) */
unsigned long long ezb_option_epoch(void);
/* debugging aid: with option "gemm_debug"=1, CTA 0 of each pair-GEMM accumulates cycle counters; this reads and resets them */
int ezb_debug_read(unsigned long long* out8);
/* accounting: kernels launched by this library so far (process-wide); per-GEMM CUDA-event timing for bench.py's roofline leg */
unsigned long long ezb_launch_count(void);
void ezb_launch_count_add(unsigned long long n); /* launches replayed from a captured CUDA graph */
int ezb_prof_gemm_begin(void);
int ezb_prof_gemm_end(int* launches, double* flops, double* ms);
int ezb_prof_gemm_stats(double min_flops, int* launches, double* flops, double* ms); /* subset of the last profile */

#ifdef __cplusplus
}
#endif
#endif
