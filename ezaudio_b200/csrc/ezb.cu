// C-ABI entry points of libezb200.so (declared in include/ezb200.h).
#include "../../include/ezb200.h"

#include <algorithm>
#include "dit.cuh"
#include "vae.cuh"
#include "t5.cuh"
#include "host.cuh"

using namespace ezb;

namespace {
Device& device_ctx(int device) {
  static Device devs[16];
  Device& d = devs[device & 15];
  if (d.num_sms == 148 && d.id == 0 && device >= 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) d.num_sms = sms;
    d.id = device;
  }
  return d;
}
EpiLinearParams to_epi(const ezb_test_epilogue* e) {
  EpiLinearParams p;
  memset(&p, 0, sizeof p);
  p.bias = e->bias; p.bias_mod = e->bias_mod; p.resid = e->resid; p.ldr = e->ldr; p.gate = e->gate;
  p.gate_bstride = e->gate_bstride; p.rows_per_batch = e->rows_per_batch; p.out_f32 = e->out_f32; p.ld32 = e->ld32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(e->out_bf16); p.ld16 = e->ld16; p.split_stride = e->split_stride;
  p.act = e->act; p.act_a = e->act_a; p.act_b = e->act_b; p.out_scale = 0.f; p.phase_cols = 0; p.phase_ld16 = 0;
  return p;
}
}  // namespace

extern "C" {

__attribute__((visibility("default"))) int ezb_version(void) { return 1; }
__attribute__((visibility("default"))) const char* ezb_last_error(void) { return last_error().c_str(); }

__attribute__((visibility("default"))) int ezb_test_gemm(int device, const void* A, int lda, const void* W, int ldw, int M, int N, int K, int bn, int epi_kind,
                  const ezb_test_epilogue* e, int conv_taps, int conv_center, int conv_dil, int conv_cin_pad, int conv_T, int conv_B,
                  void* stream) {
  if (!A || !W || !e) return fail(EZB_ERR_ARG, "ezb_test_gemm: null pointer");
  EZB_CUDA(cudaSetDevice(device));
  Device& dev = device_ctx(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(A);
  const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(W);
  ConvAddr conv;
  conv.taps = conv_taps; conv.center = conv_center; conv.dilation = conv_dil; conv.cin_pad = conv_cin_pad; conv.T = conv_T; conv.B = conv_B;
  const ConvAddr* cp = conv_taps > 0 ? &conv : nullptr;
  if (epi_kind == 20) {  // swap-AB: 128 features x 256 tokens tiles, fp32 output
    if (cp) return fail(EZB_ERR_UNSUPPORTED, "swap-AB GEMM has no conv addressing");
    EpiLinearParams p = to_epi(e);
    if (opt_swap_mc()) return gemm_swapped_mc<EpiLinearT<256>, 3>(dev, st, a, lda, w, ldw, M, N, K, p);
    return gemm_swapped<EpiLinearT<256>>(dev, st, a, lda, w, ldw, M, N, K, p);
  }
  if (epi_kind == 10 || epi_kind == 11) {  // CTA-pair kernel (bn is the pair tile's N)
    if (cp) return fail(EZB_ERR_UNSUPPORTED, "pair GEMM has no conv addressing");
    if (epi_kind == 10) {
      EpiLinearParams p = to_epi(e);
      if (bn == 128) return gemm2<128, EpiLinear<128>>(dev, st, a, lda, w, ldw, M, N, K, p);
      if (bn == 256) return gemm2<256, EpiLinear<256>>(dev, st, a, lda, w, ldw, M, N, K, p);
    } else {
      EpiGegluParams p;
      memset(&p, 0, sizeof p);
      p.bias = e->bias; p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(e->out_bf16); p.ld16 = e->ld16; p.split_stride = e->split_stride;
      if (bn == 256) return gemm2<256, EpiGeglu<256>>(dev, st, a, lda, w, ldw, M, N, K, p);
    }
    return fail(EZB_ERR_UNSUPPORTED, "ezb_test_gemm pair: bn=%d", bn);
  }
  if (epi_kind == 0) {
    EpiLinearParams p = to_epi(e);
    if (bn == 64) return gemm<64, EpiLinear<64>>(dev, st, a, lda, w, ldw, M, N, K, p, cp);
    if (bn == 128) return gemm<128, EpiLinear<128>>(dev, st, a, lda, w, ldw, M, N, K, p, cp);
    if (bn == 256) return gemm<256, EpiLinear<256>>(dev, st, a, lda, w, ldw, M, N, K, p, cp);
  } else if (epi_kind == 1) {
    EpiGegluParams p;
    memset(&p, 0, sizeof p);
    p.bias = e->bias; p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(e->out_bf16); p.ld16 = e->ld16; p.split_stride = e->split_stride;
    if (bn == 128) return gemm<128, EpiGeglu<128>>(dev, st, a, lda, w, ldw, M, N, K, p, cp);
    if (bn == 256) return gemm<256, EpiGeglu<256>>(dev, st, a, lda, w, ldw, M, N, K, p, cp);
  }
  return fail(EZB_ERR_UNSUPPORTED, "ezb_test_gemm: bn=%d epi=%d", bn, epi_kind);
}


#define EZB_API __attribute__((visibility("default")))
#define ST(s) reinterpret_cast<cudaStream_t>(s)

EZB_API int ezb_dit_create(ezb_dit** out, const ezb_dit_desc* desc, int device) {
  if (!out || !desc) return fail(EZB_ERR_ARG, "ezb_dit_create: null argument");
  EZB_CUDA(cudaSetDevice(device));
  Dit* h = new Dit();
  int rc = h->init(*desc, &device_ctx(device));
  if (rc != 0) { delete h; return rc; }
  *out = reinterpret_cast<ezb_dit*>(h);
  return EZB_OK;
}
EZB_API int ezb_dit_destroy(ezb_dit* h) {
  delete reinterpret_cast<Dit*>(h);
  return EZB_OK;
}
EZB_API int ezb_dit_load_weight(ezb_dit* h, const char* key, const float* data, const int64_t* shape, int ndim, void* stream) {
  if (!h || !key || !data || !shape) return fail(EZB_ERR_ARG, "ezb_dit_load_weight: null argument");
  return reinterpret_cast<Dit*>(h)->load_weight(key, data, shape, ndim, ST(stream));
}
EZB_API int ezb_dit_finalize_weights(ezb_dit* h, void* stream) {
  if (!h) return fail(EZB_ERR_ARG, "null handle");
  (void)stream;
  return reinterpret_cast<Dit*>(h)->finalize();
}
EZB_API int ezb_dit_set_context(ezb_dit* h, const float* ctx, const uint8_t* ctx_mask, int Be, int Lc, void* stream) {
  if (!h || !ctx || !ctx_mask) return fail(EZB_ERR_ARG, "ezb_dit_set_context: null argument");
  return reinterpret_cast<Dit*>(h)->set_context(ctx, ctx_mask, Be, Lc, ST(stream));
}
EZB_API int ezb_dit_set_timesteps(ezb_dit* h, const int64_t* ts, int n, void* stream) {
  if (!h || !ts) return fail(EZB_ERR_ARG, "ezb_dit_set_timesteps: null argument");
  return reinterpret_cast<Dit*>(h)->set_timesteps(ts, n, ST(stream));
}
EZB_API int ezb_dit_forward(ezb_dit* h, const float* x, const float* gt, const uint8_t* gt_mask, const int32_t* tidx, int tall,
                            const float* const* cskips, float* out, int Be, int L, void* stream) {
  if (!h || !x || !out) return fail(EZB_ERR_ARG, "ezb_dit_forward: null argument");
  return reinterpret_cast<Dit*>(h)->forward(x, gt, gt_mask, tidx, tall, cskips, out, Be, L, ST(stream));
}
EZB_API int ezb_controlnet_forward(ezb_dit* h, const float* x, const float* gt, const uint8_t* gt_mask, const int32_t* tidx, int tall,
                                   const float* condition, float scale, float* const* skips_out, int Be, int L, void* stream) {
  if (!h || !x || !condition || !skips_out) return fail(EZB_ERR_ARG, "ezb_controlnet_forward: null argument");
  return reinterpret_cast<Dit*>(h)->controlnet_forward(x, gt, gt_mask, tidx, tall, condition, scale, skips_out, Be, L, ST(stream));
}
EZB_API int ezb_cfg_ddim_step(int device, const float* model_out, float* latents, const float* noise, int B, int C, int L, float gs, float gr,
                              const float* coef, void* stream) {
  if (!model_out || !latents || !coef || B < 1 || C < 1 || L < 1) return fail(EZB_ERR_ARG, "ezb_cfg_ddim_step: bad argument");
  if (coef[4] != 0.f && !noise) return fail(EZB_ERR_ARG, "ezb_cfg_ddim_step: sigma != 0 needs a noise tensor");
  EZB_CUDA(cudaSetDevice(device));
  const int n = C * L;
  const float* uncond = gs != 0.f ? model_out + (size_t)B * n : nullptr;
  return launch_k(cfg_ddim_kernel, dim3(B * CFG_CLUSTER), dim3(1024), 0, ST(stream), CFG_CLUSTER, model_out, uncond, latents,
                  coef[4] != 0.f ? noise : (const float*)nullptr, n, gs, gr, coef[0], coef[1], coef[2], coef[3], coef[4]);
}
EZB_API int ezb_vae_create(ezb_vae** out, const ezb_vae_desc* desc, int device) {
  if (!out || !desc) return fail(EZB_ERR_ARG, "ezb_vae_create: null argument");
  EZB_CUDA(cudaSetDevice(device));
  Vae* h = new Vae();
  int rc = h->init(*desc, &device_ctx(device));
  if (rc != 0) { delete h; return rc; }
  *out = reinterpret_cast<ezb_vae*>(h);
  return EZB_OK;
}
EZB_API int ezb_vae_destroy(ezb_vae* h) {
  delete reinterpret_cast<Vae*>(h);
  return EZB_OK;
}
EZB_API int ezb_vae_load_weight(ezb_vae* h, const char* key, const float* data, const int64_t* shape, int ndim, void* stream) {
  if (!h || !key || !data || !shape) return fail(EZB_ERR_ARG, "ezb_vae_load_weight: null argument");
  return reinterpret_cast<Vae*>(h)->load_weight(key, data, shape, ndim, ST(stream));
}
EZB_API int ezb_vae_finalize_weights(ezb_vae* h, void* stream) {
  if (!h) return fail(EZB_ERR_ARG, "null handle");
  return reinterpret_cast<Vae*>(h)->finalize(ST(stream));
}
EZB_API int ezb_vae_encode(ezb_vae* h, const float* audio, const float* noise, float* z, int B, int T, void* stream) {
  if (!h || !audio || !z) return fail(EZB_ERR_ARG, "ezb_vae_encode: null argument");
  return reinterpret_cast<Vae*>(h)->encode(audio, noise, z, B, T, ST(stream));
}
EZB_API int ezb_t5_create(ezb_t5** out, const ezb_t5_desc* desc, int device) {
  if (!out || !desc) return fail(EZB_ERR_ARG, "ezb_t5_create: null argument");
  EZB_CUDA(cudaSetDevice(device));
  T5* h = new T5();
  int rc = h->init(*desc, &device_ctx(device));
  if (rc != 0) { delete h; return rc; }
  *out = reinterpret_cast<ezb_t5*>(h);
  return EZB_OK;
}
EZB_API int ezb_t5_destroy(ezb_t5* h) {
  delete reinterpret_cast<T5*>(h);
  return EZB_OK;
}
EZB_API int ezb_t5_load_weight(ezb_t5* h, const char* key, const float* data, const int64_t* shape, int ndim, void* stream) {
  if (!h || !key || !data || !shape) return fail(EZB_ERR_ARG, "ezb_t5_load_weight: null argument");
  return reinterpret_cast<T5*>(h)->load_weight(key, data, shape, ndim, ST(stream));
}
EZB_API int ezb_t5_finalize_weights(ezb_t5* h, void* stream) {
  if (!h) return fail(EZB_ERR_ARG, "null handle");
  (void)stream;
  return reinterpret_cast<T5*>(h)->finalize();
}
EZB_API int ezb_t5_forward(ezb_t5* h, const int32_t* ids, const uint8_t* mask, const int32_t* buckets, float* out, int B, int L, void* stream) {
  if (!h || !ids || !mask || !out) return fail(EZB_ERR_ARG, "ezb_t5_forward: null argument");
  return reinterpret_cast<T5*>(h)->forward(ids, mask, buckets, out, B, L, ST(stream));
}
EZB_API int ezb_energy_condition(int device, const float* audio, float* out, int B, int T, int hop, int win, float min_db, int norm, int qlevels,
                                 void* stream) {
  if (!audio || !out) return fail(EZB_ERR_ARG, "ezb_energy_condition: null pointer");
  if (B < 1 || hop < 1 || win < hop || T < hop) return fail(EZB_ERR_SHAPE, "ezb_energy_condition: B=%d T=%d hop=%d window=%d", B, T, hop, win);
  const int pad = (win - hop) / 2, n_frames = T / hop;
  if (pad >= T) return fail(EZB_ERR_SHAPE, "ezb_energy_condition: reflect padding %d needs more than %d samples", pad, T);
  if ((size_t)n_frames * sizeof(float) > 200 * 1024) return fail(EZB_ERR_SHAPE, "ezb_energy_condition: %d frames exceed the shared-memory table", n_frames);
  EZB_CUDA(cudaSetDevice(device));
  EZB_CUDA(cudaFuncSetAttribute(energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ++launch_counter();
  energy_kernel<<<B, 1024, n_frames * sizeof(float), ST(stream)>>>(audio, out, T, n_frames, hop, win, min_db, powf(10.f, min_db / 10.f), norm, qlevels);
  EZB_CUDA(cudaGetLastError());
  return EZB_OK;
}
EZB_API int ezb_wave_prepare(int device, const float* in, float* out, int B, int T_in, int T_out, int normalize, float gate, void* stream) {
  if (!in || !out || B < 1 || T_in < 1 || T_out < 1) return fail(EZB_ERR_ARG, "ezb_wave_prepare: bad argument");
  EZB_CUDA(cudaSetDevice(device));
  ++launch_counter();
  wave_prepare_kernel<<<B, 1024, 0, ST(stream)>>>(in, out, T_in, T_out, 1e-9f, gate, normalize);
  EZB_CUDA(cudaGetLastError());
  return EZB_OK;
}
EZB_API int ezb_wave_splice(int device, float* dst, long long dst_len, const float* src, long long start, long long n, void* stream) {
  if (!dst || !src) return fail(EZB_ERR_ARG, "ezb_wave_splice: null pointer");
  if (start < 0 || n < 0 || start + n > dst_len) return fail(EZB_ERR_SHAPE, "ezb_wave_splice: [%lld, %lld) outside a clip of %lld samples", start, start + n, dst_len);
  if (n == 0) return EZB_OK;
  EZB_CUDA(cudaSetDevice(device));
  ++launch_counter();
  wave_splice_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(dst, src, start, n);
  EZB_CUDA(cudaGetLastError());
  return EZB_OK;
}
EZB_API int ezb_wave_to_pcm16(int device, const float* in, int16_t* out, long long n, void* stream) {
  if (!in || !out || n < 0) return fail(EZB_ERR_ARG, "ezb_wave_to_pcm16: bad argument");
  if (n == 0) return EZB_OK;
  EZB_CUDA(cudaSetDevice(device));
  ++launch_counter();
  wave_to_pcm16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(in, out, n);
  EZB_CUDA(cudaGetLastError());
  return EZB_OK;
}
EZB_API int ezb_vae_decode(ezb_vae* h, const float* z, float* wav, int B, int L, void* stream) {
  if (!h || !z || !wav) return fail(EZB_ERR_ARG, "ezb_vae_decode: null argument");
  return reinterpret_cast<Vae*>(h)->decode(z, wav, B, L, ST(stream));
}
// impl 0: fp32 CUDA-core kernel (q,k,v fp32 [B,H,L,dh]); impl 1: tcgen05 kernel (q,k bf16 [B*H,L,DHP], vt bf16 [B*H,DVP,Lkpad])
EZB_API int ezb_test_attention(int device, const void* q, const void* k, const void* v, const uint8_t* key_mask, void* out, int B, int H, int Lq,
                               int Lk, int dh, int impl, void* stream) {
  if (!q || !k || !v || !out) return fail(EZB_ERR_ARG, "ezb_test_attention: null pointer");
  EZB_CUDA(cudaSetDevice(device));
  const float scale = 1.0f / sqrtf((float)dh);
  if (impl == 0) {
    if (dh % 4) return fail(EZB_ERR_UNSUPPORTED, "fp32 attention: head dimension %d is not a multiple of 4", dh);
    EZB_CUDA(cudaFuncSetAttribute(attn_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    dim3 grid((Lq + SA_WARPS * SA_QW - 1) / (SA_WARPS * SA_QW), B * H);
    ++launch_counter();
    attn_simt_kernel<<<grid, SA_WARPS * 32, attn_simt_smem(dh), ST(stream)>>>(reinterpret_cast<const float*>(q), reinterpret_cast<const float*>(k),
                                                                            reinterpret_cast<const float*>(v), key_mask,
                                                                            reinterpret_cast<__nv_bfloat16*>(out), H, Lq, Lk, dh, scale, 1);
    EZB_CUDA(cudaGetLastError());
    return EZB_OK;
  }
  // impl 1 / 5: q, k rows of `dhp` elements: 64-multiple (round-1 layout) or, with impl >= 100 (impl - 100 = kernel), 80 for dh = 72
  int dhp = (dh + 63) / 64 * 64;
  if (impl >= 100) { impl -= 100; if (dh == 72) dhp = 80; }
  const int dvp = (dh + 15) / 16 * 16, lkpad = (Lk + 7) / 8 * 8;
  if (impl == 7 || (impl == 1 && opt_attn7()))
    return attention_tc7(device_ctx(device), ST(stream), reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
                         reinterpret_cast<const __nv_bfloat16*>(v), key_mask, reinterpret_cast<__nv_bfloat16*>(out), B, H, Lq, Lk, lkpad, dh, dhp, dvp, scale);
  if (impl == 6 || (impl == 1 && (opt_attn6() & 1)))
    return attention_tc6(device_ctx(device), ST(stream), reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
                         reinterpret_cast<const __nv_bfloat16*>(v), key_mask, reinterpret_cast<__nv_bfloat16*>(out), B, H, Lq, Lk, lkpad, dh, dhp, dvp, scale);
  return attention_tc4(device_ctx(device), ST(stream), reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
                       reinterpret_cast<const __nv_bfloat16*>(v), key_mask, reinterpret_cast<__nv_bfloat16*>(out), B, H, Lq, Lk, lkpad, dh, dhp, dvp, scale);
}


// runtime switches (A/B testing): "pair_gemm" 0/1 -- read when a handle is created
EZB_API unsigned long long ezb_option_epoch(void) { return option_epoch(); }
EZB_API int ezb_set_option(const char* name, int value) {
  ++option_epoch();   // captured CUDA graphs bake the kernel selection in: the host layer keys its graph cache on this counter
  if (name && !strcmp(name, "pair_gemm")) { opt_pair_gemm() = value; return EZB_OK; }
  if (name && !strcmp(name, "pdl")) { opt_pdl() = value; return EZB_OK; }
  if (name && !strcmp(name, "swap_ab")) { opt_swap_ab() = value; return EZB_OK; }
  if (name && !strcmp(name, "qkv3")) { opt_qkv3() = value; return EZB_OK; }
  if (name && !strcmp(name, "ln_tail")) { opt_ln_tail() = value; return EZB_OK; }
  if (name && !strcmp(name, "heads_dbg")) { opt_heads_dbg() = value; return EZB_OK; }
  if (name && !strcmp(name, "mlp2_pair")) { opt_mlp2_pair() = value; return EZB_OK; }
  if (name && !strcmp(name, "cq_single")) { opt_cq_single() = value; return EZB_OK; }
  if (name && !strcmp(name, "ksub2")) { opt_ksub2() = value; return EZB_OK; }
  if (name && !strcmp(name, "attn_res")) { opt_attn_res() = value; return EZB_OK; }
  if (name && !strcmp(name, "attn_pp")) { opt_attn_pp() = value; return EZB_OK; }
  if (name && !strcmp(name, "attn6")) { opt_attn6() = value; return EZB_OK; }
  if (name && !strcmp(name, "attn7")) { opt_attn7() = value; return EZB_OK; }
  if (name && !strcmp(name, "w_prefetch")) { opt_w_prefetch() = value; return EZB_OK; }
  if (name && !strcmp(name, "attn_dbg")) { opt_attn_dbg() = value; return EZB_OK; }
  if (name && !strcmp(name, "ln_variant")) { opt_ln_variant() = value; return EZB_OK; }
  if (name && !strcmp(name, "mlp_fused")) { opt_mlp_fused() = value; return EZB_OK; }
  if (name && !strcmp(name, "heads_direct")) { opt_heads_direct() = value; return EZB_OK; }
  if (name && !strcmp(name, "dhp80")) { opt_dhp80() = value; return EZB_OK; }
  if (name && !strcmp(name, "ln_fold")) { opt_fold() = value; return EZB_OK; }
  if (name && !strcmp(name, "skip")) { opt_skip() = value; return EZB_OK; }
  if (name && !strcmp(name, "swap_mc")) { opt_swap_mc() = value; return EZB_OK; }
  if (name && !strcmp(name, "attn_poly")) { opt_attn_poly() = value; return EZB_OK; }
  if (name && !strcmp(name, "rope_mufu")) { opt_rope_mufu() = value; return EZB_OK; }
  if (name && !strcmp(name, "gemm_debug")) {  // cycle counters of CTA 0 of every pair-GEMM launch (accumulated)
    if (value && !gemm_dbg_buf()) { EZB_CUDA(cudaMalloc(&gemm_dbg_buf(), 64)); EZB_CUDA(cudaMemset(gemm_dbg_buf(), 0, 64)); }
    if (!value && gemm_dbg_buf()) { cudaFree(gemm_dbg_buf()); gemm_dbg_buf() = nullptr; }
    return EZB_OK;
  }
  return fail(EZB_ERR_ARG, "unknown option");
}
EZB_API int ezb_debug_read(unsigned long long* out8) {
  if (!gemm_dbg_buf() || !out8) return fail(EZB_ERR_STATE, "gemm_debug is off");
  EZB_CUDA(cudaDeviceSynchronize());
  EZB_CUDA(cudaMemcpy(out8, gemm_dbg_buf(), 64, cudaMemcpyDeviceToHost));
  EZB_CUDA(cudaMemset(gemm_dbg_buf(), 0, 64));
  return EZB_OK;
}
// ---- accounting / profiling hooks (bench.py)
EZB_API unsigned long long ezb_launch_count(void) { return launch_counter(); }
// kernels replayed through a captured CUDA graph never pass the launch helpers: the host layer reports them here
EZB_API void ezb_launch_count_add(unsigned long long n) { launch_counter() += n; }
EZB_API int ezb_prof_gemm_begin(void) {
  GemmProf& gp = gemm_prof();
  gp.on = true; gp.used = 0; gp.flops.clear();
  return EZB_OK;
}
// Synchronises the device; returns the number of GEMM launches recorded since begin, their total algorithmic FLOPs and the
// sum of their CUDA-event durations (ms).
EZB_API int ezb_prof_gemm_end(int* launches, double* flops, double* ms) {
  GemmProf& gp = gemm_prof();
  gp.on = false;
  EZB_CUDA(cudaDeviceSynchronize());
  double f = 0, t = 0;
  for (size_t i = 0; i < gp.flops.size(); ++i) {
    float m = 0.f;
    EZB_CUDA(cudaEventElapsedTime(&m, gp.ev[2 * i], gp.ev[2 * i + 1]));
    f += gp.flops[i]; t += m;
  }
  if (launches) *launches = (int)gp.flops.size();
  if (flops) *flops = f;
  if (ms) *ms = t;
  return EZB_OK;
}

// After ezb_prof_gemm_end: the same statistics restricted to launches with at least `min_flops` algorithmic FLOPs (the dominant
// GEMM of the step is the GEGLU MLP-in projection, the largest single launch).
EZB_API int ezb_prof_gemm_stats(double min_flops, int* launches, double* flops, double* ms) {
  GemmProf& gp = gemm_prof();
  double f = 0, t = 0;
  int n = 0;
  for (size_t i = 0; i < gp.flops.size(); ++i) {
    if (gp.flops[i] < min_flops) continue;
    float m = 0.f;
    EZB_CUDA(cudaEventElapsedTime(&m, gp.ev[2 * i], gp.ev[2 * i + 1]));
    f += gp.flops[i]; t += m; ++n;
  }
  if (launches) *launches = n;
  if (flops) *flops = f;
  if (ms) *ms = t;
  return EZB_OK;
}

}  // extern "C"
