// extern "C" surface of libfpnmt.so (see include/fpnmt.h).
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "engine.cuh"
#include "tgemm.cuh"

namespace fpnmt {
static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
static int g_pdl_mode = 1;   // 0 = off, 1 = every kernel, 2 = tcgen05 GEMM kernels only (fpnmt_config.kernel_opts)
void set_pdl_mode(int mode) { g_pdl_mode = mode; }
bool pdl_enabled() { return g_pdl_mode != 0; }
bool pdl_small_enabled() { return g_pdl_mode == 1; }
}  // namespace fpnmt

using namespace fpnmt;

struct fpnmt_handle {
  Engine* eng;                   // lane 0: every single-batch entry point runs here
  std::vector<Engine*> lanes;    // lanes[0] == eng; fpnmt_submit / fpnmt_collect address the others
  cudaEvent_t last_encode = nullptr;   // encoder-done event of the most recent fpnmt_submit (owned by that lane)
};

extern "C" {

FPNMT_API const char* fpnmt_version(void) {
  return "fpnmt 0.1 (sm_100a; tcgen05 implicit-GEMM igemm_kernel<32|64|128|256> + skinny-row tgemm_kernel<32|64> (cluster LayerNorm epilogue), TMA, TMEM; CUDA-core tail kernels)";
}
FPNMT_API const char* fpnmt_last_error(void) { return g_last_error.c_str(); }

FPNMT_API int fpnmt_create(const fpnmt_config* cfg, int device, fpnmt_handle** out) {
  if (!cfg || !out) {
    set_last_error("fpnmt_create: NULL argument");
    return FPNMT_ERR_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_last_error("fpnmt_create: no CUDA device (this library has no CPU fallback)");
    return FPNMT_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    set_last_error("fpnmt_create: bad device index");
    return FPNMT_ERR_INVALID;
  }
  if (cfg->lanes < 0 || cfg->lanes > 16) {
    set_last_error("fpnmt_create: lanes must be 0..16");
    return FPNMT_ERR_INVALID;
  }
  const int L = cfg->lanes < 1 ? 1 : cfg->lanes;
  fpnmt_handle* h = new fpnmt_handle{nullptr, {}};
  for (int l = 0; l < L; ++l) {
    Engine* e = new (std::nothrow) Engine(*cfg, device);
    int rc = e ? e->init() : FPNMT_ERR_INVALID;
    if (rc) {
      delete e;
      for (size_t i = h->lanes.size(); i-- > 0;) delete h->lanes[i];
      delete h;
      return rc;
    }
    if (!h->lanes.empty()) e->share_weights_of(h->lanes[0]);   // one device copy of the GEMM weights for all lanes
    h->lanes.push_back(e);
  }
  h->eng = h->lanes[0];
  *out = h;
  return FPNMT_OK;
}

FPNMT_API int fpnmt_destroy(fpnmt_handle* h) {
  if (!h) return FPNMT_OK;
  for (size_t i = h->lanes.size(); i-- > 0;) delete h->lanes[i];   // the lead lane (owner of the shared weights) last
  delete h;
  return FPNMT_OK;
}

#define CHECK_H(h)                                 \
  if (!(h) || !(h)->eng) {                         \
    set_last_error("NULL handle");                 \
    return FPNMT_ERR_INVALID;                      \
  }

FPNMT_API int fpnmt_set_weight(fpnmt_handle* h, const char* key, const float* data, const int64_t* shape, int ndim) {
  CHECK_H(h);
  for (Engine* e : h->lanes) {
    const int rc = e->set_weight(key, data, shape, ndim);
    if (rc) return rc;
  }
  return FPNMT_OK;
}
FPNMT_API int fpnmt_finalize_weights(fpnmt_handle* h) {
  CHECK_H(h);
  for (Engine* e : h->lanes) {
    const int rc = e->finalize();
    if (rc) return rc;
  }
  return FPNMT_OK;
}
FPNMT_API int fpnmt_lanes(fpnmt_handle* h) { return h ? (int)h->lanes.size() : 0; }
FPNMT_API int fpnmt_submit(fpnmt_handle* h, int lane, const float* images, int on_host, int early_stop, void* stream) {
  CHECK_H(h);
  if (lane < 0 || lane >= (int)h->lanes.size()) {
    set_last_error("fpnmt_submit: lane out of range (fpnmt_config.lanes)");
    return FPNMT_ERR_INVALID;
  }
  const int rc = h->lanes[lane]->submit(images, on_host, early_stop, (cudaStream_t)stream, h->last_encode);
  if (!rc) h->last_encode = h->lanes[lane]->encode_done_event();
  return rc;
}
FPNMT_API int fpnmt_collect(fpnmt_handle* h, int lane, int32_t* out_ids, int32_t* out_len, int outputs_on_host, void* stream) {
  CHECK_H(h);
  if (lane < 0 || lane >= (int)h->lanes.size()) {
    set_last_error("fpnmt_collect: lane out of range (fpnmt_config.lanes)");
    return FPNMT_ERR_INVALID;
  }
  return h->lanes[lane]->collect(out_ids, out_len, outputs_on_host, (cudaStream_t)stream);
}
FPNMT_API int fpnmt_encode(fpnmt_handle* h, const float* images, int on_host, float* memory_out, void* stream) {
  CHECK_H(h);
  return h->eng->encode(images, on_host, memory_out, (cudaStream_t)stream);
}
FPNMT_API int fpnmt_features(fpnmt_handle* h, const float* images, int on_host, float* const out5[5], void* stream) {
  CHECK_H(h);
  return h->eng->features(images, on_host, out5, (cudaStream_t)stream);
}
FPNMT_API int fpnmt_get_tap(fpnmt_handle* h, const char* name, float* out, size_t capacity, size_t* count, void* stream) {
  CHECK_H(h);
  return h->eng->get_tap(name, out, capacity, count, (cudaStream_t)stream);
}
FPNMT_API int fpnmt_decode_logits(fpnmt_handle* h, const float* memory, const int32_t* tokens, int t, float* logits_out,
                        void* stream) {
  CHECK_H(h);
  return h->eng->decode_logits(memory, tokens, t, logits_out, (cudaStream_t)stream);
}
FPNMT_API int fpnmt_decode_hidden(fpnmt_handle* h, const float* memory, const int32_t* tokens, int t, float* hidden_out,
                                  void* stream) {
  CHECK_H(h);
  if (!hidden_out) {
    set_last_error("decode_hidden: hidden_out is NULL");
    return FPNMT_ERR_INVALID;
  }
  return h->eng->decode_logits(memory, tokens, t, nullptr, (cudaStream_t)stream, hidden_out);
}
FPNMT_API int fpnmt_beam_step(fpnmt_handle* h, const float* logits, const float* scores_in, int32_t* parent, int32_t* token,
                    float* scores_out, void* stream) {
  CHECK_H(h);
  return h->eng->beam_step(logits, scores_in, parent, token, scores_out, (cudaStream_t)stream);
}
FPNMT_API int fpnmt_generate(fpnmt_handle* h, const float* images, int on_host, int32_t* out_ids, int32_t* out_len,
                   int outputs_on_host, int early_stop, float* step_scores, void* stream) {
  CHECK_H(h);
  return h->eng->generate(images, on_host, out_ids, out_len, outputs_on_host, early_stop, step_scores,
                          (cudaStream_t)stream);
}
FPNMT_API int fpnmt_stage_images(fpnmt_handle* h, const float* host_images, int slot) {
  CHECK_H(h);
  return h->eng->stage_images(host_images, slot);
}
FPNMT_API int fpnmt_generate_staged(fpnmt_handle* h, int slot, int32_t* out_ids, int32_t* out_len, int outputs_on_host,
                                    int early_stop, float* step_scores, void* stream) {
  CHECK_H(h);
  return h->eng->generate_staged(slot, out_ids, out_len, outputs_on_host, early_stop, step_scores, (cudaStream_t)stream);
}
FPNMT_API int fpnmt_decode(fpnmt_handle* h, int32_t* out_ids, int32_t* out_len, int outputs_on_host, int early_stop,
                 float* step_scores, void* stream) {
  CHECK_H(h);
  return h->eng->decode(out_ids, out_len, outputs_on_host, early_stop, step_scores, (cudaStream_t)stream);
}
FPNMT_API int fpnmt_profile(fpnmt_handle* h, int iters, char* buf, size_t cap) {
  CHECK_H(h);
  return h->eng->profile(iters, buf, cap);
}
FPNMT_API int64_t fpnmt_launch_count(fpnmt_handle* h) {
  if (!h || !h->eng) return 0;
  int64_t n = 0;
  for (Engine* e : h->lanes) n += e->launches;
  return n;
}

// ---- stand-alone convolution operator -----------------------------------------------------------------
static inline uint16_t f2bf_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf2f_host(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

FPNMT_API int fpnmt_op_conv2d(int device, int precision, const float* x, int N, int H, int W, int Cin, const float* kernel, int kh,
                    int kw, int Cout, int pad_top, int pad_left, const float* bias, int act, const float* residual,
                    int res_mode, float* out, int force_bn, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (Cin % 8) {
    set_last_error("op_conv2d: Cin must be a multiple of 8");
    return FPNMT_ERR_INVALID;
  }
  FPNMT_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  FPNMT_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_last_error("op_conv2d: device is not sm_100");
    return FPNMT_ERR_CUDA;
  }
  int rc = igemm_set_attributes();
  if (rc) return rc;
  const bool split = precision == FPNMT_PREC_BF16X3;
  const int K = kh * kw * Cin;
  const size_t ldw = split ? 2 * (size_t)K : (size_t)K;
  std::vector<uint16_t> hw((size_t)Cout * ldw);
  for (int t = 0; t < kh * kw; ++t)
    for (int ci = 0; ci < Cin; ++ci)
      for (int o = 0; o < Cout; ++o) {
        const float f = kernel[((size_t)t * Cin + ci) * Cout + o];
        const uint16_t hi = f2bf_host(f);
        hw[(size_t)o * ldw + (size_t)t * Cin + ci] = hi;
        if (split) hw[(size_t)o * ldw + K + (size_t)t * Cin + ci] = f2bf_host(f - bf2f_host(hi));
      }
  std::vector<void*> tmp;
  auto dal = [&](size_t b) {
    void* p = nullptr;
    cudaMalloc(&p, b ? b : 16);
    tmp.push_back(p);
    return p;
  };
  auto cleanup = [&]() {
    for (void* p : tmp) cudaFree(p);
  };
  bf16* dw = (bf16*)dal(hw.size() * 2);
  float* dbias = nullptr;
  if (bias) {
    dbias = (float*)dal((size_t)(Cout + 32) * 4);
    cudaMemsetAsync(dbias, 0, (size_t)(Cout + 32) * 4, s);
    cudaMemcpyAsync(dbias, bias, (size_t)Cout * 4, cudaMemcpyHostToDevice, s);
  }
  cudaMemcpyAsync(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice, s);
  auto mk = [&](size_t pix, int C) {
    Act a;
    const int cs = C < 8 ? 8 : (C + 7) / 8 * 8;
    a.C = C;
    a.ld = split ? 2 * cs : cs;
    a.lo = split ? cs : 0;
    a.p = (bf16*)dal(pix * a.ld * 2);
    return a;
  };
  const size_t pix = (size_t)N * H * W;
  Act ax = mk(pix, Cin), ao = mk(pix, Cout), ar{nullptr, 0, 0, 0};
  rc = launch_f32_to_act(x, pix, Cin, ax, s);
  if (!rc && residual) {
    const size_t rpix = res_mode == RES_UP2 ? (size_t)N * (H / 2) * (W / 2) : pix;
    ar = mk(rpix, Cout);
    rc = launch_f32_to_act(residual, rpix, Cout, ar, s);
  }
  IgemmOp op;
  ConvGeom g{N, H, W, Cin, Cout, kh, kw, pad_top, pad_left};
  if (!rc)
    rc = make_igemm_op(&op, g, ax, dw, split, dbias, act, ao, nullptr, 0, residual ? res_mode : RES_NONE, ar,
                       prop.multiProcessorCount, force_bn);
  if (!rc) rc = igemm_launch(op, s);
  if (!rc) rc = launch_act_to_f32(ao, pix, out, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cleanup();
  if (!rc && e != cudaSuccess) {
    set_last_error(std::string("op_conv2d: ") + cudaGetErrorString(e));
    return FPNMT_ERR_CUDA;
  }
  return rc;
}

// ---- input preprocessing operator ------------------------------------------------------------------------
FPNMT_API int fpnmt_op_preprocess(int device, const uint8_t* images_hwc, int N, int H, int W, int S, float* out, void* stream) {
  if (!images_hwc || !out || N < 1 || H < 1 || W < 1 || S < 1) {
    set_last_error("op_preprocess: bad arguments");
    return FPNMT_ERR_INVALID;
  }
  FPNMT_CUDA_OK(cudaSetDevice(device));
  return launch_preprocess(images_hwc, N, H, W, S, out, (cudaStream_t)stream);
}

// ---- stand-alone skinny-row Dense operator (tgemm) -----------------------------------------------------
FPNMT_API int fpnmt_op_dense(int device, int precision, const float* x, int R, int K, const float* kernel, int F,
                   const float* bias, int act, const float* residual, const float* gamma, const float* beta, float eps,
                   float* out, int force_bn, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (K % 8 || F % 8) {
    set_last_error("op_dense: K and F must be multiples of 8");
    return FPNMT_ERR_INVALID;
  }
  FPNMT_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  FPNMT_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_last_error("op_dense: device is not sm_100");
    return FPNMT_ERR_CUDA;
  }
  int rc = tgemm_set_attributes();
  if (!rc) rc = tgemmw_set_attributes();
  if (rc) return rc;
  const bool split = precision == FPNMT_PREC_BF16X3;
  const bool wide = force_bn >= 128;          // the wide-row two-CTAs-per-SM kernel (tgemmw.cu)
  if (wide && !tgemmw_supports(split, residual != nullptr, gamma != nullptr, F, K)) {
    set_last_error("op_dense: the wide-row kernel needs bf16 precision and a residual only together with LayerNorm");
    return FPNMT_ERR_INVALID;
  }
  const size_t ldw = split ? 2 * (size_t)K : (size_t)K;
  std::vector<uint16_t> hw((size_t)F * ldw);
  for (int k = 0; k < K; ++k)
    for (int o = 0; o < F; ++o) {
      const float f = kernel[(size_t)k * F + o];          // Keras Dense kernel (in, out)
      const uint16_t hi = f2bf_host(f);
      hw[(size_t)o * ldw + k] = hi;
      if (split) hw[(size_t)o * ldw + K + k] = f2bf_host(f - bf2f_host(hi));
    }
  std::vector<void*> tmp;
  auto dal = [&](size_t b) {
    void* p = nullptr;
    cudaMalloc(&p, b ? b : 16);
    tmp.push_back(p);
    return p;
  };
  bf16* dw = (bf16*)dal(hw.size() * 2);
  cudaMemcpyAsync(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice, s);
  auto upl = [&](const float* h, int n) {
    float* d = (float*)dal((size_t)(n + 128) * 4);
    cudaMemsetAsync(d, 0, (size_t)(n + 128) * 4, s);
    cudaMemcpyAsync(d, h, (size_t)n * 4, cudaMemcpyHostToDevice, s);
    return d;
  };
  float* dbias = bias ? upl(bias, F) : nullptr;
  float* dg = gamma ? upl(gamma, F) : nullptr;
  float* db = beta ? upl(beta, F) : nullptr;
  auto mk = [&](size_t rows, int C) {
    Act a;
    a.C = C;
    a.ld = split ? 2 * C : C;
    a.lo = split ? C : 0;
    a.p = (bf16*)dal(rows * a.ld * 2);
    return a;
  };
  Act ax = mk(R, K), ao = mk(R, F), ar{nullptr, 0, 0, 0};
  rc = launch_f32_to_act(x, R, K, ax, s);
  if (!rc && residual) {
    ar = mk(R, F);
    rc = launch_f32_to_act(residual, R, F, ar, s);
  }
  TgemmOp op;
  if (!rc && wide)
    rc = make_tgemmw_op(&op, R, ax, dw, F, K, dbias, act, ao, nullptr, 0, residual ? &ar : nullptr, dg, db, eps, prop.multiProcessorCount);
  else if (!rc)
    rc = make_tgemm_op(&op, R, ax, dw, F, K, split, dbias, act, ao, nullptr, 0, residual ? &ar : nullptr, dg, db, eps,
                       prop.multiProcessorCount, force_bn);
  if (!rc) rc = tgemm_launch(op, s);
  if (!rc) rc = launch_act_to_f32(ao, R, out, s);
  cudaError_t e = cudaStreamSynchronize(s);
  for (void* p : tmp) cudaFree(p);
  if (!rc && e != cudaSuccess) {
    set_last_error(std::string("op_dense: ") + cudaGetErrorString(e));
    return FPNMT_ERR_CUDA;
  }
  return rc;
}

}  // extern "C"
