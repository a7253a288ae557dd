/*
 * fpnmt.h — C ABI of libfpnmt.so, the B200 (sm_100a) caption-inference engine.
 *
 * The reference (samkoesnadi/fpn-MT-image-captioning) is pure Python/TensorFlow and has no FFI; the
 * boundary this library sits behind is the reference's Python builder API.  Each entry point names the
 * reference interface it replaces (paths relative to the reference repo).  The host-side Python mirror
 * (fpn-mt-image-captioning_b200/fpnmt) binds these symbols with ctypes; see INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only (no C++/torch types); every function returns 0 on success or
 * an FPNMT_ERR_* code and leaves a message retrievable with fpnmt_last_error(); device pointers are raw
 * CUDA device addresses (e.g. torch.Tensor.data_ptr() or a DLPack capsule's data field) on the engine's
 * device; all work is enqueued on the caller's stream (cudaStream_t passed as void*; NULL = default stream);
 * one handle per GPU; calls on one handle must be serialised by the caller.  There is no CPU fallback.
 */
#ifndef FPNMT_H_
#define FPNMT_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define FPNMT_API __attribute__((visibility("default")))
#else
#define FPNMT_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

enum {
  FPNMT_OK = 0,
  FPNMT_ERR_INVALID = 1,     /* bad argument / shape / configuration            */
  FPNMT_ERR_STATE = 2,       /* call order (e.g. encode before finalize)        */
  FPNMT_ERR_CUDA = 3,        /* CUDA runtime / driver error                     */
  FPNMT_ERR_MISSING = 4      /* a required weight was never set                 */
};

enum { FPNMT_BACKBONE_MOBILENETV2 = 0, FPNMT_BACKBONE_RESNET50 = 1, FPNMT_BACKBONE_DENSENET121 = 2 };
enum { FPNMT_PREC_BF16 = 0,      /* bf16 operands, fp32 accumulate, bf16 activations                          */
       FPNMT_PREC_BF16X3 = 1 };  /* 3-term split-bf16 products (~fp32 operand precision), fp32 accumulate:
                                    the parity mode; same kernels, 3x the tensor work                        */
enum { FPNMT_SCORE_LOG = 0,      /* beam score = sum of log-probs                                             */
       FPNMT_SCORE_PROB = 1 };   /* beam score = product of softmax probabilities (reference, pipeline.py:117-123) */
/* fpnmt_config.kernel_opts: each bit switches ONE fused kernel back to its unfused equivalent (A/B timing and the
 * fused-vs-unfused parity tests).  0 = every fused path on (the product configuration). */
enum { FPNMT_OPT_NO_XATTN = 1,       /* separate q2 / cross-attention / o2+LN kernels instead of xattn_kernel          */
       FPNMT_OPT_NO_STEM = 2,        /* explicit im2col + GEMM stem instead of stem_kernel                              */
       FPNMT_OPT_NO_TGEMM = 4,       /* decoder Dense layers through igemm + separate LayerNorm                         */
       FPNMT_OPT_ENC_ATT_SIMT = 8,   /* CUDA-core encoder attention instead of the tensor-core flash kernel             */
       FPNMT_OPT_KSPLIT2 = 16,       /* split-K over an 8-CTA cluster for the K = 2048 LayerNorm Dense (slower on B200)  */
       FPNMT_OPT_NO_PDL = 32,        /* no programmatic dependent launch (process-wide: the last created engine wins)    */
       FPNMT_OPT_PDL_GEMM_ONLY = 64, /* only the tcgen05 GEMM kernels launch early                                      */
       FPNMT_OPT_DEC_ATT_SIMT = 256, /* CUDA-core decode self-attention in bf16 mode (instead of the mma.sync kernel)          */
       FPNMT_OPT_TGEMM_WIDE = 512,   /* bf16 Dense layers on tgemmw_kernel (128 rows per CTA, two CTAs per SM, TMA-store
                                        epilogue: least SM time, best when several lanes share the GPU) even with lanes < 2.
                                        Default: tgemmw_kernel when fpnmt_config.lanes >= 2, else tgemm_kernel (lowest
                                        latency of a single chain)                                                         */
       FPNMT_OPT_NO_TGEMM_WIDE = 1024, /* never tgemmw_kernel                                                             */
       FPNMT_OPT_NO_DENSE_1X1 = 8192, /* 1x1 convolutions in the 2-D pixel-tile geometry of the other convolutions (default: as a
                                        GEMM over 128 consecutive pixels per tile)                                          */
       FPNMT_OPT_NO_B_STATIONARY = 32768, /* igemm_kernel streams the weight chunks with every tile (default: a panel <= 96 KB of a
                                        single output-channel tile stays in shared memory for all tiles of the CTA)            */
       FPNMT_OPT_NO_TMA_STORE = 16384, /* igemm_kernel writes its bf16 outputs with 16-byte LSU stores instead of bulk tensor stores */
       FPNMT_OPT_NO_VSTATS = 4096,   /* keep the [rows][V] fp32 logits between the vocabulary projection and k_beam_step (default with
                                        the wide Dense kernels, log scores and beam <= 8: per-tile softmax partials + 8 candidates)  */
       FPNMT_OPT_NO_KV_SHARE = 2048, /* ancestry cache mode: every beam reads its own lineage's cache rows even when another
                                        beam of the image has the same token history (default: read the first such beam's rows;
                                        identical bits, 8x less cache traffic under the reference's beam initialisation)    */
       FPNMT_OPT_DSTEP_TAPS = 128    /* fused decoder: also keep every layer's LayerNorm outputs (fpnmt_get_tap "decL_outK") */ };
enum { FPNMT_CACHE_ANCESTRY = 0,     /* KV cache never moves; an ancestry table maps (beam, position) -> physical row    */
       FPNMT_CACHE_PHYSICAL = 1 };   /* KV cache rows are gathered by beam parent after every step (bandwidth kernel)    */
enum { FPNMT_DECODE_AUTO = 0,        /* the faster path for the configuration: today the per-operator chain                */
       FPNMT_DECODE_CHAIN = 1,       /* per-operator kernel chain (tgemm / attention / xattn / beam kernels), 38 per step   */
       FPNMT_DECODE_FUSED = 2 };     /* group-stationary fused decoder (dstep_kernel): ONE launch runs every layer, the
                                        vocabulary projection and the beam tail of all steps (bf16, log scores, beam <= 16) */

typedef struct fpnmt_config {
  int32_t backbone;      /* FPNMT_BACKBONE_* — models/mobilenet.py:43, models/resnet.py:78, models/densenet.py:73 */
  int32_t image_size;    /* IMAGE_INPUT_SIZE, common/common_definitions.py:18 (multiple of 256)                  */
  int32_t batch;         /* images per call (the reference predicts one image at a time, pipeline.py:93)         */
  int32_t beam;          /* BEAM_SEARCH_N, common/common_definitions.py:22                                       */
  int32_t vocab;         /* target_vocab_size, utils/pipeline.py:19                                              */
  int32_t max_len;       /* max_seq_len, utils/pipeline.py:12 / test.py:13                                       */
  int32_t num_layers;    /* common/common_definitions.py:56                                                      */
  int32_t d_model;       /* :57 (must be 512)                                                                    */
  int32_t num_heads;     /* :59 (must be 8)                                                                      */
  int32_t dff;           /* :58                                                                                  */
  int32_t precision;     /* FPNMT_PREC_*                                                                         */
  int32_t score_mode;    /* FPNMT_SCORE_*                                                                        */
  int32_t start_id;      /* tokenizer.word_index['<start>'], pipeline.py:89                                      */
  int32_t end_id;        /* tokenizer.word_index['<end>'],   pipeline.py:90                                      */
  int32_t true_beam;     /* 0 = reference init (all beams identical, pipeline.py:101-102); 1 = only beam 0 alive  */
  int32_t use_graphs;    /* 1 = replay the encode / decode-step programs as CUDA graphs                          */
  /* ---- the eight words below were `reserved[8]` in ABI 0.1; all-zero selects the defaults ---- */
  int32_t kernel_opts;   /* FPNMT_OPT_* bit mask                                                                  */
  int32_t cache_mode;    /* FPNMT_CACHE_*                                                                         */
  int32_t decode_path;   /* FPNMT_DECODE_*                                                                        */
  float length_penalty;  /* EXTENSION (not in the reference): alpha of the length normalisation applied when a beam
                            finishes, score / ((5 + len) / 6)^alpha; 0 = reference ordering                        */
  int32_t finished_beams;/* EXTENSION: 1 = a beam that emitted <end> is frozen and competes with its final score;
                            the image stops when its best beam is a finished one.  0 = reference (pipeline.py:143-148:
                            stop the moment the top beam emits <end>, finished lower beams keep decoding)          */
  int32_t dec_groups;    /* DECODE_CHAIN only: cut the batch into this many concurrently decoded chains (0/1 = one) */
  int32_t lanes;         /* batches in flight (fpnmt_submit / fpnmt_collect): the handle holds this many complete engines
                            (own activations, KV cache, beam state, streams; ONE shared copy of the GEMM weights) on the one GPU, 0/1 = one,
                            at most 16.  Two lanes overlap
                            the throughput-bound encoder of batch i+1 with the latency-bound decode of batch i.            */
  int32_t reserved[1];
} fpnmt_config;

typedef struct fpnmt_handle fpnmt_handle;

/* Library / build information ("sm_100a", kernel list).  Never fails. */
FPNMT_API const char* fpnmt_version(void);
/* Thread-local message of the last failing call. */
FPNMT_API const char* fpnmt_last_error(void);

/* Replaces: Transformer(...) / Encoder(...) / FeatureExtractor(...) construction
 * (models/transformer.py:344-357, :246-264; models/retinanet.py:266-304). */
FPNMT_API int fpnmt_create(const fpnmt_config* cfg, int device, fpnmt_handle** out);
FPNMT_API int fpnmt_destroy(fpnmt_handle* h);

/* Replaces: checkpoint restore / load_weights (utils/pipeline.py:38-48, models/retinanet.py:277-278).
 * `key` is the variable path of SURVEY.md Appendix B (e.g. "transformer/decoder/dec_layers/0/mha1/wq/kernel");
 * `data` is HOST float32 in Keras layout (Dense (in,out); Conv2D (kh,kw,Cin,Cout); Depthwise (kh,kw,C,1));
 * the library copies it. */
FPNMT_API int fpnmt_set_weight(fpnmt_handle* h, const char* key, const float* data, const int64_t* shape, int ndim);
/* Folds BatchNorm, concatenates projections, converts to bf16 (split into hi/lo for BF16X3), uploads, builds
 * the kernel programs and tensor maps. */
FPNMT_API int fpnmt_finalize_weights(fpnmt_handle* h);

/* Replaces: Encoder.call(x, training=False, mask=None) (models/transformer.py:266-303).
 * images: float32 NHWC [batch, S, S, 3] in [-1,1] (dataset.py:19-26), DEVICE pointer unless images_on_host != 0
 * (then a pinned/pageable HOST pointer; the H2D copy is enqueued on `stream`).
 * memory_out: optional DEVICE float32 [batch, 16*(S/512)^2.., d_model] = encoder output; may be NULL. */
FPNMT_API int fpnmt_encode(fpnmt_handle* h, const float* images, int images_on_host, float* memory_out, void* stream);

/* Replaces: FeatureExtractor.call(inp) (models/retinanet.py:306-307).  Runs the CNN part only and copies the five
 * head outputs (P3..P7 order, NHWC float32, DEVICE pointers, sizes batch*(S/16>>i)^2*d_model) out. */
FPNMT_API int fpnmt_features(fpnmt_handle* h, const float* images, int images_on_host, float* const out5[5], void* stream);

/* Copies a named intermediate of the last fpnmt_encode (e.g. "C3","C4","C5","P3".."P7","tokens0".."tokens4",
 * "enc_layer0"..) to a DEVICE float32 buffer of `capacity` floats; writes its element count to *count. */
FPNMT_API int fpnmt_get_tap(fpnmt_handle* h, const char* name, float* out, size_t capacity, size_t* count, void* stream);

/* Replaces: Transformer.call(inp=enc_output, tar, training=False, look_ahead_mask) (models/transformer.py:359-374)
 * for teacher-forced parity: memory DEVICE float32 [batch,16,d_model] (NULL = reuse the last fpnmt_encode result),
 * tokens DEVICE int32 [batch, t]; logits_out DEVICE float32 [batch, t, vocab].  Computed with the KV-cached
 * step program (position by position), which is mathematically the causal-masked full forward. */
FPNMT_API int fpnmt_decode_logits(fpnmt_handle* h, const float* memory, const int32_t* tokens, int t, float* logits_out,
                        void* stream);

/* Replaces: Decoder.call(x, enc_output, training=False, look_ahead_mask, padding_mask) (models/transformer.py:321-341): the
 * output of the last decoder layer (before final_layer) for teacher-forced tokens; arguments as fpnmt_decode_logits,
 * hidden_out DEVICE float32 [batch, t, d_model].  (The attention-weights dict is not returned: SURVEY §8b1.) */
FPNMT_API int fpnmt_decode_hidden(fpnmt_handle* h, const float* memory, const int32_t* tokens, int t, float* hidden_out,
                                  void* stream);

/* Replaces: the body of the decode loop, utils/pipeline.py:115-141, on caller-provided logits (the decode-tail
 * kernels alone): logits DEVICE float32 [batch*beam, vocab], scores_in DEVICE float32 [batch*beam];
 * outputs DEVICE int32 parent[batch*beam], token[batch*beam], float32 scores_out[batch*beam]. */
FPNMT_API int fpnmt_beam_step(fpnmt_handle* h, const float* logits, const float* scores_in, int32_t* parent, int32_t* token,
                    float* scores_out, void* stream);

/* Replaces: Pipeline.predict (utils/pipeline.py:82-154) for a batch: encoder + beam-search decode.
 * out_ids DEVICE-or-HOST int32 [batch, max_len] (0 padded), out_len int32 [batch]: exactly what predict() returns
 * per image (<start> stripped, trailing <end> stripped).  outputs_on_host != 0: results are copied to host and
 * the call synchronises the stream.  early_stop != 0: stop when every image's top beam has emitted <end>
 * (pipeline.py:147); 0: always run max_len steps (throughput runs).
 * step_scores: optional DEVICE float32 [max_len, batch] score of the top beam after each step (may be NULL). */
FPNMT_API int fpnmt_generate(fpnmt_handle* h, const float* images, int images_on_host, int32_t* out_ids, int32_t* out_len,
                   int outputs_on_host, int early_stop, float* step_scores, void* stream);
/* Replaces: the input overlap the reference gets from tf.data (dataset.py:90-92, map(load_image, AUTOTUNE) +
 * prefetch(AUTOTUNE)): double-buffered HOST input.  fpnmt_stage_images enqueues the host->device copy of a whole batch
 * (HOST float32 [batch, S, S, 3]; pinned memory for a truly asynchronous copy) into staging slot 0 or 1 on the engine's
 * own copy stream and returns at once; the buffer must stay valid until the matching fpnmt_generate_staged has been
 * issued and its stream synchronised.  fpnmt_generate_staged is fpnmt_generate on a staged slot: `stream` waits for that
 * slot's copy, runs encoder + decode, and releases the slot as soon as the encoder has consumed it.  Typical loop:
 * stage(0); for i: stage((i+1)&1, batch i+1); generate_staged(i&1, ...).  Error: FPNMT_ERR_STATE if the slot is empty. */
FPNMT_API int fpnmt_stage_images(fpnmt_handle* h, const float* host_images, int slot);
FPNMT_API int fpnmt_generate_staged(fpnmt_handle* h, int slot, int32_t* out_ids, int32_t* out_len, int outputs_on_host,
                                    int early_stop, float* step_scores, void* stream);
/* Replaces: the caller's loop over batches (utils/pipeline.py:156-175 evaluate; test.py:14-21) with `lanes` batches in
 * flight.  fpnmt_submit enqueues one whole batch (host->device copy when images_on_host, encoder, decode) on lane `lane`'s
 * own stream and returns at once; fpnmt_collect hands that batch's result over (outputs as fpnmt_generate; host outputs
 * synchronise the lane's stream, device outputs make `stream` wait for them).  The image buffer must stay valid until the
 * matching collect; device images are ordered after the work already enqueued on `stream`.  With early_stop != 0 the
 * decode needs the host to poll the finished-image counter: submit enqueues the encoder only and collect runs the decode
 * (the encoder of the batch submitted on another lane meanwhile still overlaps it).  A lane holds one batch at a time:
 * FPNMT_ERR_STATE on a second submit before collect, or a collect without submit.  Results are bit-identical to
 * fpnmt_generate on the same images (same kernels, same order per lane).  Typical loop, L = fpnmt_lanes(h):
 *   for i: if (i >= L) collect(i % L, ...batch i-L...); submit(i % L, batch i);   then collect the last L. */
FPNMT_API int fpnmt_lanes(fpnmt_handle* h);
FPNMT_API int fpnmt_submit(fpnmt_handle* h, int lane, const float* images, int images_on_host, int early_stop, void* stream);
FPNMT_API int fpnmt_collect(fpnmt_handle* h, int lane, int32_t* out_ids, int32_t* out_len, int outputs_on_host, void* stream);
/* Same, from an encoder output already in the engine (after fpnmt_encode) — the decode half only. */
FPNMT_API int fpnmt_decode(fpnmt_handle* h, int32_t* out_ids, int32_t* out_len, int outputs_on_host, int early_stop,
                 float* step_scores, void* stream);

/* Per-kernel timing of the encode and decode-step programs (CUDA events, `iters` repetitions each).  Writes a JSON
 * document into buf (NUL terminated, truncated to cap).  Reports algorithmic FLOPs/bytes per kernel. */
FPNMT_API int fpnmt_profile(fpnmt_handle* h, int iters, char* buf, size_t cap);
/* Number of kernel launches issued by this handle since creation (graph replays count their kernel nodes). */
FPNMT_API int64_t fpnmt_launch_count(fpnmt_handle* h);

/* ---- stand-alone operators (used by the parity tests to check single kernels) ------------------------------- */
/* out[pix, co] = act(conv(x) + bias [+ residual]); x DEVICE float32 NHWC [N,H,W,Cin]; kernel HOST float32
 * (kh,kw,Cin,Cout); bias HOST float32 [Cout] or NULL; residual DEVICE float32 (same shape as out, or
 * [N,H/2,W/2,Cout] when res_mode == 2) or NULL; out DEVICE float32 [N,H,W,Cout].  stride 1, zero padding
 * (pad_top, pad_left; the rest implied by the output size == input size). act: 0 none,1 relu,2 leaky(0.2),3 relu6 */
FPNMT_API int fpnmt_op_conv2d(int device, int precision, const float* x, int N, int H, int W, int Cin, const float* kernel,
                    int kh, int kw, int Cout, int pad_top, int pad_left, const float* bias, int act,
                    const float* residual, int res_mode, float* out, int force_bn, void* stream);

/* Replaces: dataset.load_image after the JPEG decode (dataset.py:21-24): tf.image.resize(img, (S, S)) — bilinear,
 * half-pixel centres, no antialias — followed by mobilenet_v2.preprocess_input (x / 127.5 - 1).
 * images_hwc DEVICE uint8 [N, H, W, 3]; out DEVICE float32 NHWC [N, S, S, 3] in [-1, 1] (the layout fpnmt_encode takes). */
FPNMT_API int fpnmt_op_preprocess(int device, const uint8_t* images_hwc, int N, int H, int W, int S, float* out, void* stream);

/* Replaces: dataset.load_image as a whole (dataset.py:19-26) for a batch of JPEG byte strings: tf.io.read_file's result ->
 * tf.image.decode_jpeg(channels=3) -> resize to (S, S) -> x / 127.5 - 1.  jpegs: n HOST pointers to the encoded files,
 * lengths their byte counts; decode runs on the GPU (nvJPEG), the RGB image never visits the host; out DEVICE float32 NHWC
 * [n, S, S, 3]; sizes_out optional HOST int32 [n][2] = decoded (height, width).  FPNMT_ERR_INVALID names the first image that
 * is not a decodable JPEG.  Work is enqueued on `stream`; the encoded bytes must stay valid until it has run. */
FPNMT_API int fpnmt_op_decode_jpeg(int device, const uint8_t* const* jpegs, const size_t* lengths, int n, int S, float* out,
                                   int32_t* sizes_out, void* stream);

/* out[r, f] = epilogue( x[r, :] @ kernel[:, f] + bias[f] [+ residual[r, f]] ) with the skinny-row Dense kernel of the
 * decoder step (tf.keras.layers.Dense, models/transformer.py:117-122, 165-168, 211-214, 357): x DEVICE float32 [R, K];
 * kernel HOST float32 (K, F) Keras layout; bias HOST [F] or NULL; residual DEVICE float32 [R, F] or NULL; out DEVICE
 * float32 [R, F].  gamma/beta HOST [F] != NULL (F must be 512): the epilogue is LayerNormalization(epsilon=eps) of
 * (dense + residual) (models/transformer.py:192,198,230,235,241), computed by a 4-CTA cluster. act as fpnmt_op_conv2d.
 * force_bn: 0 = automatic, 32 / 64 = rows per CTA of tgemm_kernel, 128 = the wide-row kernel of FPNMT_OPT_TGEMM_WIDE. */
FPNMT_API int fpnmt_op_dense(int device, int precision, const float* x, int R, int K, const float* kernel, int F,
                   const float* bias, int act, const float* residual, const float* gamma, const float* beta, float eps,
                   float* out, int force_bn, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FPNMT_H_ */
