// Fused stem convolution (bf16 mode): the first convolution of the backbones reads the fp32 NHWC image itself
// (Cin = 3, stride 2) and writes bf16 NHWC activations; BatchNorm is folded into the weights, the activation is applied in
// the epilogue.  Replaces the `stem_im2col` + `stem_conv` pair (an explicit [pixels][152] im2col matrix in HBM, 1.3 GB of
// extra traffic for a 64 x 512 x 512 batch) for
//   ResNet-50 / DenseNet-121  conv1   7x7 s2 pad 3, 3 -> 64   (keras.applications resnet.py / densenet.py stems, as built by
//                                     models/resnet.py, models/densenet.py of the reference)
//   MobileNetV2               Conv1   3x3 s2 pad (0,1), 3 -> 32  (models/mobilenet.py)
// One persistent CTA per SM walks 2 x 64 output-pixel tiles: producer warps stage the input patch with 16-byte cp.async
// and build the 128 x K im2col tile directly in the 128B-swizzled K-major shared-memory layout tcgen05.mma expects, one
// thread issues the MMAs (M = 128 pixels, N = Cout) into a double-buffered TMEM accumulator, four epilogue warps add the
// bias, apply the activation and write 4 KB contiguous bf16 per warp.
#pragma once
#include "common.cuh"

namespace fpnmt {

struct StemParams {
  const float* const* img_slot;   // device slot holding the image pointer (fp32 NHWC, 16-byte aligned)
  int N, H, W, Ho, Wo;
  int tiles_x, tiles_y, tiles;    // 64-pixel columns, 2-pixel rows, total tiles
  const float* bias;              // [Cout] (BatchNorm shift)
  int act;
  Act out;                        // [N*Ho*Wo][Cout] bf16, ld == Cout
  long long* dbg;                 // optional globaltimer stamps (FPNMT_DBG_OP), else nullptr
};

struct StemOp {
  CUtensorMap tmW;
  StemParams p;
  int kh, cout, grid;
  double flops;
};

// wt: device [Cout][Kp] bf16, k = (ky*KH + kx)*3 + c, zero padded to Kp; wt_s2d: device buffer of Cout * ceil(KH/2)^2 * 12
// bf16 that receives the same weights in the kernel's space-to-depth K order (filled here, synchronously).
int make_stem_op(StemOp* op, const float* const* img_slot, int N, int H, int W, int kh, int pad, int cout, const bf16* wt, int Kp,
                 bf16* wt_s2d, const float* bias, int act, const Act& out, int num_sms);
int stem_set_attributes();
int stem_launch(const StemOp& op, cudaStream_t stream);

}  // namespace fpnmt
