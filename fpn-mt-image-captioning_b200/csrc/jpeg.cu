// GPU JPEG decode for the input side of the hot path: replaces the first half of dataset.load_image
// (/root/reference/dataset.py:19-21, tf.io.read_file + tf.image.decode_jpeg(channels=3)) with nvJPEG and feeds k_preprocess
// (bilinear resize + x / 127.5 - 1, dataset.py:22-24), so a JPEG byte string goes to the engine's NHWC float input without
// the decoded image ever visiting the host.  nvJPEG is a CUDA-toolkit library (like cuBLAS: library code, not a kernel of
// this repo); the resize / normalise kernel is ours.
#include <nvjpeg.h>

#include <mutex>
#include <vector>

#include "../../include/fpnmt.h"
#include "kernels.cuh"

namespace fpnmt {

struct JpegCtx {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  uint8_t* rgb = nullptr;      // device staging for the decoded RGB images of one call
  size_t rgb_cap = 0;
  int device = -1;
};
static std::mutex g_jpeg_mu;
static JpegCtx g_jpeg[16];

static int jpeg_fail(const char* what, int st) {
  set_last_error(std::string("nvjpeg: ") + what + " failed with status " + std::to_string(st));
  return FPNMT_ERR_CUDA;
}

}  // namespace fpnmt

using namespace fpnmt;

extern "C" FPNMT_API int fpnmt_op_decode_jpeg(int device, const uint8_t* const* jpegs, const size_t* lengths, int n, int S,
                                              float* out, int32_t* sizes_out, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!jpegs || !lengths || !out || n < 1 || S < 1 || device < 0 || device >= 16) {
    set_last_error("op_decode_jpeg: bad arguments");
    return FPNMT_ERR_INVALID;
  }
  FPNMT_CUDA_OK(cudaSetDevice(device));
  std::lock_guard<std::mutex> lock(g_jpeg_mu);
  JpegCtx& c = g_jpeg[device];
  if (!c.handle) {
    nvjpegStatus_t st = nvjpegCreateSimple(&c.handle);
    if (st != NVJPEG_STATUS_SUCCESS) return jpeg_fail("nvjpegCreateSimple", (int)st);
    st = nvjpegJpegStateCreate(c.handle, &c.state);
    if (st != NVJPEG_STATUS_SUCCESS) return jpeg_fail("nvjpegJpegStateCreate", (int)st);
    c.device = device;
  }
  // pass 1: headers -> sizes and staging offsets
  std::vector<int> W(n), H(n);
  std::vector<size_t> off(n);
  size_t total = 0;
  for (int i = 0; i < n; ++i) {
    int ncomp = 0, w[NVJPEG_MAX_COMPONENT], h[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t sub;
    const nvjpegStatus_t st = nvjpegGetImageInfo(c.handle, jpegs[i], lengths[i], &ncomp, &sub, w, h);
    if (st != NVJPEG_STATUS_SUCCESS) {
      set_last_error("op_decode_jpeg: image " + std::to_string(i) + " is not a decodable JPEG (status " + std::to_string((int)st) + ")");
      return FPNMT_ERR_INVALID;
    }
    W[i] = w[0];
    H[i] = h[0];
    off[i] = total;
    total += ((size_t)W[i] * H[i] * 3 + 255) / 256 * 256;
    if (sizes_out) {
      sizes_out[2 * i] = H[i];
      sizes_out[2 * i + 1] = W[i];
    }
  }
  if (total > c.rgb_cap) {
    FPNMT_CUDA_OK(cudaStreamSynchronize(s));
    if (c.rgb) cudaFree(c.rgb);
    c.rgb = nullptr;
    c.rgb_cap = 0;
    FPNMT_CUDA_OK(cudaMalloc(&c.rgb, total));
    c.rgb_cap = total;
  }
  // pass 2: decode to interleaved RGB on the device (grayscale / CMYK sources are converted by nvJPEG: channels=3), then
  // resize + normalise each image into its slot of the output batch
  for (int i = 0; i < n; ++i) {
    nvjpegImage_t img{};
    img.channel[0] = c.rgb + off[i];
    img.pitch[0] = (size_t)W[i] * 3;
    const nvjpegStatus_t st = nvjpegDecode(c.handle, c.state, jpegs[i], lengths[i], NVJPEG_OUTPUT_RGBI, &img, s);
    if (st != NVJPEG_STATUS_SUCCESS) return jpeg_fail("nvjpegDecode", (int)st);
    const int rc = launch_preprocess(c.rgb + off[i], 1, H[i], W[i], S, out + (size_t)i * S * S * 3, s);
    if (rc) return rc;
  }
  return FPNMT_OK;
}
