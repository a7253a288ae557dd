// CUtensorMap construction through the driver entry point (no link-time dependency on libcuda).
#include "tensormap.cuh"
#include <stdlib.h>

#include <cudaTypedefs.h>

#include <mutex>

namespace fpnmt {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
  auto fn = get_encode();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return 3;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "cuTensorMapEncodeTiled failed (%d): rank %d base %p dims [%llu %llu %llu %llu] strides [%llu %llu %llu] "
             "box [%u %u %u %u]",
             (int)r, rank, base, (unsigned long long)dims[0], (unsigned long long)dims[1],
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
             (unsigned long long)strides_bytes[0], (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
             (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0,
             rank > 3 ? box[3] : 0);
    set_last_error(buf);
    return 3;
  }
  return 0;
}

// NHWC activation viewed as a rank-4 tensor (C, W, H, N); box = (64 channels, tw, th, bn) = one 128 x 64 A tile.
int encode_tmap_act(CUtensorMap* m, const bf16* base, int C, int ld, int W, int H, int N, int tw, int th, int bn) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)IG_BK, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)bn};
  return encode(m, base, 4, dims, strides, box);
}

// Row-major [outer][ld] matrix, logical width `inner`; box = (64, box_outer).
int encode_tmap_2d(CUtensorMap* m, const bf16* base, uint64_t inner, uint64_t outer, uint64_t ld_elems, int box_outer) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)IG_BK, (cuuint32_t)box_outer};
  return encode(m, base, 2, dims, strides, box);
}

static int np2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

int make_igemm_op(IgemmOp* op, const ConvGeom& g, const Act& in, const bf16* wt, bool split, const float* bias, int act,
                  const Act& out, float* out_f32, int ld_f32, int res_mode, const Act& res, int num_sms, int force_bn, bool tma_store,
                  bool b_stationary) {
  IgemmParams& p = op->p;
  p = IgemmParams{};
  p.N = g.N; p.H = g.H; p.W = g.W;

  if (g.H == 1 && g.N == 1) {           // dense: rows along W
    p.tw = 128; p.th = 1; p.bn = 1;
  } else {
    p.tw = np2(g.W) < 16 ? np2(g.W) : 16;
    int th = 128 / p.tw;
    p.th = np2(g.H) < th ? np2(g.H) : th;
    p.bn = 128 / (p.tw * p.th);
  }
  p.tiles_x = (g.W + p.tw - 1) / p.tw;
  p.tiles_y = (g.H + p.th - 1) / p.th;
  p.tiles_n = (g.N + p.bn - 1) / p.bn;
  p.taps_y = g.kh; p.taps_x = g.kw; p.pad_y = g.pad_y; p.pad_x = g.pad_x;
  p.Cin = g.Cin;
  p.kchunks = (g.Cin + IG_BK - 1) / IG_BK;
  p.nterms = split ? 3 : 1;
  const int ktot = g.kh * g.kw * g.Cin;
  p.b_lo_off = split ? ktot : 0;
  p.Cout = g.Cout;
  p.bias = bias;
  p.act = act;
  p.out = out;
  p.out_f32 = out_f32;
  p.ld_f32 = ld_f32;
  p.res_mode = res_mode;
  p.res = res;

  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_n;
  int bn_sel = force_bn;
  if (!bn_sel) {
    const int cands[4] = {256, 128, 64, 32};
    int cap = np2(g.Cout) < 32 ? 32 : np2(g.Cout);
    bn_sel = 32;
    for (int i = 0; i < 4; ++i) {
      const int bn = cands[i];
      if (bn > cap) continue;
      const long tiles = (long)tiles_m * ((g.Cout + bn - 1) / bn);
      if (tiles >= num_sms || bn == 32) {
        bn_sel = bn;
        break;
      }
    }
  }
  op->BN = bn_sel;
  p.tiles_co = (g.Cout + bn_sel - 1) / bn_sel;
  const long total = (long)tiles_m * p.tiles_co;
  op->grid = (int)(total < num_sms ? total : num_sms);
  op->flops = 2.0 * (double)g.N * g.H * g.W * (double)g.Cout * (double)ktot;

  if (in.C != g.Cin) {
    set_last_error("make_igemm_op: input view channel count != Cin");
    return 1;
  }
  if (split && !in.lo) {
    set_last_error("make_igemm_op: split mode needs a split input view");
    return 1;
  }
  int rc = encode_tmap_act(&op->tmA_hi, in.p, in.C, in.ld, g.W, g.H, g.N, p.tw, p.th, p.bn);
  if (rc) return rc;
  rc = encode_tmap_act(&op->tmA_lo, split ? in.p + in.lo : in.p, in.C, in.ld, g.W, g.H, g.N, p.tw, p.th, p.bn);
  if (rc) return rc;
  // Output map for the bulk-tensor-store epilogue (kept in the tmA_lo slot, which bf16 mode does not use): an epilogue warp owns
  // 32 consecutive rows of the 128-row tile = a (bx, by, bz) sub-box of the (tw, th, bn) pixel box, 64 channels per store.
  p.tma_store = 0;
  if (!split && out.p && !out.lo && !out_f32 && bn_sel >= 128 && (g.Cout % 8) == 0 && tma_store) {
    const int bx = p.tw < 32 ? p.tw : 32;
    const int by = p.th < 32 / bx ? p.th : 32 / bx;
    const int bz = 32 / (bx * by);
    cuuint64_t dims[4] = {(cuuint64_t)g.Cout, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)out.ld * 2, (cuuint64_t)g.W * out.ld * 2, (cuuint64_t)g.H * g.W * out.ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)IG_BK, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
    rc = encode(&op->tmA_lo, out.p, 4, dims, strides, box);
    if (rc) return rc;
    p.tma_store = 1;
  }
  // Stationary weight panel: one output-channel tile, bf16 mode, at least two tiles per CTA, panel <= 96 KB and the launch fits
  p.b_stat = 0;
  {
    const int kit = p.taps_y * p.taps_x * p.kchunks;
    const size_t need = igemm_smem_bytes(bn_sel, kit);
    if (!split && p.tiles_co == 1 && total >= 2L * op->grid && (size_t)kit * bn_sel * IG_BK * 2 <= 96 * 1024 && need <= 226 * 1024 &&
        b_stationary)
      p.b_stat = 1;
    op->smem_bytes = (int)(p.b_stat ? need : igemm_smem_bytes(bn_sel));
  }
  const uint64_t kw_total = split ? 2 * (uint64_t)ktot : (uint64_t)ktot;
  return encode_tmap_2d(&op->tmB, wt, kw_total, (uint64_t)g.Cout, kw_total, bn_sel);
}

}  // namespace fpnmt
