"""GPU JPEG decode feeding the resize / normalise kernel: fpnmt_op_decode_jpeg == dataset.load_image
(/root/reference/dataset.py:19-26) on encoded files, checked against a host decode of the SAME bytes (PIL / libjpeg-turbo, the
decoder family TensorFlow's decode_jpeg uses) followed by the golden-pinned host mirror of resize + preprocess_input.

Tolerance (stated): JPEG decoders are not bit-identical - IDCT rounding and chroma upsampling differ by a few grey levels.  In
grey levels of the decoded image (1 level = 1/127.5 of the network's input range), measured on B200 (nvJPEG 12.9 GPU backend vs
libjpeg-turbo ISLOW + fancy upsampling): 4:4:4 files max 2.55-2.75 / mean 0.49-0.50 (IDCT arithmetic), bound 4 / 0.75; a smooth
4:2:0 file max 3.0 / mean 0.52, bound 4 / 0.75; a 4:2:0 file of pure chroma NOISE max 79 / mean 9.0 (nvJPEG replicates chroma
samples, libjpeg interpolates them: the worst case for that difference, kept as a stated bound of 100 / 12); grayscale max
0.94 / mean 0.02, bound 2 / 0.1.  Bit-exact libjpeg parity would need our own entropy decoder + ISLOW IDCT; not claimed."""
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _jpeg(arr, **kw):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, "JPEG", **kw)
    return buf.getvalue()


def _host_reference(data, size):
    from PIL import Image
    from fpnmt.dataset import preprocess_input, resize_bilinear_tf2
    arr = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"), dtype=np.float32)
    return preprocess_input(resize_bilinear_tf2(arr, size, size))


def test_decode_jpeg_batch_matches_host_decode():
    from fpnmt.engine import decode_jpeg
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:300, 0:420]
    smooth = np.stack([(yy * 255 / 300), (xx * 255 / 420), ((yy + xx) * 255 / 720)], -1).astype(np.uint8)
    noise = rng.integers(0, 256, (97, 64, 3), dtype=np.uint8)
    gray = (rng.integers(0, 256, (128, 160), dtype=np.uint8))
    files = [_jpeg(smooth, quality=95, subsampling=0), _jpeg(smooth, quality=90), _jpeg(noise, quality=95, subsampling=0),
             _jpeg(noise, quality=85), _jpeg(gray, quality=95)]
    out, sizes = decode_jpeg(files, size=256, return_sizes=True)
    assert tuple(out.shape) == (5, 256, 256, 3) and out.dtype == torch.float32 and out.is_cuda
    assert sizes.tolist() == [[300, 420], [300, 420], [97, 64], [97, 64], [128, 160]]
    lvl = 1.0 / 127.5
    bounds = [(4 * lvl, 0.75 * lvl), (4 * lvl, 0.75 * lvl), (4 * lvl, 0.75 * lvl), (100 * lvl, 12 * lvl), (2 * lvl, 0.1 * lvl)]
    errs = []
    for i, data in enumerate(files):
        ref = torch.from_numpy(_host_reference(data, 256))
        err = (out[i].cpu() - ref).abs()
        errs.append((float(err.max()) / lvl, float(err.mean()) / lvl))
    print("jpeg decode error vs PIL, grey levels (max, mean):", [(round(a, 2), round(b, 3)) for a, b in errs])
    for i, (mx, mean) in enumerate(errs):
        assert mx * lvl <= bounds[i][0] and mean * lvl <= bounds[i][1], (i, errs)
    assert float(out.min()) >= -1.0 and float(out.max()) <= 1.0


def test_decode_jpeg_feeds_the_engine_and_rejects_garbage():
    import fpnmt_oracle as O
    from fpnmt._lib import FpnmtError
    from fpnmt.engine import Engine, decode_jpeg
    imgs = ((O.test_images(2, 256, seed=5).numpy() + 1.0) * 127.5).round().clip(0, 255).astype(np.uint8)
    files = [_jpeg(a, quality=98, subsampling=0) for a in imgs]
    x = decode_jpeg(files, size=256)
    w = O.test_weights("mobilenet224_1.0", vocab=512, layers=2, seed=0)
    eng = Engine(w, backbone="mobilenet224_1.0", batch=2, beam=4, vocab=512, max_len=8, num_layers=2, image_size=256, precision="bf16x3")
    ref_in = torch.stack([torch.from_numpy(_host_reference(f, 256)) for f in files])
    assert float((x.cpu() - ref_in).abs().max()) <= 4 / 127.5       # the engine's input: same bound as the 4:4:4 files above
    ids, lens = eng.generate(x)                                     # device tensor straight into the path, no host round trip
    eng.close()
    assert tuple(ids.shape)[0] == 2 and int(lens.min()) >= 1
    with pytest.raises(FpnmtError) as ei:
        decode_jpeg([files[0], b"this is not a jpeg"], size=256)
    assert "image 1" in str(ei.value)
