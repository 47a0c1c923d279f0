"""JPEG front end.  Parity chain:
  cv2.imdecode / PIL (the real libjpeg-turbo the reference calls through cv2.imread, vltk/compat.py:573-579)
    == oracle/jpeg_oracle.py (numpy + pure-Python restatement)                      [CPU, pins the oracle]
    == host entropy decoder (C++) coefficients vs the oracle's                       [CPU]
    == GPU reconstruction / GPU entropy decoder vs cv2.imdecode and the oracle       [GPU, bit-exact]
Fixtures are encoded on the fly with cv2.imencode from the seeded synthetic images (all sampling modes, qualities,
odd sizes, restart intervals, grayscale, EXIF orientation)."""
import io

import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")

SS = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
      "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}


def raw_image(h, w, seed):
    from vltk_b200 import synthetic
    return synthetic.make_raw_image(max(h, 32), max(w, 32), seed).numpy()[:h, :w].copy()


def encode(img, q=85, ss="420", rst=0, gray=False, progressive=False):
    if gray:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    args = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SS[ss]]
    if rst:
        args += [cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
    if progressive:
        args += [cv2.IMWRITE_JPEG_PROGRESSIVE, 1]
    ok, buf = cv2.imencode(".jpg", img, args)
    assert ok
    return buf.tobytes()


def cv2_decode(b):
    return cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION)


SMALL = [(48, 64, "420", 85, 0), (37, 53, "420", 50, 0), (37, 53, "422", 92, 0), (40, 40, "444", 75, 0),
         (9, 3, "420", 85, 0), (16, 16, "420", 30, 0), (64, 41, "422", 85, 3), (50, 70, "420", 95, 2),
         (33, 47, "444", 60, 5), (8, 8, "420", 85, 0), (17, 2, "422", 85, 0)]


@pytest.mark.parametrize("h,w,ss,q,rst", SMALL)
def test_oracle_is_bit_exact_with_libjpeg_turbo(h, w, ss, q, rst):
    from oracle import jpeg_oracle as J
    b = encode(raw_image(h, w, h * 100 + w), q, ss, rst)
    ref = cv2_decode(b)
    assert np.array_equal(J.decode(b), ref)
    from PIL import Image                       # the other decoder the reference uses (processing/image.py:62-70)
    pil = np.asarray(Image.open(io.BytesIO(b)).convert("RGB"))[:, :, ::-1]
    assert np.array_equal(ref, pil)


def test_oracle_grayscale():
    from oracle import jpeg_oracle as J
    b = encode(raw_image(40, 56, 3), 80, gray=True)
    assert np.array_equal(J.decode(b), cv2_decode(b))


@pytest.mark.parametrize("h,w,ss,q,rst", SMALL + [(120, 200, "420", 98, 0), (120, 200, "420", 20, 7)])
def test_host_entropy_decoder_matches_oracle(h, w, ss, q, rst):
    from oracle import jpeg_oracle as J
    from vltk_b200 import jpeg
    b = encode(raw_image(h, w, h * 100 + w), q, ss, rst)
    info, co = jpeg.coefficients(b)
    I, ref = J.coefficients(b)
    assert (info.width, info.height, info.ncomp, info.restart_interval) == (w, h, 3, rst)
    assert np.array_equal(co, np.concatenate([r.reshape(-1) for r in ref]))
    for c in range(3):
        assert np.array_equal(np.asarray(info.qt[c][:]), I["qt"][I["comps"][c]["tq"]])


def cmyk_jpeg():
    from PIL import Image
    buf = io.BytesIO()
    Image.new("CMYK", (32, 32), (10, 20, 30, 40)).save(buf, format="JPEG")
    return buf.getvalue()


@pytest.mark.parametrize("h,w,ss,q,rst", [(48, 64, "420", 85, 0), (37, 53, "422", 60, 0), (40, 40, "444", 92, 0),
                                          (9, 3, "420", 85, 0), (120, 200, "420", 30, 0), (120, 200, "420", 97, 5),
                                          (65, 129, "420", 75, 2)])
def test_progressive_host_decoder_is_bit_exact_with_libjpeg_turbo(h, w, ss, q, rst):
    """Progressive (SOF2) files: the C++ host entropy decoder walks every scan (DC/AC, first/refinement, EOB runs);
    its coefficients pushed through the oracle's IDCT/upsampling/colour stages must give cv2.imdecode's pixels."""
    from oracle import jpeg_oracle as J
    from vltk_b200 import jpeg
    b = encode(raw_image(h, w, h * 100 + w), q, ss, rst, progressive=True)
    info, co = jpeg.coefficients(b)
    assert info.progressive == 1 and (info.width, info.height) == (w, h)
    I = J.geometry(J.parse(b))
    coefs = [co[int(info.coef_offset[c]): int(info.coef_offset[c]) + info.blocks_w[c] * info.blocks_h[c] * 64]
             .reshape(info.blocks_h[c], info.blocks_w[c], 64) for c in range(info.ncomp)]
    assert np.array_equal(J.reconstruct(I, coefs), cv2_decode(b))


def test_progressive_grayscale_host_decoder():
    from oracle import jpeg_oracle as J
    from vltk_b200 import jpeg
    b = encode(raw_image(50, 70, 9), 80, gray=True, progressive=True)
    info, co = jpeg.coefficients(b)
    I = J.geometry(J.parse(b))
    assert np.array_equal(J.reconstruct(I, [co.reshape(info.blocks_h[0], info.blocks_w[0], 64)]), cv2_decode(b))


def test_unsupported_and_corrupt_files_are_reported_not_decoded():
    from vltk_b200 import _lib, jpeg
    with pytest.raises(jpeg.UnsupportedJpeg):
        jpeg.parse(cmyk_jpeg())
    with pytest.raises(_lib.LibraryError):
        jpeg.parse(b"\x89PNG\r\n\x1a\n" + b"\0" * 32)
    good = encode(raw_image(32, 32, 1))
    with pytest.raises(_lib.LibraryError):
        jpeg.coefficients(good[:200])           # truncated inside the tables


def test_exif_orientation_is_parsed():
    from vltk_b200 import jpeg
    b = encode(raw_image(24, 40, 2))
    tiff = b"II*\x00\x08\x00\x00\x00" + b"\x01\x00" + b"\x12\x01\x03\x00\x01\x00\x00\x00\x06\x00\x00\x00" + b"\x00\x00\x00\x00"
    app1 = b"Exif\x00\x00" + tiff
    seg = b"\xff\xe1" + (len(app1) + 2).to_bytes(2, "big") + app1
    b6 = b[:2] + seg + b[2:]
    assert jpeg.parse(b).orientation == 0 and jpeg.parse(b6).orientation == 6
    ref = cv2.imdecode(np.frombuffer(b6, np.uint8), cv2.IMREAD_COLOR)      # cv2.imread applies the tag
    plain = torch.from_numpy(cv2_decode(b6))
    assert np.array_equal(jpeg.apply_orientation(plain, 6).numpy(), ref)
    for o in range(2, 9):
        tiff_o = tiff.replace(b"\x06\x00\x00\x00", bytes([o, 0, 0, 0]))
        bo = b[:2] + b"\xff\xe1" + (len(app1) + 2).to_bytes(2, "big") + b"Exif\x00\x00" + tiff_o + b[2:]
        ref = cv2.imdecode(np.frombuffer(bo, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(jpeg.apply_orientation(plain, o).numpy(), ref), o


def test_decoder_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vltk_b200 import _lib, jpeg
    with pytest.raises(_lib.LibraryError):
        jpeg.JpegDecoder()


GPU_CASES = SMALL + [(600, 1000, "420", 90, 0), (600, 1000, "422", 75, 0), (601, 999, "444", 85, 0),
                     (800, 1333, "420", 92, 0), (333, 500, "420", 85, 16)]


@pytest.mark.gpu
def test_gpu_entropy_decoder_matches_host_decoder():
    """Self-synchronising parallel Huffman decoding on the device == the serial C++ decoder, coefficient for
    coefficient, on every sampling mode / size / quality (restart-interval files take the host path inside)."""
    from vltk_b200 import jpeg
    dec = jpeg.JpegDecoder(entropy="gpu")
    cases = GPU_CASES + [(600, 1000, "420", 98, 0), (600, 1000, "420", 10, 0), (1333, 800, "444", 95, 0)]
    datas = [encode(raw_image(h, w, h * 100 + w), q, ss, rst) for (h, w, ss, q, rst) in cases]
    datas.append(encode(raw_image(40, 56, 3), 80, gray=True))
    datas.append(encode(np.full((200, 300, 3), 117, np.uint8), 85))          # flat image: the shortest possible codes
    infos, offs, coef = dec.coefficients(datas)
    coef = coef.cpu().numpy()
    iters = dec.last_iterations.cpu().numpy()
    for i, b in enumerate(datas):
        _, ref = jpeg.coefficients(b)
        got = coef[offs[i]: offs[i] + ref.size]
        bad = np.nonzero(got != ref)[0]
        assert bad.size == 0, (i, cases[i] if i < len(cases) else "extra", bad[:8], got[bad[:8]], ref[bad[:8]], int(iters[i]))
    print("sync iterations per image:", iters.tolist())
    assert iters.max() <= 8      # self-synchronisation: a handful of sweeps, not one per subsequence


@pytest.mark.gpu
@pytest.mark.parametrize("entropy", ["gpu", "host"])
def test_gpu_decode_is_bit_exact_with_cv2_imread(entropy):
    from vltk_b200 import jpeg
    dec = jpeg.JpegDecoder(entropy=entropy)
    datas = [encode(raw_image(h, w, h * 100 + w), q, ss, rst) for (h, w, ss, q, rst) in GPU_CASES]
    datas.append(encode(raw_image(40, 56, 3), 80, gray=True))
    datas.append(encode(raw_image(333, 500, 77), 85, "420", progressive=True))   # host entropy decode, GPU IDCT/colour
    datas.append(encode(raw_image(61, 47, 78), 50, "444", 3, progressive=True))
    for rep in range(2):                         # second pass reuses the pinned staging buffer
        outs = dec.decode(datas)
        for b, o in zip(datas, outs):
            assert o.is_cuda and o.dtype == torch.uint8
            assert np.array_equal(o.cpu().numpy(), cv2_decode(b))


@pytest.mark.gpu
def test_jpeg_bytes_to_features_equals_cv2_decoded_path():
    """forward_jpeg_stream (JPEG bytes -> GPU entropy decode -> IDCT/colour -> fused resize/normalise/pad -> model)
    returns exactly what the reference-style path returns for the same files decoded by cv2 on the host: the decoded
    pixels are bit-identical, so every output is too.  Mixed sizes, 3 batches decoded in groups of 2."""
    from oracle import cases
    from vltk_b200.frcnn import FRCNN
    from vltk_b200.preprocess import Preprocess
    from tests.util import weights
    cfg = cases.case_config("mixed")
    model = FRCNN.from_pretrained(state_dict=weights(0), config=cfg, mode="fp32")
    model.roi_outputs.min_detections, model.roi_outputs.max_detections = cfg.min_detections, cfg.max_detections
    model.roi_outputs.nms_thresh = list(cfg.nms_thresh_test)
    pre = Preprocess(cfg)
    shapes = [[(150, 200), (240, 160)], [(180, 180), (120, 260)], [(200, 150)]]
    batches = [[encode(raw_image(h, w, 10 * bi + j), 90, "420") for j, (h, w) in enumerate(b)] for bi, b in enumerate(shapes)]
    got = list(model.forward_jpeg_stream(iter(batches), pre, group=2))
    assert len(got) == 3
    for b, g in zip(batches, got):
        raws = [torch.from_numpy(cv2_decode(d)) for d in b]
        ids, images, sizes, scales = pre(raws)
        ref = model(images, sizes, scales_yx=scales, padding="max_detections", return_tensors="np")
        for k in ("obj_ids", "attr_ids", "boxes", "roi_features", "obj_probs", "preds_per_image", "normalized_boxes"):
            assert np.array_equal(g[k], ref[k]), k
    # Preprocess itself takes encoded bytes / .jpg paths and routes them through the GPU front end
    ids, images, sizes, scales = pre(batches[0])
    ids2, images2, _, _ = pre([torch.from_numpy(cv2_decode(d)) for d in batches[0]])
    assert torch.equal(images, images2)
    from vltk_b200 import jpeg
    with pytest.raises(jpeg.UnsupportedJpeg):
        pre([cmyk_jpeg()])
    ok = Preprocess(cfg, host_decode_unsupported=True)([cmyk_jpeg()])
    assert ok[1].shape[0] == 1
    # progressive files: entropy-decoded by the C++ host decoder, everything else on the GPU as usual
    prog = encode(raw_image(150, 200, 0), 90, "420", progressive=True)
    assert torch.equal(pre([prog])[1], pre([torch.from_numpy(cv2_decode(prog))])[1])


@pytest.mark.gpu
def test_gpu_decode_randomised_sweep():
    """40 random (size, quality, sampling, restart interval) combinations, decoded in two front-end calls and
    compared bit-for-bit with cv2: exercises block-grid padding, every MCU geometry, tiny and odd images, both
    entropy paths of the device decoder (self-synchronising and per-interval)."""
    from vltk_b200 import jpeg
    rng = np.random.default_rng(20261018)
    datas = []
    for i in range(40):
        h, w = int(rng.integers(1, 260)), int(rng.integers(1, 340))
        ss = ["444", "422", "420"][int(rng.integers(0, 3))]
        q = int(rng.integers(5, 100))
        rst = int(rng.choice([0, 0, 1, 2, 5, 11]))
        datas.append(encode(raw_image(h, w, 7000 + i), q, ss, rst))
    dec = jpeg.JpegDecoder()
    for part in (datas[:23], datas[23:]):
        for b, o in zip(part, dec.decode(part)):
            ref = cv2_decode(b)
            assert o.shape == ref.shape and np.array_equal(o.cpu().numpy(), ref)


@pytest.mark.gpu
def test_preprocess_reads_jpeg_files_from_disk(tmp_path):
    """Preprocess(paths): .jpg files go through the GPU front end (EXIF orientation honoured like cv2.imread), other
    formats are read with cv2 on the host as in the reference (compat.py:573-579); both give cv2.imread's pixels."""
    from oracle import cases
    from vltk_b200.preprocess import Preprocess
    cfg = cases.case_config("mixed")
    pre = Preprocess(cfg)
    img = raw_image(150, 200, 42)
    pj, pp = str(tmp_path / "a.jpg"), str(tmp_path / "b.png")
    cv2.imwrite(pj, img, [cv2.IMWRITE_JPEG_QUALITY, 88])
    cv2.imwrite(pp, img)
    ids, images, sizes, scales = pre([pj, pp], ["a", "b"])
    ref = pre([torch.from_numpy(cv2.imread(pj)), torch.from_numpy(cv2.imread(pp))], ["a", "b"])
    assert ids == ["a", "b"] and torch.equal(images, ref[1]) and torch.equal(sizes, ref[2])
