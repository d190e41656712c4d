"""GPU parity: encoder/decoder layers through the C-ABI vs the CPU oracle.
Tolerances are on fp32-class arithmetic: 2e-5 * (1 + |ref|)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from wct_tf_b200 import _capi
from tests import gpu_util as U

pytestmark = pytest.mark.gpu


def conv_ref64(x_nhwc, w_hwio, b, relu):
    """ops.py:12-19 Conv2DReflect in float64 on the CPU."""
    x = torch.from_numpy(np.asarray(x_nhwc, dtype=np.float64)).permute(0, 3, 1, 2)
    w = torch.from_numpy(np.asarray(w_hwio, dtype=np.float64)).permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, torch.from_numpy(np.asarray(b, dtype=np.float64)))
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).numpy()


def assert_close(got, ref, tol=2e-5, name="", **dump):
    err = np.abs(got - ref) / (1.0 + np.abs(ref))
    bad = ~np.isfinite(got)
    if bad.any() or err.max() > tol:
        U.dump("fail_" + name, got=got, ref=ref, **dump)
        idx = np.unravel_index(np.nanargmax(np.where(bad, np.inf, err)), err.shape)
        raise AssertionError("%s: max rel err %.3e at %s (got %r ref %r), non-finite %d" %
                             (name, np.nanmax(err), idx, got[idx], ref[idx], int(bad.sum())))


def test_image_pre_post():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (2, 5, 7, 3), dtype=np.uint8)
    d = U.dev(img)
    f = torch.empty(img.shape, dtype=torch.float32, device="cuda")
    _capi.check(U.lib().wctb200_image_u8_to_f32(d.data_ptr(), d.numel(), f.data_ptr(), U.stream()))
    ref = (img / 255.0).astype(np.float32)          # wct.py:64 (float64 divide, fed to an fp32 placeholder)
    assert np.abs(f.cpu().numpy() - ref).max() <= 6e-8
    x = rng.normal(0.5, 0.6, (3, 9, 4, 3)).astype(np.float32)
    xd = U.dev(x)
    o = torch.empty(x.shape, dtype=torch.uint8, device="cuda")
    _capi.check(U.lib().wctb200_image_f32_to_u8(xd.data_ptr(), xd.numel(), o.data_ptr(), U.stream()))
    ref8 = np.uint8(np.clip(x, 0, 1) * np.float32(255))   # wct.py:68, truncation
    assert np.array_equal(o.cpu().numpy(), ref8)


@pytest.mark.parametrize("shape", [(1, 2, 2, 8), (2, 5, 7, 16), (1, 8, 3, 64)])
def test_act_roundtrip_and_reflect_halo(shape):
    rng = np.random.default_rng(1)
    x = rng.normal(0, 3, shape).astype(np.float32)
    buf = U.act_from_numpy(x)
    n, h, w, c = shape
    back = U.act_to_numpy(buf, n, h, w, c)
    assert np.abs(back - x).max() <= 1e-6 * (1 + np.abs(x).max())
    padded = U.act_raw_padded(buf, n, h, w, c)
    ref = np.pad(U.split_repr(x), ((0, 0), (1, 1), (1, 1), (0, 0)), mode="reflect")   # tf.pad REFLECT, ops.py:12-15
    assert np.array_equal(padded, ref)


@pytest.mark.parametrize("impl", ["tensor-core", "simt"])
@pytest.mark.parametrize("shape", [(1, 6, 5), (2, 16, 12), (1, 33, 20), (3, 37, 53), (1, 128, 130), (2, 2, 2)])
def test_conv_head_matches_preprocess_plus_conv1_1(shape, impl):
    """the head runs on the tensor cores with an operand built in shared memory (conv_head_tc.cu); the SIMT kernel stays
    covered through the debug knob"""
    if impl == "simt":
        U.lib().wctb200_debug_set_conv_head_tc(0)
    try:
        _conv_head_case(shape)
    finally:
        U.lib().wctb200_debug_set_conv_head_tc(1)


def _conv_head_case(shape):
    from wct_tf_b200.weights import make_synthetic_weights
    n, h, w = shape
    wts = make_synthetic_weights(3)
    vgg = {l["name"]: l for l in wts["vgg"]}
    rng = np.random.default_rng(2)
    img = rng.random((n, h, w, 3)).astype(np.float32)
    w0 = vgg["preprocess"]["weight"].astype(np.float64)[:, :, 0, 0]
    b0 = vgg["preprocess"]["bias"].astype(np.float64)
    w1 = vgg["conv1_1"]["weight"].astype(np.float64)
    b1 = vgg["conv1_1"]["bias"].astype(np.float64)
    wf = np.einsum("ojyx,ji->yxio", w1, w0).reshape(27, 64).astype(np.float32)
    bf = (b1 + np.einsum("ojyx,j->o", w1, b0)).astype(np.float32)
    out = U.act_alloc(n, h, w, 64)
    d_img, d_wf, d_bf = U.dev(img), U.dev(wf), U.dev(bf)      # keep the device tensors alive across the async call
    _capi.check(U.lib().wctb200_conv_head(d_img.data_ptr(), n, h, w, d_wf.data_ptr(), d_bf.data_ptr(),
                                          out.data_ptr(), U.stream()))
    got = U.act_to_numpy(out, n, h, w, 64)
    # reference order of operations: conv0 (1x1), reflect pad, conv1_1, relu (vgg_normalised.py:25-40)
    y0 = np.einsum("nhwi,ji->nhwj", img.astype(np.float64), w0) + b0
    ref = conv_ref64(y0, np.transpose(w1, (2, 3, 1, 0)), b1, True)
    assert_close(got, ref, name="conv_head")
    padded = U.act_raw_padded(out, n, h, w, 64)
    assert np.isfinite(padded).all()
    assert np.array_equal(padded, np.pad(padded[:, 1:-1, 1:-1], ((0, 0), (1, 1), (1, 1), (0, 0)), mode="reflect"))


CONV_CASES = [
    # N, H, W, Cin, Cout, relu
    (1, 8, 8, 64, 64, True),
    (1, 5, 9, 64, 128, True),
    (2, 7, 6, 128, 64, False),
    (1, 16, 16, 128, 128, True),
    (1, 12, 10, 256, 256, True),
    (2, 9, 11, 256, 512, True),
    (1, 8, 8, 512, 512, True),
    (3, 34, 30, 64, 64, True),
    (1, 64, 64, 64, 128, True),        # flat tiling, two TMA boxes per plane
    (1, 8, 130, 64, 64, True),         # 2-D tiles (4 x 30), ragged right edge
    (2, 9, 128, 128, 128, True),       # 2-D tiles, ragged bottom edge, batch
    (1, 37, 260, 64, 64, False),
    (1, 13, 40, 64, 64, True),
    (3, 16, 48, 64, 128, True),        # two cout tiles at BN=64
    (1, 24, 16, 192, 64, True),        # three 64-channel slices
]


def _conv_inputs(case, seed):
    n, h, w, cin, cout, relu = case
    rng = np.random.default_rng(seed)
    x = np.maximum(rng.normal(0.3, 1.0, (n, h, w, cin)), 0).astype(np.float32)
    k = (rng.normal(0, 1, (3, 3, cin, cout)) * np.sqrt(2.0 / (9 * cin))).astype(np.float32)
    b = rng.normal(0, 0.3, cout).astype(np.float32)
    return x, k, b


@pytest.mark.parametrize("case", CONV_CASES[:4], ids=lambda c: "x".join(map(str, c)))
def test_conv3x3_ref_kernel(case):
    n, h, w, cin, cout, relu = case
    x, k, b = _conv_inputs(case, 5)
    xin = U.act_from_numpy(x)
    out = U.act_alloc(n, h, w, cout)
    d_k, d_b = U.dev(k), U.dev(b)
    _capi.check(U.lib().wctb200_conv3x3_ref(xin.data_ptr(), n, h, w, cin, d_k.data_ptr(), d_b.data_ptr(), cout,
                                            _capi.RELU if relu else 0, out.data_ptr(), U.stream()))
    got = U.act_to_numpy(out, n, h, w, cout)
    ref = conv_ref64(U.split_repr(x), k, b, relu)
    assert_close(got, ref, name="conv_ref_%d_%d" % (cin, cout))


def weight_repr(k):
    """The value the library stores for fp32 weights k: split-fp16 pair of k * 2^S, S such that max|k| * 2^S is in
    [512, 1024) (csrc/layers.cu: k_prep_weights), divided by 2^S again."""
    k = np.asarray(k, dtype=np.float32)
    amax = float(np.abs(k).max())
    if amax <= 0:
        return k.astype(np.float64)
    _, e = np.frexp(np.float32(amax))
    sc = 2.0 ** int(np.clip(10 - int(e), -14, 40))
    return U.split_repr(k * np.float32(sc)) / sc


@pytest.mark.parametrize("fuse", [-1, 0, 1])     # fused [b_hi|b_lo] MMA: auto / off / forced on
@pytest.mark.parametrize("bn", [0, 64, 256])
@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv3x3_tensor_core(case, bn, fuse):
    n, h, w, cin, cout, relu = case
    if bn and cout % bn:
        pytest.skip("tile does not divide Cout")
    x, k, b = _conv_inputs(case, 7)
    xin = U.act_from_numpy(x)
    wsplit = torch.empty(U.lib().wctb200_conv_weight_bytes(9, cin, cout), dtype=torch.uint8, device="cuda")
    d_k, d_b = U.dev(k), U.dev(b)
    _capi.check(U.lib().wctb200_prep_conv_weights(d_k.data_ptr(), 9, cin, cout, wsplit.data_ptr(), U.stream()))
    out = U.act_alloc(n, h, w, cout)
    U.lib().wctb200_debug_set_conv_bn(bn)
    U.lib().wctb200_debug_set_conv_fuse(fuse)
    try:
        _capi.check(U.lib().wctb200_conv3x3(xin.data_ptr(), n, h, w, cin, wsplit.data_ptr(), d_b.data_ptr(), cout,
                                            _capi.RELU if relu else 0, out.data_ptr(), U.stream()))
        U.check_device()
    finally:
        U.lib().wctb200_debug_set_conv_bn(0)
        U.lib().wctb200_debug_set_conv_fuse(-1)
    got = U.act_to_numpy(out, n, h, w, cout)
    ref = conv_ref64(U.split_repr(x), weight_repr(k), b, relu)
    err = np.abs(got - ref) / (1.0 + np.abs(ref))
    print("K=%d: max rel err %.2e, mean signed err %.2e" % (9 * cin, err.max(), (got - ref).mean()))
    assert_close(got, ref, tol=1e-5, name="conv_tc_%d_%d_bn%d" % (cin, cout, bn), x=x, k=k, b=b)
    # the stored weights carry >= 21 bits: the conv sits within 4e-6 of the conv with the EXACT fp32 weights
    assert_close(got, conv_ref64(U.split_repr(x), k, b, relu), tol=4e-6, name="conv_tc_exactw_%d_%d" % (cin, cout))
    padded = U.act_raw_padded(out, n, h, w, cout)
    assert np.isfinite(padded).all(), "halo cells left unwritten"
    assert np.array_equal(padded, np.pad(padded[:, 1:-1, 1:-1], ((0, 0), (1, 1), (1, 1), (0, 0)), mode="reflect"))


UP2_CASES = [
    # N, H, W (low resolution), Cin, Cout, relu
    (1, 4, 4, 64, 64, True),
    (2, 5, 7, 128, 64, True),
    (1, 8, 6, 128, 128, False),
    (1, 2, 2, 64, 64, True),           # smallest map: every pixel touches the border
    (2, 9, 17, 256, 128, True),
    (1, 16, 16, 512, 512, True),
    (1, 33, 20, 64, 64, True),
]


@pytest.mark.parametrize("bn", [0, 64])
@pytest.mark.parametrize("case", UP2_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv3x3_up2_equals_upsample_then_conv(case, bn):
    """wctb200_conv3x3_up2 (UpSampling2D folded into the conv: 4 parity kernels with pre-summed taps over the
    low-resolution input with an EDGE halo) == UpSampling2D -> Conv2DReflect of model.py:291-293 in float64."""
    n, h, w, cin, cout, relu = case
    if bn and cout % bn:
        pytest.skip("tile does not divide Cout")
    rng = np.random.default_rng(11)
    # the low-resolution input is itself produced by a conv launched with HALO_EDGE (as the engine does)
    x0, k0, b0 = _conv_inputs((n, h, w, cin, cin, True), 12)
    k = (rng.normal(0, 1, (3, 3, cin, cout)) * np.sqrt(2.0 / (9 * cin))).astype(np.float32)
    b = rng.normal(0, 0.3, cout).astype(np.float32)
    lib = U.lib()
    x0in = U.act_from_numpy(x0)
    w0 = torch.empty(lib.wctb200_conv_weight_bytes(9, cin, cin), dtype=torch.uint8, device="cuda")
    d_k0, d_b0, d_k, d_b = U.dev(k0), U.dev(b0), U.dev(k), U.dev(b)
    _capi.check(lib.wctb200_prep_conv_weights(d_k0.data_ptr(), 9, cin, cin, w0.data_ptr(), U.stream()))
    low = U.act_alloc(n, h, w, cin)
    _capi.check(lib.wctb200_conv3x3(x0in.data_ptr(), n, h, w, cin, w0.data_ptr(), d_b0.data_ptr(), cin,
                                    _capi.RELU | _capi.HALO_EDGE, low.data_ptr(), U.stream()))
    lowp = U.act_raw_padded(low, n, h, w, cin)
    assert np.isfinite(lowp).all()
    assert np.array_equal(lowp, np.pad(lowp[:, 1:-1, 1:-1], ((0, 0), (1, 1), (1, 1), (0, 0)), mode="edge")), "HALO_EDGE"
    wup = torch.empty(lib.wctb200_conv_weight_bytes(16, cin, cout), dtype=torch.uint8, device="cuda")
    _capi.check(lib.wctb200_prep_conv_weights_up2(d_k.data_ptr(), cin, cout, wup.data_ptr(), U.stream()))
    out = U.act_alloc(n, 2 * h, 2 * w, cout)
    lib.wctb200_debug_set_conv_bn(bn)
    try:
        _capi.check(lib.wctb200_conv3x3_up2(low.data_ptr(), n, h, w, cin, wup.data_ptr(), d_b.data_ptr(), cout,
                                            _capi.RELU if relu else 0, out.data_ptr(), U.stream()))
        U.check_device()
    finally:
        lib.wctb200_debug_set_conv_bn(0)
    got = U.act_to_numpy(out, n, 2 * h, 2 * w, cout)
    xl = lowp[:, 1:-1, 1:-1]                                            # what the device holds (float64 of hi+lo)
    up = np.repeat(np.repeat(xl, 2, axis=1), 2, axis=2)                 # UpSampling2D, model.py:293
    ref = conv_ref64(up, k, b, relu)
    assert_close(got, ref, tol=1e-5, name="conv_up2_%d_%d" % (cin, cout))
    padded = U.act_raw_padded(out, n, 2 * h, 2 * w, cout)
    assert np.isfinite(padded).all(), "halo cells left unwritten"
    assert np.array_equal(padded, np.pad(padded[:, 1:-1, 1:-1], ((0, 0), (1, 1), (1, 1), (0, 0)), mode="reflect"))


@pytest.mark.parametrize("shape", [(1, 8, 8, 64), (2, 7, 5, 64), (1, 9, 12, 128), (1, 2, 3, 8)])
def test_maxpool_same_and_upsample(shape):
    n, h, w, c = shape
    rng = np.random.default_rng(9)
    x = rng.normal(0, 2, shape).astype(np.float32)
    xin = U.act_from_numpy(x)
    xs = U.split_repr(x)
    ho, wo = (h + 1) // 2, (w + 1) // 2
    if ho >= 2 and wo >= 2:
        out = U.act_alloc(n, ho, wo, c)
        _capi.check(U.lib().wctb200_maxpool2(xin.data_ptr(), n, h, w, c, out.data_ptr(), U.stream()))
        t = torch.from_numpy(xs).permute(0, 3, 1, 2)
        ref = F.max_pool2d(t, 2, 2, ceil_mode=True).permute(0, 2, 3, 1).numpy()   # MaxPooling2D(padding='same')
        padded = U.act_raw_padded(out, n, ho, wo, c)
        assert np.array_equal(padded[:, 1:-1, 1:-1], ref)
        assert np.array_equal(padded, np.pad(ref, ((0, 0), (1, 1), (1, 1), (0, 0)), mode="reflect"))
    out = U.act_alloc(n, 2 * h, 2 * w, c)
    _capi.check(U.lib().wctb200_upsample2(xin.data_ptr(), n, h, w, c, out.data_ptr(), U.stream()))
    ref = xs.repeat(2, axis=1).repeat(2, axis=2)                                    # UpSampling2D, model.py:293
    padded = U.act_raw_padded(out, n, 2 * h, 2 * w, c)
    assert np.array_equal(padded, np.pad(ref, ((0, 0), (1, 1), (1, 1), (0, 0)), mode="reflect"))


@pytest.mark.parametrize("impl", ["tensor-core", "simt"])
@pytest.mark.parametrize("clip", [False, True])
@pytest.mark.parametrize("shape", [(1, 8, 8, 64), (2, 5, 7, 64), (1, 33, 3, 64),
                                   (1, 32, 32, 64), (2, 37, 45, 64), (1, 64, 96, 64), (1, 40, 33, 128),   # >= 32x32: tiled SIMT kernel
                                   (1, 70, 300, 64), (2, 65, 127, 64), (1, 31, 253, 64)])                 # several strips / row chunks
def test_conv_tail(shape, clip, impl):
    """64-channel tails run as the transposed tensor-core product (conv_tail_tc.cu); the SIMT kernels serve other widths and
    stay covered through the debug knob."""
    n, h, w, c = shape
    if impl == "simt":
        U.lib().wctb200_debug_set_conv_tail_tc(0)
    try:
        _conv_tail_case(n, h, w, c, clip)
    finally:
        U.lib().wctb200_debug_set_conv_tail_tc(1)


def _conv_tail_case(n, h, w, c, clip):
    shape = (n, h, w, c)
    rng = np.random.default_rng(11)
    x = np.maximum(rng.normal(0.3, 1.0, shape), 0).astype(np.float32)
    k = (rng.normal(0, 1, (3, 3, c, 3)) * 0.05).astype(np.float32)
    b = np.array([0.5, 0.4, 0.6], dtype=np.float32)
    xin = U.act_from_numpy(x)
    img = torch.full((n, h, w, 3), float("nan"), dtype=torch.float32, device="cuda")
    d_k, d_b = U.dev(k.reshape(9 * c, 3)), U.dev(b)
    _capi.check(U.lib().wctb200_conv_tail(xin.data_ptr(), n, h, w, c, d_k.data_ptr(), d_b.data_ptr(),
                                          _capi.CLIP01 if clip else 0, img.data_ptr(), U.stream()))
    ref = conv_ref64(U.split_repr(x), k, b, False)
    if clip:
        ref = np.clip(ref, 0, 1)
    assert_close(img.cpu().numpy(), ref, name="conv_tail")


POOL_CASES = [
    # N, H, W, Cin, Cout
    (1, 8, 8, 64, 64),
    (2, 6, 10, 64, 64),
    (1, 7, 9, 64, 64),          # odd sizes: 'same' pooling keeps the ragged last row / column (ceil)
    (1, 4, 4, 128, 128),
    (2, 9, 70, 128, 128),       # two column segments, the second ragged
    (1, 33, 130, 64, 64),       # three segments, odd height
    (1, 16, 16, 256, 256),
    (1, 12, 20, 512, 512),
    (3, 5, 5, 64, 128),
]


@pytest.mark.parametrize("bn", [0, 64])
@pytest.mark.parametrize("case", POOL_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv3x3_pool2_equals_conv_then_maxpool(case, bn):
    """WCTB200_POOL2 (MaxPooling2D 2x2/2 'same' of vgg_normalised.py:41-42 folded into the conv epilogue) gives the SAME bits as
    wctb200_conv3x3 -> wctb200_maxpool2 (max commutes with the monotone scale/bias/ReLU/split), reflect halo of the pooled map
    included, and matches the float64 reference."""
    n, h, w, cin, cout = case
    if bn and cout % bn:
        pytest.skip("tile does not divide Cout")
    x, k, b = _conv_inputs((n, h, w, cin, cout, True), 21)
    lib = U.lib()
    xin = U.act_from_numpy(x)
    wsp = torch.empty(lib.wctb200_conv_weight_bytes(9, cin, cout), dtype=torch.uint8, device="cuda")
    d_k, d_b = U.dev(k), U.dev(b)
    _capi.check(lib.wctb200_prep_conv_weights(d_k.data_ptr(), 9, cin, cout, wsp.data_ptr(), U.stream()))
    ho, wo = (h + 1) // 2, (w + 1) // 2
    full = U.act_alloc(n, h, w, cout)
    two_step = U.act_alloc(n, ho, wo, cout)
    fused = U.act_alloc(n, ho, wo, cout)
    lib.wctb200_debug_set_conv_bn(bn)
    try:
        _capi.check(lib.wctb200_conv3x3(xin.data_ptr(), n, h, w, cin, wsp.data_ptr(), d_b.data_ptr(), cout, _capi.RELU,
                                        full.data_ptr(), U.stream()))
        _capi.check(lib.wctb200_maxpool2(full.data_ptr(), n, h, w, cout, two_step.data_ptr(), U.stream()))
        _capi.check(lib.wctb200_conv3x3(xin.data_ptr(), n, h, w, cin, wsp.data_ptr(), d_b.data_ptr(), cout,
                                        _capi.RELU | _capi.POOL2, fused.data_ptr(), U.stream()))
        U.check_device()
    finally:
        lib.wctb200_debug_set_conv_bn(0)
    a = U.act_raw_padded(fused, n, ho, wo, cout)
    assert np.isfinite(a).all(), "pooled cells left unwritten"
    assert np.array_equal(a, U.act_raw_padded(two_step, n, ho, wo, cout))
    assert np.array_equal(a, np.pad(a[:, 1:-1, 1:-1], ((0, 0), (1, 1), (1, 1), (0, 0)), mode="reflect"))
    ref = conv_ref64(U.split_repr(x), k, b, True)
    t = torch.from_numpy(ref).permute(0, 3, 1, 2)
    ref = F.max_pool2d(t, 2, 2, ceil_mode=True).permute(0, 2, 3, 1).numpy()
    assert_close(a[:, 1:-1, 1:-1], ref, tol=1e-5, name="conv_pool_%d_%d" % (cin, cout))


@pytest.mark.parametrize("case", [(4, 200, 300, 64, 64), (2, 130, 262, 128, 128), (1, 96, 96, 512, 512)], ids=lambda c: "x".join(map(str, c)))
def test_conv3x3_pool2_many_tiles_per_cta(case):
    """More tiles than CTAs (the persistent loop, the row-pair exchange buffer and the accumulator ring are reused back to back):
    POOL2 == conv -> maxpool by value on every cell, halo included."""
    n, h, w, cin, cout = case
    x, k, b = _conv_inputs((n, h, w, cin, cout, True), 31)
    lib = U.lib()
    xin = U.act_from_numpy(x)
    wsp = torch.empty(lib.wctb200_conv_weight_bytes(9, cin, cout), dtype=torch.uint8, device="cuda")
    d_k, d_b = U.dev(k), U.dev(b)
    _capi.check(lib.wctb200_prep_conv_weights(d_k.data_ptr(), 9, cin, cout, wsp.data_ptr(), U.stream()))
    ho, wo = (h + 1) // 2, (w + 1) // 2
    full, two_step, fused = U.act_alloc(n, h, w, cout), U.act_alloc(n, ho, wo, cout), U.act_alloc(n, ho, wo, cout)
    _capi.check(lib.wctb200_conv3x3(xin.data_ptr(), n, h, w, cin, wsp.data_ptr(), d_b.data_ptr(), cout, _capi.RELU, full.data_ptr(), U.stream()))
    _capi.check(lib.wctb200_maxpool2(full.data_ptr(), n, h, w, cout, two_step.data_ptr(), U.stream()))
    _capi.check(lib.wctb200_conv3x3(xin.data_ptr(), n, h, w, cin, wsp.data_ptr(), d_b.data_ptr(), cout, _capi.RELU | _capi.POOL2,
                                    fused.data_ptr(), U.stream()))
    U.check_device()
    a = U.act_raw_padded(fused, n, ho, wo, cout)
    assert np.isfinite(a).all()
    assert np.array_equal(a, U.act_raw_padded(two_step, n, ho, wo, cout))
