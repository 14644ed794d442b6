"""tcgen05 implicit-GEMM convolution vs a plain PyTorch fp32 reference of the same op (on the same
bf16-rounded operands), plus the SIMT cross-check kernel.  Shapes cover every conv class of the two backbones
and the deconv head."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf16_round(t):
    return t.to(torch.bfloat16).float()


def _nhwc(x_nchw, cpad=None):
    B, C, H, W = x_nchw.shape
    cpad = cpad or C
    out = torch.zeros(B, H, W, cpad, dtype=torch.bfloat16, device=x_nchw.device)
    out[..., :C] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out.contiguous()


def _report(name, got, ref, tol_rel=2.0 ** -7, tol_abs=2e-2):
    got = got.float()
    err = (got - ref).abs()
    lim = tol_abs + tol_rel * ref.abs()
    bad = err > lim
    nbad = int(bad.sum())
    if nbad:
        idx = bad.nonzero()[:8].tolist()
        rows_bad = bad.reshape(bad.shape[0], -1, bad.shape[-1]).any(-1).sum().item()
        chans_bad = bad.reshape(-1, bad.shape[-1]).any(0).nonzero().flatten()[:32].tolist()
        msg = (f"{name}: {nbad}/{bad.numel()} mismatches, max err {err.max().item():.4g}, "
               f"ref absmax {ref.abs().max().item():.4g}, got absmax {got.abs().max().item():.4g}, "
               f"bad pixels {rows_bad}, first bad idx {idx}, bad channels {chans_bad}")
        vals = [(got[tuple(i)].item(), ref[tuple(i)].item()) for i in idx[:4]]
        raise AssertionError(msg + f" samples(got,ref)={vals}")
    return err.max().item()


CASES = [
    # name, B, Cin, H, W, Cout, k, stride, pad
    ("gemm_1x1_64", 2, 64, 64, 64, 64, 1, 1, 0),
    ("gemm_1x1_256_to_64", 2, 256, 64, 64, 64, 1, 1, 0),
    ("gemm_1x1_64_to_256", 1, 64, 64, 64, 256, 1, 1, 0),
    ("conv3_64_16x16", 3, 64, 16, 16, 64, 3, 1, 1),
    ("conv3_32_64x64_sw64", 2, 32, 64, 64, 32, 3, 1, 1),
    ("conv3_128_16x16", 2, 128, 16, 16, 128, 3, 1, 1),
    ("conv3_256_8x8_bn2", 3, 256, 8, 8, 256, 3, 1, 1),
    ("conv3_s2_64_to_128", 2, 64, 32, 32, 128, 3, 2, 1),
    ("conv3_s2_32_to_64", 2, 32, 64, 64, 64, 3, 2, 1),
    ("conv1_s2_256_to_512", 2, 256, 32, 32, 512, 1, 2, 0),
    ("final_1x1_256_to_448", 1, 256, 64, 64, 448, 1, 1, 0),
    ("gemm_1024_to_2048_8x8", 3, 1024, 8, 8, 2048, 1, 1, 0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_vs_torch(case, hrp_lib):
    from horopose_b200 import ops
    name, B, Cin, H, W, Cout, k, stride, pad = case
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)
    x = _bf16_round(torch.randn(B, Cin, H, W, generator=g)).cuda()
    w = _bf16_round(torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5)
    scale = (torch.rand(Cout, generator=g) + 0.5)
    bias = torch.randn(Cout, generator=g) * 0.1
    op = ops.ConvOp(_nhwc(x), w, stride=stride, pad=pad, relu=True, scale=scale, bias=bias)
    ref = F.conv2d(x, w.cuda(), stride=stride, padding=pad)
    ref = torch.relu(ref * scale.cuda()[None, :, None, None] + bias.cuda()[None, :, None, None])
    ref = ref.permute(0, 2, 3, 1).contiguous()
    got_simt = op.run(ops.IMPL_SIMT_CHECK).clone()
    torch.cuda.synchronize()
    _report(name + "[simt]", got_simt, ref)
    op.out.zero_()
    got = op.run(ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    _report(name + "[tcgen05]", got, ref)


def test_epilogue_addends(hrp_lib):
    """pre (residual), nearest-upsampled addends (HRNet fuse, HRnet.py:197-208,254-263) and post-ReLU addend
    (cls head, HRnet.py:558-560)."""
    from horopose_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 2, 64, 32, 32
    x = _bf16_round(torch.randn(B, C, H, W, generator=g)).cuda()
    w = _bf16_round(torch.randn(C, C, 3, 3, generator=g) / (C * 9) ** 0.5)
    res = _bf16_round(torch.randn(B, C, H, W, generator=g)).cuda()
    up1 = _bf16_round(torch.randn(B, C, H // 2, W // 2, generator=g)).cuda()
    up2 = _bf16_round(torch.randn(B, C, H // 4, W // 4, generator=g)).cuda()
    post = _bf16_round(torch.randn(B, C, H, W, generator=g)).cuda()
    op = ops.ConvOp(_nhwc(x), w, stride=1, pad=1, relu=True, pre=[_nhwc(res)],
                    up=[(_nhwc(up1), 1), (_nhwc(up2), 2)], post=_nhwc(post))
    ref = F.conv2d(x, w.cuda(), padding=1) + res
    ref = ref + F.interpolate(up1, scale_factor=2, mode="nearest") + F.interpolate(up2, scale_factor=4, mode="nearest")
    ref = (torch.relu(ref) + post).permute(0, 2, 3, 1).contiguous()
    got = op.run()
    torch.cuda.synchronize()
    _report("addends", got, ref)


def test_residual_only_epilogue(hrp_lib):
    """BasicBlock / Bottleneck tail: conv + BN + residual + ReLU (HRnet.py:41-57), ragged batch tile (B=3 @ 8x8)."""
    from horopose_b200 import ops
    g = torch.Generator().manual_seed(11)
    B, C, H = 3, 256, 8
    x = _bf16_round(torch.randn(B, C, H, H, generator=g)).cuda()
    w = _bf16_round(torch.randn(C, C, 3, 3, generator=g) / (C * 9) ** 0.5)
    res = _bf16_round(torch.randn(B, C, H, H, generator=g)).cuda()
    res2 = _bf16_round(torch.randn(B, C, H, H, generator=g)).cuda()
    scale = torch.rand(C, generator=g) + 0.5
    bias = torch.randn(C, generator=g) * 0.1
    op = ops.ConvOp(_nhwc(x), w, stride=1, pad=1, relu=True, scale=scale, bias=bias, pre=[_nhwc(res), _nhwc(res2)])
    ref = F.conv2d(x, w.cuda(), padding=1) * scale.cuda()[None, :, None, None] + bias.cuda()[None, :, None, None]
    ref = torch.relu(ref + res + res2).permute(0, 2, 3, 1).contiguous()
    got = op.run()
    torch.cuda.synchronize()
    _report("residual", got, ref)
    # no ReLU variant (HRNet fuse partial sums / downsample branches)
    op2 = ops.ConvOp(_nhwc(x), w, stride=1, pad=1, relu=False, scale=scale, bias=bias)
    ref2 = (F.conv2d(x, w.cuda(), padding=1) * scale.cuda()[None, :, None, None] + bias.cuda()[None, :, None, None])
    got2 = op2.run()
    torch.cuda.synchronize()
    _report("norelu", got2, ref2.permute(0, 2, 3, 1).contiguous())


def test_pooled_epilogue(hrp_lib):
    """conv + BN + ReLU + global average pool in the epilogue (HRnet.py:562-568)."""
    from horopose_b200 import ops
    g = torch.Generator().manual_seed(6)
    B, Cin, Cout = 3, 128, 256
    x = _bf16_round(torch.randn(B, Cin, 8, 8, generator=g)).cuda()
    w = _bf16_round(torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5)
    op = ops.ConvOp(_nhwc(x), w, relu=True, pool=True, write_out=False)
    ref = torch.relu(F.conv2d(x, w.cuda())).mean(dim=(2, 3))
    got = op.run()
    torch.cuda.synchronize()
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), (got - ref).abs().max()


@pytest.mark.parametrize("cin,hin", [(256, 8), (64, 16)])
def test_deconv_k4s2p1(cin, hin, hrp_lib):
    """ConvTranspose2d(k4,s2,p1)+BN+ReLU as four sub-pixel phases (full_net.py:194-216)."""
    from horopose_b200 import ops
    g = torch.Generator().manual_seed(7)
    B, Cout = 2, 64
    x = _bf16_round(torch.randn(B, cin, hin, hin, generator=g)).cuda()
    w = _bf16_round(torch.randn(cin, Cout, 4, 4, generator=g) / (cin * 4) ** 0.5)
    scale = torch.rand(Cout, generator=g) + 0.5
    bias = torch.randn(Cout, generator=g) * 0.1
    op = ops.ConvOp(_nhwc(x), w, kind=ops.DECONV_K4S2P1, relu=True, scale=scale, bias=bias)
    ref = F.conv_transpose2d(x, w.cuda(), stride=2, padding=1)
    ref = torch.relu(ref * scale.cuda()[None, :, None, None] + bias.cuda()[None, :, None, None])
    ref = ref.permute(0, 2, 3, 1).contiguous()
    got_simt = op.run(ops.IMPL_SIMT_CHECK).clone()
    torch.cuda.synchronize()
    _report("deconv[simt]", got_simt, ref)
    op.out.zero_()
    got = op.run()
    torch.cuda.synchronize()
    _report("deconv[tcgen05]", got, ref)


@pytest.mark.parametrize("k,pad", [(7, 3), (3, 1)])
def test_stem_s2d(k, pad, hrp_lib):
    """Stride-2 stem convs with C_in=3 through the space-to-depth input packing
    (Resnet.py:21 k7 s2 p3; HRnet.py:284 k3 s2 p1)."""
    from horopose_b200 import ops
    g = torch.Generator().manual_seed(8)
    B, H = 2, 64
    x = torch.rand(B, 3, H, H, generator=g).cuda()
    w = _bf16_round(torch.randn(64, 3, k, k, generator=g) / (3 * k * k) ** 0.5)
    xs = ops.pack_input_s2d(x)
    op = ops.ConvOp(xs, w, kind=ops.STEM_S2D, stride=2, pad=pad, relu=True)
    ref = torch.relu(F.conv2d(_bf16_round(x), w.cuda(), stride=2, padding=pad)).permute(0, 2, 3, 1).contiguous()
    got_simt = op.run(ops.IMPL_SIMT_CHECK).clone()
    torch.cuda.synchronize()
    _report("stem[simt]", got_simt, ref)
    op.out.zero_()
    got = op.run()
    torch.cuda.synchronize()
    _report("stem[tcgen05]", got, ref)


def test_maxpool_and_bridges(hrp_lib):
    from horopose_b200 import ops
    g = torch.Generator().manual_seed(9)
    x = _bf16_round(torch.randn(2, 64, 32, 32, generator=g)).cuda()
    xn = ops.nchw_to_nhwc_bf16(x)
    assert torch.equal(xn.float(), x.permute(0, 2, 3, 1))
    back = ops.nhwc_bf16_to_nchw(xn)
    assert torch.equal(back, x)
    mp = ops.maxpool3x3s2(xn)
    ref = F.max_pool2d(x, 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(mp.float(), ref)
    # the network's size and a ragged one
    for shape in ((3, 64, 128, 128), (2, 16, 12, 20)):
        y = _bf16_round(torch.randn(*shape, generator=g)).cuda()
        got = ops.maxpool3x3s2(ops.nchw_to_nhwc_bf16(y))
        assert torch.equal(got.float(), F.max_pool2d(y, 3, 2, 1).permute(0, 2, 3, 1)), shape


# ---- halo-tile kernel (conv_halo.cu): 3x3 stride-1 convs with Cin = Cout in {32, 64} ----
HALO_CASES = [
    # name, B, C, H, W, residual, relu
    ("halo_32_64x64", 3, 32, 64, 64, False, True),
    ("halo_32_64x64_res", 2, 32, 64, 64, True, True),
    ("halo_64_32x32_res", 5, 64, 32, 32, True, True),
    ("halo_64_64x64", 2, 64, 64, 64, False, True),
    ("halo_64_16x16_norelu", 7, 64, 16, 16, True, False),
    ("halo_32_ragged_32x16", 3, 32, 32, 16, True, True),   # non-square; 32*18 = 576 positions = 4.5 tiles
    ("halo_32_8x8", 9, 32, 8, 8, False, True),
    ("halo_64_many_units", 150, 64, 32, 32, True, True),    # more units than SMs: band buffers and accumulators wrap
]


@pytest.mark.parametrize("case", HALO_CASES, ids=[c[0] for c in HALO_CASES])
def test_halo_conv_vs_torch(case, hrp_lib):
    import ctypes as C
    import zlib
    from horopose_b200 import _lib, ops
    name, B, Cc, H, W, residual, relu = case
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)
    x = _bf16_round(torch.randn(B, Cc, H, W, generator=g)).cuda()
    w = _bf16_round(torch.randn(Cc, Cc, 3, 3, generator=g) / (Cc * 9) ** 0.5)
    scale = (torch.rand(Cc, generator=g) + 0.5)
    bias = torch.randn(Cc, generator=g) * 0.1
    res = _bf16_round(torch.randn(B, Cc, H, W, generator=g)).cuda() if residual else None
    pre = (_nhwc(res),) if residual else ()
    op = ops.ConvOp(_nhwc(x), w, stride=1, pad=1, relu=relu, scale=scale, bias=bias, pre=pre)
    L = _lib.lib()
    _lib.check(L.hrp_conv_set_variant(op.handle, C.c_int32(2)))  # raises if the halo kernel is not eligible
    assert L.hrp_conv_variant(op.handle) == 2
    ref = F.conv2d(x, w.cuda(), padding=1) * scale.cuda()[None, :, None, None] + bias.cuda()[None, :, None, None]
    if residual:
        ref = ref + res
    if relu:
        ref = torch.relu(ref)
    ref = ref.permute(0, 2, 3, 1).contiguous()
    op.out.fill_(float("nan"))  # every valid pixel must be written, junk positions must not spill anywhere
    got = op.run(ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all(), f"{name}: unwritten output pixels"
    _report(name + "[halo]", got, ref)
    # bit-exact against the one-tile-per-CTA tcgen05 kernel (same operands, same fp32 accumulation order per tap)
    _lib.check(L.hrp_conv_set_variant(op.handle, C.c_int32(0)))
    got_tile = op.run(ops.IMPL_TCGEN05).clone()
    _lib.check(L.hrp_conv_set_variant(op.handle, C.c_int32(2)))
    got2 = op.run(ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    diff = (got2.float() - got_tile.float()).abs().max().item()
    assert diff <= 2.0 ** -6 * max(1.0, ref.abs().max().item()), f"{name}: halo vs tile kernel differ by {diff}"


@pytest.mark.parametrize("shape", [(3, 64, 64, True), (2, 32, 16, False), (5, 8, 8, True)])
def test_halo_pixel_pair_vs_plain(shape, hrp_lib, monkeypatch):
    """32-channel layers run as 64-channel pixel pairs with zero weight blocks (HaloParams.pair); the plain
    32-channel instantiation of the same kernel (HRP_HALO_PAIR=0) must agree up to fp32 summation order."""
    import ctypes as C
    from horopose_b200 import _lib, ops
    B, H, W, residual = shape
    g = torch.Generator().manual_seed(H * 131 + W)
    x = _bf16_round(torch.randn(B, 32, H, W, generator=g)).cuda()
    w = _bf16_round(torch.randn(32, 32, 3, 3, generator=g) / (32 * 9) ** 0.5)
    scale, bias = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.1
    res = _bf16_round(torch.randn(B, 32, H, W, generator=g)).cuda() if residual else None
    pre = (_nhwc(res),) if residual else ()
    outs = []
    for pair in ("1", "0"):
        monkeypatch.setenv("HRP_HALO_PAIR", pair)
        op = ops.ConvOp(_nhwc(x), w, stride=1, pad=1, relu=True, scale=scale, bias=bias, pre=pre)
        _lib.check(_lib.lib().hrp_conv_set_variant(op.handle, C.c_int32(2)))
        op.out.fill_(float("nan"))
        outs.append(op.run(ops.IMPL_TCGEN05).clone())
    torch.cuda.synchronize()
    assert torch.isfinite(outs[0].float()).all() and torch.isfinite(outs[1].float()).all()
    diff = (outs[0].float() - outs[1].float()).abs().max().item()
    assert diff <= 2.0 ** -6 * max(1.0, outs[1].float().abs().max().item()), f"pair vs plain halo differ by {diff}"


def test_halo_not_eligible_raises(hrp_lib):
    import ctypes as C
    from horopose_b200 import _lib, ops
    x = torch.randn(1, 16, 16, 128, device="cuda").to(torch.bfloat16)
    op = ops.ConvOp(x, torch.randn(128, 128, 3, 3) * 0.03, stride=1, pad=1, relu=True)
    with pytest.raises(_lib.HrpError):
        _lib.check(_lib.lib().hrp_conv_set_variant(op.handle, C.c_int32(2)))


# ---- persistent kernel (conv_gemm_persistent): forced with hrp_conv_set_variant(1) ----
# ~35 % of a 512-image step runs on this kernel, but at the small shapes of CASES the planner may prefer the
# one-tile-per-CTA kernel, so the variant is pinned here and checked against fp32 F.conv2d on every conv class, with the
# batch enlarged where needed so that a persistent CTA walks several tiles (accumulator / staging rings wrap).
@pytest.mark.parametrize("case", CASES, ids=[c[0] + "_persist" for c in CASES])
def test_persistent_conv_vs_torch(case, hrp_lib):
    import ctypes as C
    import zlib
    from horopose_b200 import _lib, ops
    name, B, Cin, H, W, Cout, k, stride, pad = case
    B = max(B, 3) * (8 if H * W <= 1024 else 2)      # >= ~200 M tiles x N tiles: more tiles than SMs
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(zlib.crc32((name + "p").encode()) % 1000)
    x = _bf16_round(torch.randn(B, Cin, H, W, generator=g)).cuda()
    w = _bf16_round(torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5)
    scale = (torch.rand(Cout, generator=g) + 0.5)
    bias = torch.randn(Cout, generator=g) * 0.1
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    conv = F.conv2d(x, w.cuda(), stride=stride, padding=pad) * scale.cuda()[None, :, None, None] + \
        bias.cuda()[None, :, None, None]
    L = _lib.lib()
    # (a) plain epilogue, (b) one residual (EPI_RES: TMA-prefetched into the staging ring), (c) two addends (EPI_PRE)
    res1 = _bf16_round(torch.randn(B, Cout, Ho, Wo, generator=g)).cuda()
    res2 = _bf16_round(torch.randn(B, Cout, Ho, Wo, generator=g)).cuda()
    for tag, pre, ref in (("plain", (), conv), ("res", (res1,), conv + res1), ("pre2", (res1, res2), conv + res1 + res2)):
        op = ops.ConvOp(_nhwc(x), w, stride=stride, pad=pad, relu=True, scale=scale, bias=bias,
                        pre=[_nhwc(t) for t in pre])
        _lib.check(L.hrp_conv_set_variant(op.handle, C.c_int32(1)))
        assert L.hrp_conv_variant(op.handle) == 1
        op.out.fill_(float("nan"))
        got = op.run(ops.IMPL_TCGEN05)
        torch.cuda.synchronize()
        assert torch.isfinite(got.float()).all(), f"{name}[{tag}]: unwritten output"
        _report(f"{name}[persist,{tag}]", got, torch.relu(ref).permute(0, 2, 3, 1).contiguous())
        # same operands as the one-tile-per-CTA kernel (the K order differs with vertical tap sharing): one bf16 ulp
        got = got.clone()
        _lib.check(L.hrp_conv_set_variant(op.handle, C.c_int32(0)))
        got_tile = op.run(ops.IMPL_TCGEN05)
        torch.cuda.synchronize()
        diff = (got_tile.float() - got.float()).abs().max().item()
        assert diff <= 2.0 ** -6 * max(1.0, ref.abs().max().item()), f"{name}[{tag}]: persistent vs tile differ by {diff}"


@pytest.mark.parametrize("cfg", [
    {"HRP_CONV_NPROD": "1", "HRP_CONV_DUAL": "0", "HRP_CONV_RES_STORE": "0"},   # one producer does everything
    {"HRP_CONV_DUAL": "0", "HRP_CONV_RES_STORE": "0"},                          # second producer fetches the addends
    {"HRP_CONV_DUAL": "0"},                                                     # store warp fetches them, one ring
    {"HRP_CONV_KSTAGE": "2"},                                                   # 32 KiB stages
], ids=["one_producer", "addend_producer", "store_warp_addends", "big_stages"])
def test_persistent_pipeline_organisation_does_not_change_a_bit(cfg, hrp_lib, monkeypatch):
    """The persistent kernel's default organisation (two producers feeding two stage rings with two MMA issuers, the
    TMA-store warp fetching residual / addend tiles) only changes WHO moves the data and when: every accumulator still
    receives the same MMAs in the same order, so each simpler organisation must give identical bits.  Covers the plain,
    residual and staged-addend flavours on shapes with more tiles than SMs and >= 4 stages."""
    import ctypes as C
    from horopose_b200 import _lib, ops
    L = _lib.lib()
    g = torch.Generator().manual_seed(23)
    B, Cin, H, Cout = 40, 256, 16, 256
    x = _nhwc(_bf16_round(torch.randn(B, Cin, H, H, generator=g)).cuda())
    w = _bf16_round(torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5)
    res = _nhwc(_bf16_round(torch.randn(B, Cout, H, H, generator=g)).cuda())
    xs = _nhwc(_bf16_round(torch.randn(B, 32, 2 * H, 2 * H, generator=g)).cuda())       # 3x3 stride-2 32 -> 64 fuse conv
    ws = _bf16_round(torch.randn(64, 32, 3, 3, generator=g) / 288 ** 0.5)
    pre = _nhwc(_bf16_round(torch.randn(B, 64, H, H, generator=g)).cuda())
    up = _nhwc(_bf16_round(torch.randn(B, 64, H // 2, H // 2, generator=g)).cuda())

    def run_all():
        outs, descs = [], []
        for kw in (dict(x=x, weight=w, relu=True), dict(x=x, weight=w, relu=True, pre=[res]),
                   dict(x=xs, weight=ws, stride=2, pad=1, relu=True, pre=[pre], up=[(up, 1)])):
            op = ops.ConvOp(kw.pop("x"), kw.pop("weight"), **kw)
            _lib.check(L.hrp_conv_set_variant(op.handle, C.c_int32(1)))
            buf = C.create_string_buffer(256)
            _lib.check(L.hrp_conv_describe(op.handle, buf, 256))
            descs.append(buf.value.decode())
            outs.append(op.run().clone())
        torch.cuda.synchronize()
        return outs, descs

    ref, ref_desc = run_all()
    assert all("persistent" in d for d in ref_desc), ref_desc
    for k, v in cfg.items():
        monkeypatch.setenv(k, v)
    got, got_desc = run_all()
    for a, b, d0, d1 in zip(ref, got, ref_desc, got_desc):
        assert torch.equal(a, b), (d0, d1)


def test_persistent_full_epilogue_deconv_and_pool(hrp_lib):
    """Persistent kernel on the remaining epilogue / geometry classes: nearest-upsampled + post addends (EPI_FULL), the
    4-phase deconv, and the pooled fp32 output."""
    import ctypes as C
    from horopose_b200 import _lib, ops
    L = _lib.lib()
    g = torch.Generator().manual_seed(15)
    B, Cc, H = 6, 64, 32
    x = _bf16_round(torch.randn(B, Cc, H, H, generator=g)).cuda()
    w = _bf16_round(torch.randn(Cc, Cc, 3, 3, generator=g) / (Cc * 9) ** 0.5)
    res = _bf16_round(torch.randn(B, Cc, H, H, generator=g)).cuda()
    up1 = _bf16_round(torch.randn(B, Cc, H // 2, H // 2, generator=g)).cuda()
    post = _bf16_round(torch.randn(B, Cc, H, H, generator=g)).cuda()
    op = ops.ConvOp(_nhwc(x), w, stride=1, pad=1, relu=True, pre=[_nhwc(res)], up=[(_nhwc(up1), 1)], post=_nhwc(post))
    _lib.check(L.hrp_conv_set_variant(op.handle, C.c_int32(1)))
    ref = F.conv2d(x, w.cuda(), padding=1) + res + F.interpolate(up1, scale_factor=2, mode="nearest")
    ref = (torch.relu(ref) + post).permute(0, 2, 3, 1).contiguous()
    _report("persist[full]", op.run(), ref)
    # deconv
    xd = _bf16_round(torch.randn(4, 256, 16, 16, generator=g)).cuda()
    wd = _bf16_round(torch.randn(256, 64, 4, 4, generator=g) / (256 * 4) ** 0.5)
    opd = ops.ConvOp(_nhwc(xd), wd, kind=ops.DECONV_K4S2P1, relu=True)
    _lib.check(L.hrp_conv_set_variant(opd.handle, C.c_int32(1)))
    refd = torch.relu(F.conv_transpose2d(xd, wd.cuda(), stride=2, padding=1)).permute(0, 2, 3, 1).contiguous()
    _report("persist[deconv]", opd.run(), refd)
    # pooled output, 24 images of 8x8: 12 tiles, pool accumulators of different images in one tile
    xp = _bf16_round(torch.randn(24, 128, 8, 8, generator=g)).cuda()
    wp = _bf16_round(torch.randn(256, 128, 1, 1, generator=g) / 128 ** 0.5)
    opp = ops.ConvOp(_nhwc(xp), wp, relu=True, pool=True, write_out=False)
    _lib.check(L.hrp_conv_set_variant(opp.handle, C.c_int32(1)))
    refp = torch.relu(F.conv2d(xp, wp.cuda())).mean(dim=(2, 3))
    gotp = opp.run().clone()
    torch.cuda.synchronize()
    assert torch.allclose(gotp, refp, rtol=1e-4, atol=1e-4), (gotp - refp).abs().max()
    assert torch.equal(opp.run(), gotp)   # two contributions per (image, channel): order-independent, bitwise stable


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("nb", [2, 4])
def test_upsampling_conv_carries_the_branch0_fuse(nb, variant, hrp_lib, monkeypatch):
    """HRNet fuse layer, output branch 0 (HRnet.py:254-263): y0 = relu(x0 + up2(bn(conv1x1(x1))) + up4(t2) + up8(t3)).  The
    1x1 conv of the j = 1 term runs as an upsampling conv (HRP_CONV_UP2: four output phases share one weight matrix) whose
    epilogue adds x0 (output resolution) and the remaining low-resolution terms and applies the ReLU."""
    import ctypes as C
    from horopose_b200 import _lib, ops
    g = torch.Generator().manual_seed(40 + nb)
    B, H = 5, 64
    x0 = _bf16_round(torch.randn(B, 32, H, H, generator=g)).cuda()
    x1 = _bf16_round(torch.randn(B, 64, H // 2, H // 2, generator=g)).cuda()
    w = _bf16_round(torch.randn(32, 64, 1, 1, generator=g) / 8.0)
    scale, bias = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.1
    ups, ref_up = [], 0.0
    for j in range(2, nb):
        t = _bf16_round(torch.randn(B, 32, H >> j, H >> j, generator=g)).cuda()
        ups.append((_nhwc(t), j))
        ref_up = ref_up + F.interpolate(t, scale_factor=2 ** j, mode="nearest")
    monkeypatch.setenv("HRP_CONV_STAGED", "1" if variant == 1 else "0")   # persistent: TMA-staged addends; tile: gathers
    op = ops.ConvOp(_nhwc(x1), w, kind=ops.CONV_UP2, relu=True, scale=scale, bias=bias, pre=[_nhwc(x0)], up=ups)
    _lib.check(_lib.lib().hrp_conv_set_variant(op.handle, C.c_int32(variant)))
    y = F.conv2d(x1, w.cuda()) * scale.cuda()[None, :, None, None] + bias.cuda()[None, :, None, None]
    ref = torch.relu(x0 + F.interpolate(y, scale_factor=2, mode="nearest") + ref_up).permute(0, 2, 3, 1).contiguous()
    got_simt = op.run(ops.IMPL_SIMT_CHECK).clone()
    torch.cuda.synchronize()
    _report("up2[simt]", got_simt, ref)
    op.out.fill_(float("nan"))
    got = op.run(ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all()
    _report(f"up2[tcgen05,{variant}]", got, ref)
