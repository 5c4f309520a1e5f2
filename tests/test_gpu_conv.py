"""(a4)(a5)(a8)(a9)(a11) fused conv -> BN -> LIF layers on the GPU vs the CPU oracle.

Real-valued outputs: tolerance stated per test.  Spikes: north_star's margin rule (bit-exact where the reference's
|h - v_th| > 1e-4, flip rate <= 1e-4).  The tcgen05 kernel is checked twice: through its linear read-out mode (real
outputs, which isolates the GEMM arithmetic) and through its LIF epilogue, for both weight-split settings."""
import pytest
import torch
import torch.nn as nn

from conftest import assert_spikes_match
from oracle import snn_oracle as O
from spiking_diffusion_b200 import _lib, engine
from spiking_diffusion_b200.activation_based import layer, neuron, surrogate

pytestmark = pytest.mark.gpu


def make_block(cin, cout, k=3, stride=1, pad=1, transposed=False, op=0, seed=0, bn=True, lif=True, rate=0.1, beta=0.45):
    g = torch.Generator().manual_seed(seed)
    if transposed:
        conv = layer.ConvTranspose2d(cin, cout, k, stride=stride, padding=pad, output_padding=op)
    else:
        conv = layer.Conv2d(cin, cout, k, stride=stride, padding=pad)
    with torch.no_grad():
        conv.weight.copy_((torch.rand(conv.weight.shape, generator=g) * 2 - 1) / (cin * k * k) ** 0.5)
        conv.bias.copy_((torch.rand(conv.bias.shape, generator=g) * 2 - 1) * 0.1)
    mods = [conv]
    p = {"c.weight": conv.weight.detach().clone(), "c.bias": conv.bias.detach().clone()}
    if bn:
        b = layer.BatchNorm2d(cout)
        with torch.no_grad():
            b.weight.copy_(torch.rand(cout, generator=g) + 0.5)
            b.bias.copy_(torch.rand(cout, generator=g) * 0.2 + beta)
            b.running_mean.copy_(torch.rand(cout, generator=g) * 0.1)
            b.running_var.copy_((torch.rand(cout, generator=g) + 0.5) * rate / 3)
        mods.append(b)
        for k_ in ("weight", "bias", "running_mean", "running_var"):
            p["b." + k_] = getattr(b, k_).detach().clone()
    if lif:
        mods.append(neuron.LIFNode(surrogate_function=surrogate.ATan()))
    seq = layer.SpikingSequential(*mods)
    for m in seq:
        m.step_mode = "m"
    return seq.eval().cuda(), p


def oracle_layer(x_seq, p, bn=True, **kw):
    cur = O.conv_bn(x_seq, p, "c", "b" if bn else None, **kw)
    s, _, h = O.lif_multi_step(cur, return_h=True)
    return cur, s, h


def spikes(shape, rate, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(shape, generator=g) < rate).float()


# ---- CUDA-core path ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [
    dict(cin=1, cout=32, k=3, stride=2, pad=1, H=28, real=True),          # enc.conv1
    dict(cin=32, cout=64, k=3, stride=2, pad=1, H=14),                    # enc.conv2
    dict(cin=64, cout=16, k=1, stride=1, pad=0, H=7),                     # enc.conv3
    dict(cin=16, cout=16, k=1, stride=1, pad=0, H=7, real=True),          # vq.poisson
    dict(cin=16, cout=64, k=3, stride=2, pad=1, H=7, transposed=True, op=1),   # dec.convT1
    dict(cin=64, cout=32, k=3, stride=2, pad=1, H=14, transposed=True, op=1),  # dec.convT2
    dict(cin=2, cout=64, k=3, stride=1, pad=1, H=7, real=True),           # den.conv1
    dict(cin=3, cout=32, k=3, stride=2, pad=1, H=32, real=True),          # CIFAR-shape enc.conv1
])
@pytest.mark.parametrize("T", [4, 16])
def test_simt_layer_via_module_api(cfg, T):
    B, H = 3, cfg["H"]
    seq, p = make_block(cfg["cin"], cfg["cout"], cfg["k"], cfg["stride"], cfg["pad"], cfg.get("transposed", False),
                        cfg.get("op", 0), seed=T)
    if cfg.get("real"):
        x = (torch.rand((1, B, cfg["cin"], H, H), generator=torch.Generator().manual_seed(1)) - 0.5).repeat(T, 1, 1, 1, 1)
        x = x * 3
    else:
        x = spikes((T, B, cfg["cin"], H, H), 0.12, 2)
    out = seq(x.cuda())
    cur, s_ref, h_ref = oracle_layer(x, p, stride=cfg["stride"], padding=cfg["pad"],
                                     transposed=cfg.get("transposed", False), output_padding=cfg.get("op", 0))
    assert 0.01 < float(s_ref.mean()) < 0.6
    assert_spikes_match(out, s_ref, h_ref, str(cfg))
    lif = seq[2]
    assert isinstance(lif.v, torch.Tensor) and lif.v.shape == s_ref.shape[1:]   # state protocol (neuron.py:260-263)


def test_simt_state_persists_across_calls_until_reset():
    from spiking_diffusion_b200.activation_based import functional
    seq, p = make_block(32, 64, 3, 2, 1, seed=9)
    x = spikes((4, 2, 32, 14, 14), 0.12, 3)
    s1 = seq(x.cuda()).cpu()
    s2 = seq(x.cuda()).cpu()          # second call starts from the first call's membrane potential
    cur = O.conv_bn(x, p, "c", "b", stride=2, padding=1)
    r1, v1, h1 = O.lif_multi_step(cur, return_h=True)
    r2, v2, h2 = O.lif_multi_step(cur, v1, return_h=True)
    assert_spikes_match(s1, r1, h1, "call1")
    assert_spikes_match(s2, r2, h2, "call2")
    assert not torch.equal(s1, s2)
    functional.reset_net(seq)
    assert_spikes_match(seq(x.cuda()).cpu(), r1, h1, "after reset")


def test_unfused_layers_match_torch_semantics():
    conv = layer.Conv2d(8, 12, 3, stride=1, padding=1, step_mode="m").cuda().eval()
    bn = layer.BatchNorm2d(12, step_mode="m").cuda().eval()
    with torch.no_grad():
        bn.running_mean.uniform_(-0.5, 0.5); bn.running_var.uniform_(0.5, 2.0); bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-1, 1)
    x = torch.randn(3, 2, 8, 9, 9)
    y = conv(x.cuda())
    p = {"c.weight": conv.weight.detach().cpu(), "c.bias": conv.bias.detach().cpu()}
    ref = O.conv_bn(x, p, "c", None, stride=1, padding=1)
    assert float((y.cpu() - ref).abs().max()) <= 2e-5
    z = bn(y)
    p.update({"b." + k: getattr(bn, k).detach().cpu() for k in ("weight", "bias", "running_mean", "running_var")})
    assert float((z.cpu() - O.conv_bn(x, p, "c", "b", stride=1, padding=1)).abs().max()) <= 5e-5
    with pytest.raises(ValueError):
        conv(torch.zeros(2, 8, 9, 9).cuda())   # 'm' mode wants [T, N, C, H, W]  (layer.py:169-170)
    convT = layer.ConvTranspose2d(8, 4, 3, stride=2, padding=1, output_padding=1, step_mode="s").cuda().eval()
    xs = torch.randn(2, 8, 7, 7)
    ref = torch.nn.functional.conv_transpose2d(xs, convT.weight.detach().cpu(), convT.bias.detach().cpu(), stride=2,
                                               padding=1, output_padding=1)
    assert float((convT(xs.cuda()).cpu() - ref).abs().max()) <= 2e-5


# ---- tcgen05 path -----------------------------------------------------------------------------------------
def tc_layer(seq, T, B, H, out_kind, nsplit, impl="tc", **kw):
    conv = seq[0]
    bn = seq[1] if len(seq) > 1 and isinstance(seq[1], nn.BatchNorm2d) else None
    lif = seq[-1] if isinstance(seq[-1], neuron.LIFNode) else None
    return engine.FusedLayer(conv, bn, lif, T=T, B=B, H_in=H, W_in=H, in_kind=_lib.IN_STF, out_kind=out_kind,
                             impl=impl, nsplit=nsplit, **kw)


@pytest.mark.parametrize("nsplit,tol", [(2, 2e-5), (1, 2e-3)])
@pytest.mark.parametrize("cin,cout,B,H", [(64, 128, 2, 7), (64, 128, 5, 7), (320, 128, 4, 7), (128, 512, 3, 8)])
def test_tc_linear_readout_matches_oracle(nsplit, tol, cin, cout, B, H):
    """GEMM arithmetic in isolation: T-summed spike counts in, mean-over-T real values out (conv6's mode)."""
    T = 4
    seq, p = make_block(cin, cout, seed=cin + cout, bn=False, lif=False)
    fl = tc_layer(seq, T, B, H, _lib.OUT_MEAN_T, nsplit, in_T=1)
    s = spikes((T, B, cin, H, H), 0.15, 7)
    counts = s.sum(0, keepdim=True)
    x_stf = engine.stf_from_nchw(counts.cuda())
    out = fl.run(x_stf, fl.alloc_out()).cpu()                       # [B, H, W, C_out]
    ref = (O.conv_bn(s, p, "c", None, stride=1, padding=1).sum(0) / T).permute(0, 2, 3, 1)
    err = float((out - ref).abs().max())
    assert err <= tol, f"max abs err {err:.3e}"


def test_tc_linear_readout_concat_two_sources():
    T, B, H = 4, 3, 7
    seq, p = make_block(320, 128, seed=5, bn=False, lif=False)
    fl = tc_layer(seq, T, B, H, _lib.OUT_MEAN_T, 2, in_T=1, C_in0=256)
    s5, s1 = spikes((T, B, 256, H, H), 0.1, 1), spikes((T, B, 64, H, H), 0.2, 2)
    a = engine.stf_from_nchw(s5.sum(0, keepdim=True).cuda())
    b = engine.stf_from_nchw(s1.sum(0, keepdim=True).cuda())
    out = fl.run(a, fl.alloc_out(), x2=b).cpu()
    ref = (O.conv_bn(torch.cat((s5, s1), dim=2), p, "c", None, stride=1, padding=1).sum(0) / T).permute(0, 2, 3, 1)
    assert float((out - ref).abs().max()) <= 2e-5


@pytest.mark.parametrize("nsplit", [2, 1])
@pytest.mark.parametrize("cin,cout,B,H,T", [
    (64, 128, 2, 7, 4), (64, 128, 7, 7, 4), (128, 256, 4, 7, 4), (256, 512, 4, 7, 4), (512, 256, 3, 7, 4),
    (64, 128, 3, 8, 4), (64, 128, 3, 7, 8), (64, 64, 2, 7, 16), (128, 128, 2, 7, 2), (64, 128, 300, 7, 4),
])
def test_tc_conv_bn_lif_matches_oracle(nsplit, cin, cout, B, H, T):
    seq, p = make_block(cin, cout, seed=cin * 3 + cout + T)
    fl = tc_layer(seq, T, B, H, _lib.OUT_LIF, nsplit)
    s_in = spikes((T, B, cin, H, H), 0.1, B)
    x_stf = engine.stf_from_nchw(s_in.cuda())
    out, osum = fl.alloc_out(), fl.alloc_sum()
    fl.run(x_stf, out, out_sum=osum)
    got = engine.stf_to_nchw(out, T, B, cout, H, H).cpu()
    cur, s_ref, h_ref = oracle_layer(s_in, p, stride=1, padding=1)
    assert 0.02 < float(s_ref.mean()) < 0.5
    if nsplit == 2:
        assert_spikes_match(got, s_ref, h_ref, f"tc {cin}->{cout} B{B} H{H} T{T}")
    else:   # single fp16 weight term: 11-bit weights, reported, looser bar
        rate = float((got != s_ref).float().mean())
        assert rate <= 2e-3, rate
    cnt = engine.stf_to_nchw(osum, 1, B, cout, H, H).cpu()[0]
    assert torch.equal(cnt, got.sum(0))
    # pad rows of the STF output stay zero (the next layer's halo reads them)
    assert float(out.float().sum()) == float(got.sum())


def test_tc_matches_simt_implementation_and_state():
    """Same layer through both implementations, with a carried membrane state."""
    T, B, H, cin, cout = 4, 4, 7, 64, 128
    seq, p = make_block(cin, cout, seed=77)
    a, b = tc_layer(seq, T, B, H, _lib.OUT_LIF, 2), tc_layer(seq, T, B, H, _lib.OUT_LIF, 2, impl="simt")
    s_in = spikes((T, B, cin, H, H), 0.1, 4)
    x_stf = engine.stf_from_nchw(s_in.cuda())
    va, vb = a.alloc_state(), b.alloc_state()
    oa, ob = a.alloc_out(), b.alloc_out()
    for _ in range(2):
        a.run(x_stf, oa, v=va); b.run(x_stf, ob, v=vb)
    ga, gb = engine.stf_to_nchw(oa, T, B, cout, H, H), engine.stf_to_nchw(ob, T, B, cout, H, H)
    assert float((ga != gb).float().mean()) <= 1e-4
    assert float((va - vb).abs().max()) <= 1e-4 or float(((va - vb).abs() > 1e-4).float().mean()) <= 1e-4


def test_tc_rejects_unsupported_shapes():
    seq, _ = make_block(30, 64, seed=1)
    with pytest.raises(ValueError):
        tc_layer(seq, 4, 2, 7, _lib.OUT_LIF, 2)       # C_in not a multiple of 16
    seq, _ = make_block(32, 64, 3, 2, 1, seed=1)
    with pytest.raises(ValueError):
        tc_layer(seq, 4, 2, 14, _lib.OUT_LIF, 2)      # stride 2


# ---- stride-2 transposed convolution on the tcgen05 kernel (decoder) -----------------------------------------
@pytest.mark.parametrize("T,B,C,H,W", [(4, 3, 16, 7, 7), (1, 2, 8, 5, 3), (16, 1, 24, 14, 14)])
def test_stf_upsample2x_is_zero_insertion(T, B, C, H, W):
    s = spikes((T, B, C, H, W), 0.3, 11).cuda()
    out = engine.stf_empty(T, B, C, 2 * H, 2 * W, s.device)
    out.fill_(7.0)   # every valid row must be overwritten
    _lib.check(_lib.lib().sd_stf_upsample2x(_lib.ptr(engine.stf_from_nchw(s)), _lib.ptr(out), T, B, C, H, W,
                                            _lib.stream_ptr()))
    got = engine.stf_to_nchw(out, T, B, C, 2 * H, 2 * W)
    ref = torch.zeros_like(got)
    ref[..., ::2, ::2] = s
    assert torch.equal(got, ref)


@pytest.mark.parametrize("cin,cout,B,H,T", [(16, 64, 5, 7, 4), (64, 32, 3, 14, 4), (64, 32, 2, 14, 16), (16, 64, 3, 8, 8)])
def test_transposed_conv_as_upsampled_tc_conv_matches_oracle_and_simt(cin, cout, B, H, T):
    """ConvTranspose2d(k3, s2, p1, op1) + BN + LIF: the tcgen05 route (zero-insertion + flipped 3x3 conv) against the
    oracle's F.conv_transpose2d and against the CUDA-core transposed kernel."""
    seq, p = make_block(cin, cout, 3, stride=2, pad=1, transposed=True, op=1, seed=5)
    conv, bn, lif = seq[0], seq[1], seq[2]
    assert engine._UpsampledConvT.eligible(conv)
    s_in = spikes((T, B, cin, H, H), 0.15, 9)
    cur, s_ref, h_ref = oracle_layer(s_in, p, stride=2, padding=1, transposed=True, output_padding=1)
    x_stf = engine.stf_from_nchw(s_in.cuda())
    up = engine.stf_empty(T, B, cin, 2 * H, 2 * H, x_stf.device)
    _lib.check(_lib.lib().sd_stf_upsample2x(_lib.ptr(x_stf), _lib.ptr(up), T, B, cin, H, H, _lib.stream_ptr()))
    tc = engine.FusedLayer(engine._UpsampledConvT(conv), bn, lif, T=T, B=B, H_in=2 * H, W_in=2 * H,
                           in_kind=_lib.IN_STF, out_kind=_lib.OUT_LIF, impl="tc")
    simt = engine.FusedLayer(conv, bn, lif, T=T, B=B, H_in=H, W_in=H, in_kind=_lib.IN_STF, out_kind=_lib.OUT_LIF,
                             impl="simt")
    o_tc, o_simt = tc.alloc_out(), simt.alloc_out()
    tc.run(up, o_tc)
    simt.run(x_stf, o_simt)
    g_tc = engine.stf_to_nchw(o_tc, T, B, cout, 2 * H, 2 * H).cpu()
    g_simt = engine.stf_to_nchw(o_simt, T, B, cout, 2 * H, 2 * H).cpu()
    assert_spikes_match(g_tc, s_ref, h_ref, "convT on tcgen05")
    assert_spikes_match(g_simt, s_ref, h_ref, "convT on CUDA cores")
    assert float((g_tc != g_simt).float().mean()) <= 1e-4


@pytest.mark.parametrize("T,B,C,H,W", [(4, 3, 16, 14, 14), (1, 2, 8, 7, 5), (16, 1, 24, 8, 8)])
def test_stf_subsample2x_keeps_even_positions(T, B, C, H, W):
    s = spikes((T, B, C, H, W), 0.3, 12).cuda()
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    out = engine.stf_empty(T, B, C, Ho, Wo, s.device)
    _lib.check(_lib.lib().sd_stf_subsample2x(_lib.ptr(engine.stf_from_nchw(s)), _lib.ptr(out), T, B, C, H, W,
                                             _lib.stream_ptr()))
    assert torch.equal(engine.stf_to_nchw(out, T, B, C, Ho, Wo), s[..., ::2, ::2])


@pytest.mark.parametrize("cin,cout,B,H,T", [(32, 64, 5, 14, 4), (32, 64, 2, 14, 16), (16, 32, 3, 7, 8)])
def test_stride2_conv_as_subsampled_tc_conv_matches_oracle_and_simt(cin, cout, B, H, T):
    """Conv2d(k3, s2, p1) + BN + LIF: stride-1 tcgen05 conv followed by the even-position pick, against the oracle's
    strided F.conv2d and against the CUDA-core strided kernel."""
    seq, p = make_block(cin, cout, 3, stride=2, pad=1, seed=6)
    conv, bn, lif = seq[0], seq[1], seq[2]
    assert engine._Stride1Conv.eligible(conv)
    s_in = spikes((T, B, cin, H, H), 0.15, 10)
    cur, s_ref, h_ref = oracle_layer(s_in, p, stride=2, padding=1)
    Ho = (H + 1) // 2
    x_stf = engine.stf_from_nchw(s_in.cuda())
    tc = engine.FusedLayer(engine._Stride1Conv(conv), bn, lif, T=T, B=B, H_in=H, W_in=H, in_kind=_lib.IN_STF,
                           out_kind=_lib.OUT_LIF, impl="tc")
    simt = engine.FusedLayer(conv, bn, lif, T=T, B=B, H_in=H, W_in=H, in_kind=_lib.IN_STF, out_kind=_lib.OUT_LIF,
                             impl="simt")
    full, o_simt = tc.alloc_out(), simt.alloc_out()
    tc.run(x_stf, full)
    o_tc = engine.stf_empty(T, B, cout, Ho, Ho, x_stf.device)
    _lib.check(_lib.lib().sd_stf_subsample2x(_lib.ptr(full), _lib.ptr(o_tc), T, B, cout, H, H, _lib.stream_ptr()))
    simt.run(x_stf, o_simt)
    g_tc = engine.stf_to_nchw(o_tc, T, B, cout, Ho, Ho).cpu()
    g_simt = engine.stf_to_nchw(o_simt, T, B, cout, Ho, Ho).cpu()
    assert_spikes_match(g_tc, s_ref, h_ref, "strided conv on tcgen05")
    assert_spikes_match(g_simt, s_ref, h_ref, "strided conv on CUDA cores")
    assert float((g_tc != g_simt).float().mean()) <= 1e-4


@pytest.mark.parametrize("impl", ["tc", "tc-i8", "simt"])
@pytest.mark.parametrize("tau,v_th,v_reset", [(2.0, 1.0, None), (3.0, 0.7, 0.0), (2.0, 1.0, -0.25), (4.0, 0.5, None)])
def test_fused_layer_with_general_lif_parameters(impl, tau, v_th, v_reset):
    """Soft reset, tau that is not a power of two, non-zero reset value and threshold: the general LIF branch of the
    fused epilogues (the fast branch covers only the reference's own tau=2, hard reset to 0)."""
    T, B, H, cin, cout = 4, 6, 7, 64, 128
    seq, p = make_block(cin, cout, seed=21)
    lif = neuron.LIFNode(tau=tau, v_threshold=v_th, v_reset=v_reset, surrogate_function=surrogate.ATan(), step_mode="m")
    i8 = impl == "tc-i8"     # the int8-digit kernel (two passes at T = 4: the potential crosses a pass boundary in shared memory)
    layer_ = engine.FusedLayer(seq[0], seq[1], lif, T=T, B=B, H_in=H, W_in=H, in_kind=_lib.IN_STF, out_kind=_lib.OUT_LIF,
                               impl="tc" if i8 else impl, nsplit=3 if i8 else 2)
    s_in = spikes((T, B, cin, H, H), 0.12, 3)
    cur = O.conv_bn(s_in, p, "c", "b", padding=1)
    s_ref, v_ref, h_ref = O.lif_multi_step(cur, tau=tau, v_threshold=v_th, v_reset=v_reset, return_h=True)
    v = layer_.alloc_state()
    x_in = engine.stf8_from_nchw(s_in.cuda()) if i8 else engine.stf_from_nchw(s_in.cuda())
    out = layer_.run(x_in, layer_.alloc_out(), v=v)
    got = engine.stf8_to_nchw(out, T, B, cout, H, H) if i8 else engine.stf_to_nchw(out, T, B, cout, H, H)
    assert_spikes_match(got, s_ref, h_ref, f"{impl} tau={tau} v_th={v_th} v_reset={v_reset}", v_th=v_th)
    v_got = torch.empty((B, cout, H, H), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().sd_state_convert(_lib.ptr(v), _lib.ptr(v_got), B, cout, H, H, 0, _lib.stream_ptr()))
    same = (got.cpu() == s_ref).all(dim=0)
    assert float((v_got.cpu() - v_ref)[same].abs().max()) <= 2e-5


@pytest.mark.parametrize("T,B", [(8, 4), (16, 3), (8, 40)])
def test_tc_multi_pass_layers_carry_the_membrane_state_across_calls(T, B):
    """T = 8 / 16 run as passes of 4 timesteps (small batches: T-parallel passes + a separate LIF kernel).  Two
    consecutive calls with a caller-owned state must continue the recurrence exactly like the oracle does."""
    H, cin, cout = 7, 64, 128
    seq, p = make_block(cin, cout, seed=31)
    lyr = tc_layer(seq, T, B, H, _lib.OUT_LIF, 2)
    v = lyr.alloc_state()
    out, osum = lyr.alloc_out(), lyr.alloc_sum()
    v_ref = None
    for call in range(2):
        s_in = spikes((T, B, cin, H, H), 0.12, 40 + call)
        cur = O.conv_bn(s_in, p, "c", "b", padding=1)
        s_ref, v_ref, h_ref = O.lif_multi_step(cur, v=v_ref, return_h=True)
        lyr.run(engine.stf_from_nchw(s_in.cuda()), out, out_sum=osum, v=v)
        got = engine.stf_to_nchw(out, T, B, cout, H, H)
        assert_spikes_match(got, s_ref, h_ref, f"call {call}")
        if torch.equal(got.cpu(), s_ref):
            cnt = engine.stf_to_nchw(osum, 1, B, cout, H, H)[0]
            assert torch.equal(cnt.cpu(), s_ref.sum(0))
            v_got = torch.empty((B, cout, H, H), dtype=torch.float32, device="cuda")
            _lib.check(_lib.lib().sd_state_convert(_lib.ptr(v), _lib.ptr(v_got), B, cout, H, H, 0, _lib.stream_ptr()))
            assert float((v_got.cpu() - v_ref).abs().max()) <= 2e-5


# ---- kind::i8 path (nsplit = 3): u8 spikes, three int8 weight digits, exact int32 accumulation ------------------------
@pytest.mark.parametrize("T,B,C,H,W", [(4, 3, 16, 7, 7), (2, 2, 64, 5, 3), (8, 1, 32, 8, 8)])
def test_stf8_round_trip(T, B, C, H, W):
    s = spikes((T, B, C, H, W), 0.3, T + C)
    buf = engine.stf8_from_nchw(s.cuda())
    assert buf.dtype == torch.uint8 and buf.numel() == _lib.lib().sd_stf_bytes(T, B, C, H, W)
    back = engine.stf8_to_nchw(buf, T, B, C, H, W).cpu()
    assert torch.equal(back, s)
    vals = set(torch.unique(buf).tolist())
    assert vals <= {0, 1, 128}
    assert int((buf == 1).sum()) == int((buf == 128).sum()) == int(s.sum())     # one byte per spike in each plane


@pytest.mark.parametrize("cin,cout,B,H,T", [
    (64, 128, 2, 7, 4), (64, 128, 7, 7, 4), (128, 256, 4, 7, 4), (256, 512, 4, 7, 4), (512, 256, 3, 7, 4),
    (64, 128, 3, 8, 4), (64, 128, 3, 7, 8), (64, 64, 2, 7, 16), (128, 128, 2, 7, 2), (64, 128, 300, 7, 4),
    (32, 48, 5, 7, 4), (96, 16, 40, 8, 2),
])
def test_tc_int8_conv_bn_lif_matches_oracle(cin, cout, B, H, T):
    """The same layers as the fp16 test through the int8-digit kernel, to the same bar: spikes bit-exact wherever the
    oracle's |h - v_th| > 1e-4, flip rate <= 1e-4; the T-sum equals the sum of the written spikes; pad rows stay 0."""
    seq, p = make_block(cin, cout, seed=cin * 3 + cout + T)
    fl = tc_layer(seq, T, B, H, _lib.OUT_LIF, 3)
    assert fl.desc.in_kind == _lib.IN_STF8 and fl.desc.out_kind == _lib.OUT_LIF8 and fl.impl == "tc"
    s_in = spikes((T, B, cin, H, H), 0.1, B)
    x8 = engine.stf8_from_nchw(s_in.cuda())
    out, osum = fl.alloc_out(), fl.alloc_sum()
    fl.run(x8, out, out_sum=osum)
    got = engine.stf8_to_nchw(out, T, B, cout, H, H).cpu()
    assert not bool(torch.isnan(got).any()), "the s and 128*s planes disagree"
    cur, s_ref, h_ref = oracle_layer(s_in, p, stride=1, padding=1)
    assert 0.02 < float(s_ref.mean()) < 0.5
    assert_spikes_match(got, s_ref, h_ref, f"tc-i8 {cin}->{cout} B{B} H{H} T{T}")
    cnt = engine.stf_to_nchw(osum, 1, B, cout, H, H).cpu()[0]
    assert torch.equal(cnt, got.sum(0))
    assert int((out == 1).sum()) == int(got.sum()) and int((out == 128).sum()) == int(got.sum())
    assert set(torch.unique(out).tolist()) <= {0, 1, 128}


def test_tc_int8_matches_fp16_path_and_carries_state():
    """Same layer through the int8-digit and the two-fp16-term kernels with a caller-owned membrane state over two
    calls: identical spikes (no near-threshold neuron in this fixture), potentials equal to fp32 rounding."""
    T, B, H, cin, cout = 4, 6, 7, 128, 256
    seq, p = make_block(cin, cout, seed=5)
    a, b = tc_layer(seq, T, B, H, _lib.OUT_LIF, 3), tc_layer(seq, T, B, H, _lib.OUT_LIF, 2)
    va, vb = a.alloc_state(), b.alloc_state()
    oa, ob = a.alloc_out(), b.alloc_out()
    for call in range(2):
        s_in = spikes((T, B, cin, H, H), 0.1, 9 + call).cuda()
        a.run(engine.stf8_from_nchw(s_in), oa, v=va)
        b.run(engine.stf_from_nchw(s_in), ob, v=vb)
        ga, gb = engine.stf8_to_nchw(oa, T, B, cout, H, H), engine.stf_to_nchw(ob, T, B, cout, H, H)
        assert float((ga != gb).float().mean()) <= 1e-4
        assert float((va - vb).abs().max()) <= 1e-4 or float(((va - vb).abs() > 1e-4).float().mean()) <= 1e-4


def test_int8_weight_digits_reconstruct_the_weights():
    """sd_conv_pack_weights_tc (nsplit = 3): 256 * (128 * d0 + d1) + d2, times the returned channel scale, equals the fp32
    weight to 2^-21 of the channel's largest weight, and the L1 bound that rules out int32 overflow holds."""
    import ctypes
    cin, cout, T, B, H = 64, 128, 4, 4, 7
    seq, p = make_block(cin, cout, seed=11)
    fl = tc_layer(seq, T, B, H, _lib.OUT_LIF, 3)
    L = _lib.lib()
    d = fl.desc
    w = seq[0].weight.detach().float().contiguous()
    chan = torch.empty(cout, device="cuda")
    packed = torch.empty(L.sd_conv_weight_bytes_tc(ctypes.byref(d)), dtype=torch.uint8, device="cuda")
    _lib.check(L.sd_conv_pack_weights_tc(ctypes.byref(d), w.data_ptr(), packed.data_ptr(), chan.data_ptr(), _lib.stream_ptr()))
    layout = L.sd_conv_weight_layout_tc(ctypes.byref(d))
    n_tile, kblk, pair = layout >> 16, (layout >> 4) & 0xFFF, layout & 1
    halves, chunks = (2 if pair else 1), kblk // 16
    nh = n_tile // halves
    main = (cout // n_tile if cout >= n_tile else 1) * (cin // kblk) * 9 * 3 * chunks * n_tile * 16
    dg = packed[:main].view(torch.int8).reshape(-1, cin // kblk, 9, halves, 3, chunks, nh, 16).cpu().int()
    wfix = 256 * (128 * dg[:, :, :, :, 0] + dg[:, :, :, :, 1]) + dg[:, :, :, :, 2]       # [nt, kb, tap, hf, chunk, n, 16]
    # -> [co, ci, tap]
    wfix = wfix.permute(0, 3, 5, 1, 4, 6, 2).reshape(-1, cin, 9)[:cout]
    rec = wfix.double() * chan.cpu().double()[:, None, None]
    wref = w.cpu().reshape(cout, cin, 9).double()
    err = (rec - wref).abs().amax(dim=(1, 2))
    assert bool((err <= wref.abs().amax(dim=(1, 2)) * 2.0 ** -21).all())
    assert int(dg[:, :, :, :, 1].abs().max()) <= 64 and int(wfix.abs().sum(dim=(1, 2)).max()) < 2 ** 31
