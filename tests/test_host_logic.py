"""Host-side logic that needs no GPU: module structure / state_dict compatibility with the reference, the
step-mode and memory protocol, BN folding arithmetic, shard bookkeeping."""
import os

import pytest
import torch

from oracle import snn_oracle as O
from spiking_diffusion_b200 import engine, synth
from spiking_diffusion_b200.activation_based import base, functional, layer, neuron, surrogate
from spiking_diffusion_b200.snn_model.vae_model import SNN_VQVAE, VectorQuantizer
from spiking_diffusion_b200.snn_model.vq_diffusion import AbsorbingDiffusion, DummyModel

REF_VQVAE_KEYS = (
    [f"encoder.snn_convs.{i}.{k}" for i in (0, 3, 6) for k in ("weight", "bias")]
    + [f"encoder.snn_convs.{i}.{k}" for i in (1, 4, 7)
       for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")]
    + ["vq_layer.alpha", "vq_layer.memout.coef", "vq_layer.embeddings.weight"]
    + [f"vq_layer.poisson.0.{k}" for k in ("weight", "bias")]
    + [f"vq_layer.poisson.1.{k}" for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")]
    + [f"decoder.snn_convs.{i}.{k}" for i in (0, 3, 6) for k in ("weight", "bias")]
    + [f"decoder.snn_convs.{i}.{k}" for i in (1, 4)
       for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")]
    + ["memout.coef"])


def test_state_dict_keys_match_reference_checkpoint_format():
    """Key list of SURVEY.md section 8(f) rank 3 (R/main.py:199,286 checkpoints)."""
    m = SNN_VQVAE(1, 16, 128, torch.tensor(1.0))
    assert sorted(m.state_dict().keys()) == sorted(REF_VQVAE_KEYS)
    assert m.state_dict()["memout.coef"].shape == (16, 1, 1, 1, 1)          # default T = 16, as the reference
    d = DummyModel(1, 128)
    keys = sorted(d.state_dict().keys())
    want = sorted([f"conv{i}.0.{k}" for i in range(1, 7) for k in ("weight", "bias")]
                  + [f"conv{i}.1.{k}" for i in range(1, 6)
                     for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")])
    assert keys == want
    assert sum(p.numel() for p in d.parameters()) == 3101504
    assert sum(p.numel() for p in m.parameters()) == 50658
    # LIF state is not part of the state dict (SJ/activation_based/base.py:170-171)
    assert not any(k.endswith(".v") for k in keys)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference not mounted")
def test_reference_state_dicts_load_into_our_modules_and_back():
    from oracle import ref_loader
    R = ref_loader.load()
    ours, theirs = SNN_VQVAE(1, 16, 128, torch.tensor(1.0)), R.SNN_VQVAE(1, 16, 128, torch.tensor(1.0))
    assert ours.load_state_dict(theirs.state_dict()).missing_keys == []
    assert theirs.load_state_dict(ours.state_dict()).missing_keys == []
    od, td = DummyModel(1, 128), R.DummyModel(1, 128)
    od.load_state_dict(td.state_dict()); td.load_state_dict(od.state_dict())


def test_step_mode_and_memory_protocol():
    n = neuron.LIFNode(surrogate_function=surrogate.ATan())
    assert n.step_mode == "s" and n.backend == "torch" and n.v == 0.0 and n.tau == 2.0 and n.v_threshold == 1.0
    m = SNN_VQVAE(1, 16, 128, torch.tensor(1.0))
    functional.set_step_mode(m, "m")
    assert all(x.step_mode == "m" for x in m.modules() if hasattr(x, "step_mode"))
    functional.set_backend(m, "cupy")      # accepted name; same kernel, no dispatch
    n.v = torch.ones(3)
    assert "v" not in n.state_dict() and list(dict(n.named_memories())) == ["v"]
    n.reset()
    assert n.v == 0.0
    soft = neuron.LIFNode(v_reset=None)
    assert soft.v == 0.0 and soft.v_reset is None
    with pytest.raises(ValueError):
        n.step_mode = "q"
    with pytest.raises(NotImplementedError):
        n.backend = "lava"
    with pytest.raises(AssertionError):
        neuron.LIFNode(tau=2)               # isinstance(tau, float), neuron.py:707
    rep = n._replicate_for_data_parallel()
    assert rep._memories is not n._memories and rep.v == n.v


def test_sequential_grouping():
    d = DummyModel(1, 128)
    st = d.conv1._stages()
    assert len(st) == 1 and isinstance(st[0][2], neuron.LIFNode)
    assert d.conv6._stages()[0][1] is None and d.conv6._stages()[0][2] is None
    enc = SNN_VQVAE(1, 16, 128, 1.0).encoder.snn_convs
    assert [type(s[0]).__name__ for s in enc._stages()] == ["Conv2d"] * 3


def test_bn_fold_matches_conv_then_bn():
    g = torch.Generator().manual_seed(0)
    conv = torch.nn.Conv2d(5, 7, 3, padding=1)
    bn = torch.nn.BatchNorm2d(7).eval()
    with torch.no_grad():
        bn.running_mean.copy_(torch.rand(7, generator=g)); bn.running_var.copy_(torch.rand(7, generator=g) + 0.2)
        bn.weight.copy_(torch.rand(7, generator=g) + 0.5); bn.bias.copy_(torch.rand(7, generator=g))
    x = torch.randn(2, 5, 6, 6, generator=g)
    scale, shift = engine.fold_bn(conv.bias, 7, bn, "cpu")
    y = torch.nn.functional.conv2d(x, conv.weight, None, padding=1) * scale[None, :, None, None] + shift[None, :, None, None]
    assert float((y - bn(conv(x))).abs().max()) <= 1e-5
    p = {"c.bias": conv.bias.detach(), "b.weight": bn.weight.detach(), "b.bias": bn.bias.detach(),
         "b.running_mean": bn.running_mean, "b.running_var": bn.running_var}
    s2, h2 = O.bn_affine(p, "c", "b")
    assert torch.allclose(s2, scale) and torch.allclose(h2, shift)


def test_cpu_tensors_are_rejected_not_silently_computed():
    m = SNN_VQVAE(1, 16, 128, torch.tensor(1.0), T=4).eval()
    functional.set_step_mode(m, "m")
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(4, 1, 1, 28, 28), torch.zeros(1, 1, 28, 28))
    with pytest.raises(RuntimeError, match="no CPU path"):
        DummyModel(1, 128, T=4).eval()(torch.zeros(1, 1, 7, 7), torch.ones(1).long())
    with pytest.raises(RuntimeError, match="no CPU path"):
        SNN_VQVAE(1, 16, 128, 1.0)(torch.zeros(16, 1, 1, 28, 28), torch.zeros(1, 1, 28, 28))   # train mode, CPU tensors


def test_absorbing_diffusion_attributes():
    ab = AbsorbingDiffusion(DummyModel(1, 128, T=4), mask_id=128)
    assert (ab.n_samples, ab.num_timesteps, ab.shape, ab.mask_id, ab.num_classes) == (16, 49, [7, 7], 128, 128)
    ab8 = AbsorbingDiffusion(DummyModel(3, 128, T=4), mask_id=128, shape=(8, 8), n_samples=1024)
    assert ab8.num_timesteps == 64
    x0 = torch.randint(0, 128, (4, 1, 7, 7))
    x_t, ign, mask = ab.q_sample(x0, torch.tensor([49, 1, 25, 49]))
    assert bool((x_t[mask] == 128).all()) and bool((ign[~mask] == -1).all())


def test_sub_batch_plan_is_pair_aligned_and_covers_the_shard(monkeypatch):
    monkeypatch.delenv("SD_SAMPLER_STREAMS", raising=False)
    # cfg2: 256 images x 49 rows = 49 tile pairs -> 5 sub-batches ending just below 10 pairs (2560 rows)
    assert engine.plan_sub_batches(256, 49) == [52, 52, 52, 52, 48]
    assert engine.plan_sub_batches(512, 49) == [256, 256]          # more than one wave of pairs: two halves
    assert engine.plan_sub_batches(1024, 64) == [512, 512]
    assert engine.plan_sub_batches(32, 49) == [32]                 # small batch: one launch sequence
    assert engine.plan_sub_batches(64, 49) == [36, 28]
    for b, rows in [(1, 49), (63, 49), (64, 49), (100, 49), (129, 49), (200, 64), (255, 49), (384, 49), (4096, 49)]:
        for n in (None, 1, 2, 3, 5, 8):
            sizes = engine.plan_sub_batches(b, rows, n)
            assert sum(sizes) == b and all(s > 0 for s in sizes)
            if len(sizes) > 1:
                # every sub-batch but the last wastes less than one image worth of rows in its last tile pair
                for s in sizes[:-1]:
                    assert (-(s * rows) % 256) < rows
    monkeypatch.setenv("SD_SAMPLER_STREAMS", "2")
    assert engine.plan_sub_batches(256, 49) == [130, 126]


def test_tc_weight_layout_key_depends_on_tiles_not_on_concurrent_sub_batches():
    import ctypes
    from spiking_diffusion_b200 import _lib
    L = _lib.lib()

    def key(B, concurrent, C_in=512, C_out=256):
        d = _lib.ConvDesc()
        d.T, d.B, d.C_in, d.H_in, d.W_in, d.C_out, d.H_out, d.W_out = 4, B, C_in, 7, 7, C_out, 7, 7
        d.kh = d.kw = 3
        d.stride = d.pad = 1
        d.in_kind, d.out_kind, d.in_T, d.C_in0 = _lib.IN_STF, _lib.OUT_LIF, 4, C_in
        d.tau, d.v_threshold, d.v_reset, d.hard_reset, d.nsplit, d.concurrent = 2.0, 1.0, 0.0, 1, 2, concurrent
        return L.sd_conv_weight_layout_tc(ctypes.byref(d))

    full = key(256, 1)
    assert full > 0 and full & 1 == 1                      # N = 128, CTA pairs
    assert key(52, 5) == full                              # a concurrent sub-batch keeps the full-size tiles
    assert key(52, 1) == full                              # pairable tiles are never narrowed (an un-paired MMA costs
    assert key(1024, 1) == full                            # the same whatever N is)
    lone, shared = key(2, 1), key(2, 64)                   # a single M tile cannot be paired
    assert lone & 1 == 0 and shared & 1 == 0
    assert (lone >> 16) < 128 and (shared >> 16) == 128    # alone it is cut into narrower N tiles, concurrent it is not
    bad = _lib.ConvDesc()
    assert L.sd_conv_weight_layout_tc(ctypes.byref(bad)) == -1


def test_strided_layers_restated_for_the_tensor_core_kernel_are_identities():
    """The two algebraic restatements behind the tcgen05 VQ-VAE routes, checked with torch on the CPU:
    ConvTranspose2d(k3, s2, p1, op1) == Conv2d(k3, s1, p1) with transposed + flipped taps on the zero-inserted input;
    Conv2d(k3, s2, p1) == the even positions of Conv2d(k3, s1, p1)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    for H, W in ((7, 7), (8, 5)):
        convt = torch.nn.ConvTranspose2d(16, 32, 3, stride=2, padding=1, output_padding=1)
        assert engine._UpsampledConvT.eligible(convt)
        shim = engine._UpsampledConvT(convt)
        x = (torch.rand(3, 16, H, W, generator=g) < 0.3).float()
        up = torch.zeros(3, 16, 2 * H, 2 * W)
        up[..., ::2, ::2] = x
        ref = convt(x)
        got = F.conv2d(up, shim.weight, convt.bias, stride=1, padding=1)
        assert got.shape == ref.shape and float((got - ref).abs().max()) <= 1e-5
        conv = torch.nn.Conv2d(16, 32, 3, stride=2, padding=1)
        assert engine._Stride1Conv.eligible(conv)
        s1 = engine._Stride1Conv(conv)
        ref2 = conv(x)
        got2 = F.conv2d(x, s1.weight, s1.bias, stride=1, padding=1)[..., ::2, ::2]
        assert got2.shape == ref2.shape and float((got2 - ref2).abs().max()) <= 1e-5
    assert not engine._UpsampledConvT.eligible(torch.nn.ConvTranspose2d(16, 32, 3, stride=1, padding=1))
    assert not engine._UpsampledConvT.eligible(torch.nn.ConvTranspose2d(16, 3, 3, stride=2, padding=1, output_padding=1))
    assert not engine._Stride1Conv.eligible(torch.nn.Conv2d(1, 32, 3, stride=2, padding=1))


def test_lazy_memory_states_follow_the_memory_protocol():
    import copy
    import pickle
    n = neuron.LIFNode()
    assert n.memory_is_reset("v")
    calls = []

    def build():
        calls.append(1)
        return torch.arange(6.0).reshape(2, 3)

    n.v = base.LazyState(build)
    assert not n.memory_is_reset("v") and calls == []          # nothing is built until somebody looks
    assert torch.equal(n.v, torch.arange(6.0).reshape(2, 3)) and calls == [1]
    assert torch.equal(n.v, n.v) and calls == [1]               # built once, then a plain tensor
    n.v = base.LazyState(build)
    n.reset()
    assert n.memory_is_reset("v") and n.v == 0.0 and calls == [1]   # reset discards it without building
    n.v = base.LazyState(build)
    m = pickle.loads(pickle.dumps(n))                           # pickling materialises
    assert torch.equal(m.v, torch.arange(6.0).reshape(2, 3))
    n.v = base.LazyState(build)
    c = copy.deepcopy(n)
    assert torch.equal(c.v, torch.arange(6.0).reshape(2, 3))
    n.v = base.LazyState(build)
    assert [tuple(v.shape) for v in n.memories()] == [(2, 3)]


# ---- round-2 host-side checks ---------------------------------------------------------------------------
def test_memout_coefficients_bound_to_another_T_raise_on_the_fused_path():
    """engine._coef_array is what the fused plans use for the memout coefficients: a buffer bound to another T must
    raise the reference's broadcasting RuntimeError (SURVEY.md finding 1), never truncate or zero-pad."""
    from spiking_diffusion_b200 import engine
    from spiking_diffusion_b200.snn_model.snn_layers import MembraneOutputLayer
    coef16 = MembraneOutputLayer(16).coef
    arr = engine._coef_array(coef16, 16)
    assert len(arr) == 16 and abs(arr[15] - 1.0) < 1e-7 and abs(arr[14] - 0.8) < 1e-7
    for T in (4, 8, 32):
        with pytest.raises(RuntimeError, match=r"size of tensor a \(%d\) must match the size of tensor b \(16\)" % T):
            engine._coef_array(coef16, T)


def test_plan_caches_stay_out_of_deepcopy_and_pickle_and_key_covers_hyperparameters():
    import copy
    import pickle
    from spiking_diffusion_b200 import engine
    from spiking_diffusion_b200.snn_model import SNN_VQVAE, DummyModel, AbsorbingDiffusion
    den = DummyModel(1, 128, T=4)
    den._plans = {"key": 1, "plan": (lambda: 0)}            # a stand-in for ctypes pointers / CUDA graphs: not picklable
    ab = AbsorbingDiffusion(den, mask_id=128)
    ab._plans = {"key": 2, "plan": (lambda: 0)}
    vae = SNN_VQVAE(1, 16, 128, torch.tensor(1.0), T=4)
    vae._plans = {"key": 3, "plan": (lambda: 0)}
    for m in (den, ab, vae):
        c = copy.deepcopy(m)
        assert c._plans == {} and m._plans != {}
        assert pickle.loads(pickle.dumps(m))._plans == {}
        m.invalidate_plans()
        assert m._plans == {}
    k0 = engine.module_cache_key(den)
    den.conv3[2].tau = 4.0                                    # a LIF hyper-parameter baked into the plan
    k1 = engine.module_cache_key(den)
    den.conv2[1].eps = 1e-3                                   # BN eps is folded into scale/shift
    k2 = engine.module_cache_key(den)
    assert k0 != k1 and k1 != k2
    with torch.no_grad():
        den.conv2[0].weight.mul_(2.0)                         # in-place tensor op bumps _version
    assert engine.module_cache_key(den) != k2


def test_consumed_state_blocks_reads_until_reset():
    from spiking_diffusion_b200.activation_based import base, functional, neuron
    n = neuron.LIFNode()
    assert n.memory_is_reset("v")
    n.v = base.ConsumedState("test.node")
    assert not n.memory_is_reset("v")
    with pytest.raises(RuntimeError, match="consumed inside the fused kernels"):
        _ = n.v
    functional.reset_net(n)
    assert n.memory_is_reset("v") and n.v == 0.0


def test_device_helpers_without_gpu():
    from spiking_diffusion_b200 import _lib
    assert _lib.first_cuda_device(torch.zeros(2), [torch.zeros(1)], torch.nn.Linear(2, 2)) is None
    assert _lib.ptr(None) is None

    @_lib.on_device_of
    def f(x, k=1):
        return x + k
    assert float(f(torch.zeros(1), k=2)) == 2.0


def test_current_seq_output_kind_descriptor_and_layout():
    """SD_OUT_CURRENT_SEQ (the training branch's convolution on the int8 tensor-core path): accepted only with STF8 input and
    nsplit = 3, needs no workspace, and engine.currents_to_nchw reads the planar [T][C/8][R_alloc][8] buffer back."""
    import ctypes
    import torch
    from spiking_diffusion_b200 import _lib, engine
    L = _lib.lib()

    def desc(out_kind, nsplit=3, in_kind=None, C_in=64, C_out=48, T=4):
        d = _lib.ConvDesc()
        d.T, d.B, d.C_in, d.H_in, d.W_in, d.C_out, d.H_out, d.W_out = T, 3, C_in, 7, 7, C_out, 7, 7
        d.kh = d.kw = 3
        d.stride = d.pad = 1
        d.in_kind = _lib.IN_STF8 if in_kind is None else in_kind
        d.out_kind, d.in_T, d.C_in0 = out_kind, T, C_in
        d.tau, d.v_threshold, d.v_reset, d.hard_reset, d.nsplit, d.concurrent = 2.0, 1.0, 0.0, 1, nsplit, 1
        return d

    ok = desc(_lib.OUT_CURRENT_SEQ)
    assert L.sd_conv_tc_supported(ctypes.byref(ok)) == 1
    assert L.sd_conv_workspace_bytes(ctypes.byref(ok)) == 0              # the currents go straight to args.out
    assert L.sd_conv_weight_layout_tc(ctypes.byref(ok)) > 0
    assert L.sd_conv_tc_supported(ctypes.byref(desc(_lib.OUT_CURRENT_SEQ, nsplit=2, in_kind=_lib.IN_STF))) == 0
    assert L.sd_conv_tc_supported(ctypes.byref(desc(_lib.OUT_CURRENT_SEQ, T=3))) == 0      # int8 path: even T
    assert L.sd_conv_tc_supported(ctypes.byref(desc(_lib.OUT_CURRENT_SEQ, C_in=24))) == 0  # C_in % 32
    # layout: rows = guard + b*H*W + y*W + x, 8 channels innermost
    T, B, C, H, W = 2, 3, 16, 5, 4
    guard = (W + 1 + 7) // 8 * 8
    rows = L.sd_stf_bytes(T, B, C, H, W) // 2 // (T * (C // 8) * 8)
    want = torch.arange(T * B * C * H * W, dtype=torch.float32).reshape(T, B, C, H, W)
    buf = torch.full((T, C // 8, rows, 8), float("nan"))
    for t in range(T):
        for c in range(C):
            buf[t, c // 8, guard:guard + B * H * W, c % 8] = want[t, :, c].reshape(-1)
    got = engine.currents_to_nchw(buf.reshape(-1), T, B, C, H, W)
    assert torch.equal(got, want)
