"""(a12) sampling loop on the GPU.

The Philox stream is pinned against torch's own CUDA generator (the reference draws from it,
R/snn_model/vq_diffusion.py:105,118,136-138): torch.rand / Tensor.exponential_ / Categorical.sample with
torch.manual_seed must be reproduced bit for bit from (seed, offset).  Then the fused step kernel and the whole
sample() loop are compared with the CPU oracle driven by the same stream, and batch sharding is checked to
reproduce the single-GPU stream."""
import ctypes

import numpy as np
import pytest
import torch
import torch.distributions as dists

from conftest import make_denoiser, make_vqvae
from oracle import philox, snn_oracle as O
from spiking_diffusion_b200 import _lib
from spiking_diffusion_b200.snn_model.vq_diffusion import AbsorbingDiffusion

pytestmark = pytest.mark.gpu


def dev_info():
    L = _lib.lib()
    a, b, c, d = (ctypes.c_int() for _ in range(4))
    _lib.check(L.sd_device_info(a, b, c, d))
    return a.value, b.value


def ours(kind, numel, seed, offset, base=0, n=None):
    L = _lib.lib()
    n = numel - base if n is None else n
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    inc = ctypes.c_uint64()
    fn = L.sd_philox_uniform if kind == "u" else L.sd_philox_exponential
    _lib.check(fn(out.data_ptr(), n, seed, offset, base, numel, ctypes.byref(inc), _lib.stream_ptr()))
    return out, inc.value


@pytest.mark.parametrize("numel", [1, 49, 256 * 49, 303104 * 4 + 17, 256 * 49 * 128, 3_000_001])
def test_uniform_and_exponential_match_torch_cuda(numel):
    sms, thr = dev_info()
    torch.cuda.init()   # default_generators is empty until CUDA is initialised (this may be the first test to run)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(1234)
    assert gen.get_offset() == 0
    ref_u = torch.rand(numel, device="cuda")
    off1 = gen.get_offset()
    ref_e = torch.empty(numel, device="cuda").exponential_(1)
    off2 = gen.get_offset()
    u, inc_u = ours("u", numel, 1234, 0)
    e, inc_e = ours("e", numel, 1234, off1)
    assert inc_u == off1 and inc_e == off2 - off1           # generator bookkeeping identical to torch's
    assert torch.equal(u, ref_u)
    assert torch.equal(e, ref_e)
    # the numpy oracle is pinned by the same comparison: uniform exact; exponential to the accuracy of the device's
    # __logf (abs 2^-21.4 on [0.5, 2], 3 ulp elsewhere), which numpy's correctly rounded log cannot reproduce bitwise
    assert np.array_equal(philox.uniform(1234, 0, numel, sms, thr), ref_u.cpu().numpy())
    eo = philox.exponential(1234, off1, numel, sms, thr)
    assert np.allclose(eo, ref_e.cpu().numpy(), rtol=1e-6, atol=1e-6)
    # a shard of the stream equals the slice of the whole
    if numel > 1000:
        part, _ = ours("u", numel, 1234, 0, base=777, n=200)
        assert torch.equal(part, ref_u[777:977])


@pytest.mark.parametrize("K,temp", [(128, 1.0), (128, 0.65), (512, 0.3), (1024, 1.0), (32, 1.0), (96, 0.8), (200, 1.0),
                                    (1000, 0.65)])
def test_sample_step_matches_torch_categorical(K, temp):
    """Same logits, same seed: our fused step draws the same tokens as the reference's torch code path
    (rand_like -> Categorical(logits/temp).sample(), vq_diffusion.py:118-140)."""
    L = _lib.lib()
    b, h, w, t = 64, 7, 7, 5
    n = b * h * w
    g = torch.Generator().manual_seed(K)
    logits = (torch.randn((b, h, w, K), generator=g) * 2).cuda()
    x_t = torch.full((b, 1, h, w), K, dtype=torch.int64, device="cuda")
    unmasked = torch.zeros_like(x_t).bool()
    unmasked[::3] = True
    x_t[::3] = 7
    # reference code path on CUDA
    torch.manual_seed(99)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    tt = torch.full((b,), t, device="cuda", dtype=torch.long).reshape(b, 1, 1, 1).expand(b, 1, h, w)
    changes = torch.rand_like(x_t.float()) < 1 / tt.float()
    off_u = gen.get_offset()
    changes = torch.bitwise_xor(changes, torch.bitwise_and(changes, unmasked))
    unmasked_ref = torch.bitwise_or(unmasked, changes)
    x0_ref = dists.Categorical(logits=logits / temp).sample().long().unsqueeze(1)
    x_ref = x_t.clone()
    x_ref[changes] = x0_ref[changes]
    # ours
    xt = x_t.reshape(-1).clone()
    um = unmasked.reshape(-1).to(torch.uint8).clone()
    x0 = torch.empty(n, dtype=torch.int64, device="cuda")
    _lib.check(L.sd_sample_step(logits.data_ptr(), xt.data_ptr(), um.data_ptr(), x0.data_ptr(), n, K, t, temp, 99, 0,
                                off_u, 0, n, _lib.stream_ptr()))
    assert torch.equal(um.bool(), unmasked_ref.reshape(-1))
    mism = int((x0 != x0_ref.reshape(-1)).sum())
    # identical RNG stream; the only freedom is fp32 summation order inside softmax/logsumexp (~1e-7 relative),
    # which can move an argmax only on a near-tie: allow at most 2 of the b*49 draws
    assert mism <= 2, mism
    assert int((xt != x_ref.reshape(-1)).sum()) <= 2


def test_sample_loop_vs_oracle_and_sharding():
    T, K, b = 2, 128, 4
    den, sd = make_denoiser(T, K, seed=3)
    sms, thr = dev_info()
    ab = AbsorbingDiffusion(den, mask_id=K, shape=(7, 7), n_samples=b)
    x = ab.sample(temp=0.9, sample_steps=49, seed=5)
    assert x.shape == (b, 1, 7, 7) and x.dtype == torch.int64
    assert int(x.max()) < K and int(x.min()) >= 0                    # no mask token left (t=1 unmasks everything)
    # oracle driven by the same Philox stream: every image identical step by step, or its first divergence attributed
    # to a near-threshold neuron / near-tie draw (tests/_sample_parity.py; the benchmark-scale cases are in
    # tests/test_gpu_sample_parity.py)
    from _sample_parity import run_case
    rep = run_case("small_T2", den, sd, ab, T=T, K=K, hw=7, n_global=b, shard_base=0, check=list(range(b)), temp=0.9,
                   seed=5)
    assert rep["identical"] + len(rep["diverged"]) == b and rep["identical"] >= b - 1, rep
    plan = ab.plan(b)
    step_inc = plan.inc_u + plan.inc_e
    # reproducible, and consumes torch's generator like the reference's two draws per step
    torch.manual_seed(5)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    x2 = ab.sample(temp=0.9, sample_steps=49)
    assert torch.equal(x2, x) and gen.get_offset() == 49 * step_inc
    # sharding: images [0,2) and [2,4) generated separately equal the unsharded batch (no collective needed)
    ab2 = AbsorbingDiffusion(den, mask_id=K, shape=(7, 7), n_samples=2)
    lo = ab2.sample(temp=0.9, sample_steps=49, seed=5, n_global=b, shard_base=0)
    hi = ab2.sample(temp=0.9, sample_steps=49, seed=5, n_global=b, shard_base=2)
    assert torch.equal(torch.cat((lo, hi)), x)


def test_sample_then_decode_end_to_end():
    """The whole path of R/main.py:383-401 at a small batch: sample -> quantize -> poisson -> decoder -> uint8."""
    from spiking_diffusion_b200 import engine
    T, K, b = 4, 128, 8
    den, _ = make_denoiser(T, K, seed=0)
    vae, _ = make_vqvae(T, K, seed=0)
    ab = AbsorbingDiffusion(den, mask_id=K, shape=(7, 7), n_samples=b)
    sample = ab.sample(temp=0.65, sample_steps=49, seed=1).reshape(b, 7, 7)
    pred = vae.decode_indices(sample)
    img = engine.to_uint8(pred)
    assert img.shape == (b, 1, 28, 28) and img.dtype == torch.uint8
    assert 0 < float(pred.abs().max()) <= 1.0 and float(pred.std()) > 0.01


def test_chunked_large_batch_equals_one_plan():
    """n_samples above AbsorbingDiffusion.max_plan_batch is generated as consecutive chunks by one plan (one captured
    graph; the chunk's position in the global stream lives in device memory): the result must not depend on it."""
    T, K, b = 2, 128, 8
    den, _ = make_denoiser(T, K, seed=2)
    ab = AbsorbingDiffusion(den, mask_id=K, shape=(7, 7), n_samples=b)
    whole = ab.sample(temp=1.0, sample_steps=49, seed=21)
    ab.max_plan_batch = 3                        # chunks of 3, 3 and a ragged 2
    ab.invalidate_plans()
    parts = ab.sample(temp=1.0, sample_steps=49, seed=21)
    assert parts.shape == whole.shape and torch.equal(parts, whole)
    # also as a shard of a larger global batch
    ab.n_samples = 6
    lo = ab.sample(temp=1.0, sample_steps=49, seed=21, n_global=b, shard_base=0)
    assert torch.equal(lo, whole[:6])


def test_programmatic_dependent_launch_does_not_change_the_samples():
    """SD_PDL (csrc/common.cuh) lets the next kernel of the per-step chain start under the previous one's tail.  The knob is
    read once per process, so two fresh interpreters sample the same 70 images (two concurrent chains, T = 4 and the
    T-parallel T = 8 path) with it off and always on; tokens and decoded images must be identical."""
    import hashlib
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, hashlib, torch\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "from conftest import make_denoiser, make_vqvae\n"
        "from spiking_diffusion_b200.snn_model.vq_diffusion import AbsorbingDiffusion\n"
        "h = hashlib.sha256()\n"
        "for T, b in ((4, 70), (8, 12)):\n"
        "    den, _ = make_denoiser(T, 128, seed=3)\n"
        "    vae, _ = make_vqvae(T, 128, seed=3)\n"
        "    ab = AbsorbingDiffusion(den, mask_id=128, shape=(7, 7), n_samples=b)\n"
        "    x = ab.sample(temp=0.9, sample_steps=49, seed=11)\n"
        "    img = vae.decode_indices(x.reshape(b, 7, 7))\n"
        "    h.update(x.cpu().numpy().tobytes()); h.update(img.cpu().numpy().tobytes())\n"
        "print('DIGEST', h.hexdigest())\n")
    digests = []
    for mode in ("0", "2"):
        env = dict(os.environ, SD_PDL=mode)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        digests.append([l for l in r.stdout.splitlines() if l.startswith("DIGEST")][-1])
    assert digests[0] == digests[1], digests
