"""(a1)(a3)(a6) LIF neuron, state protocol and MembraneOutputLayer on the GPU vs the CPU oracle and the reference's
known-answer vectors.  LIF arithmetic is exact (fp32, same operation order): the bar is bit-exact."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import snn_oracle as O
from spiking_diffusion_b200.activation_based import functional, neuron
from spiking_diffusion_b200.snn_model.snn_layers import MembraneOutputLayer

pytestmark = pytest.mark.gpu


def test_known_answers_from_reference():
    g = golden("kat.npz")
    n = neuron.LIFNode(tau=2.0, v_threshold=1.0, v_reset=0.0, step_mode="m").eval()
    x = torch.from_numpy(g["lif_x"])[None, :].repeat(8, 1).cuda()
    s = n(x)
    assert torch.equal(s.cpu(), torch.from_numpy(g["lif_spikes"]))
    assert torch.equal(n.v.cpu(), torch.from_numpy(g["lif_v"]))


@pytest.mark.parametrize("name,vr", [("hard", 0.0), ("soft", None), ("hard_vr", -0.25)])
def test_reset_variants_and_state_carry(name, vr):
    g = golden("kat.npz")
    xr = torch.from_numpy(g["lifr_x"]).cuda()
    n = neuron.MultiStepLIFNode(tau=2.0, v_threshold=1.0, v_reset=vr).eval()
    assert isinstance(n.v, float)
    s1 = n(xr)
    assert isinstance(n.v, torch.Tensor) and n.v.shape == xr.shape[1:]
    s2 = n(xr.flip(0))   # continues from the stored membrane potential
    assert torch.equal(s1.cpu(), torch.from_numpy(g[f"lifr_{name}_s1"]))
    assert torch.equal(s2.cpu(), torch.from_numpy(g[f"lifr_{name}_s2"]))
    assert torch.equal(n.v.cpu(), torch.from_numpy(g[f"lifr_{name}_v"]))
    functional.reset_net(n)
    assert isinstance(n.v, float)
    assert torch.equal(n(xr).cpu(), torch.from_numpy(g[f"lifr_{name}_s1"]))


def test_no_decay_input_tau3():
    g = golden("kat.npz")
    n = neuron.LIFNode(tau=3.0, v_threshold=0.7, v_reset=0.0, decay_input=False, step_mode="m").eval()
    s = n(torch.from_numpy(g["lifr_x"]).cuda())
    assert torch.equal(s.cpu(), torch.from_numpy(g["lifr_tau3_s"]))
    assert torch.equal(n.v.cpu(), torch.from_numpy(g["lifr_tau3_v"]))


@pytest.mark.parametrize("shape", [(4, 64, 16, 7, 7), (8, 3, 5, 7), (16, 2, 1, 28, 28), (4, 1, 1), (1, 1000003)])
def test_random_bit_exact_vs_oracle(shape):
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(shape, generator=g) - 0.3) * 3
    s_ref, v_ref, h_ref = O.lif_multi_step(x, return_h=True)
    n = neuron.LIFNode(step_mode="m", store_v_seq=True).eval()
    s = n(x.cuda())
    assert s.shape == x.shape and s.dtype == torch.float32
    assert torch.equal(s.cpu(), s_ref) and torch.equal(n.v.cpu(), v_ref)
    v_seq_ref = 0.0 * s_ref + (1.0 - s_ref) * h_ref
    assert torch.equal(n.v_seq.cpu(), v_seq_ref)


def test_single_step_mode_and_errors():
    n = neuron.LIFNode().eval()           # step_mode 's'
    x = torch.full((5, 3), 1.5).cuda()
    assert n(x).sum() == 0 and n(x).sum() == 15   # 0.75 then 1.125 >= 1
    with pytest.raises(ValueError):
        n.step_mode = "x"
    with pytest.raises(AssertionError):
        neuron.LIFNode(tau=1.0)
    m = neuron.LIFNode(step_mode="m").eval()
    m(torch.zeros(2, 4, 4).cuda())
    with pytest.raises(ValueError):
        m(torch.zeros(2, 5, 4).cuda())      # stale state of another shape: reset() required, as in the reference
    assert m(torch.zeros(0, 4, 4).cuda()).shape == (0, 4, 4)  # empty sequence is a no-op


def test_state_moves_with_module():
    n = neuron.LIFNode(step_mode="m").eval()
    n(torch.rand(2, 8).cuda())
    n.cpu()
    assert n.v.device.type == "cpu"
    n.cuda()
    assert n.v.is_cuda


@pytest.mark.parametrize("T", [4, 8, 16])
def test_memout_matches_oracle(T):
    g = torch.Generator().manual_seed(5)
    x = torch.rand((T, 3, 2, 9, 9), generator=g) - 0.5
    m = MembraneOutputLayer(T).cuda()
    out = m(x.cuda())
    ref = O.memout(x)
    assert float((out.cpu() - ref).abs().max()) <= 1e-6
    assert float((m(x.cuda(), apply_tanh=True).cpu() - torch.tanh(ref)).abs().max()) <= 1e-6
    with pytest.raises(RuntimeError):
        MembraneOutputLayer(16).cuda()(torch.zeros(4, 1, 1, 2, 2).cuda())   # SURVEY.md finding 1: T-bound coef


@pytest.mark.parametrize("vr,detach,decay", [(0.0, False, True), (None, False, True), (-0.5, True, True), (0.0, False, False),
                                             (None, True, False)])
def test_surrogate_gradient_bptt_matches_oracle(vr, detach, decay):
    """(a2 / f.1) training branch: spikes identical to the eval kernel, gradients by surrogate BPTT (ATan, alpha=2)
    vs torch autograd through the oracle's restatement.  Tolerance: fp32, rtol 1e-5."""
    from spiking_diffusion_b200.activation_based import surrogate
    g = torch.Generator().manual_seed(9)
    x_cpu = (torch.rand(6, 3, 16, 7, 7, generator=g) - 0.3) * 3
    w_cpu = torch.rand(6, 3, 16, 7, 7, generator=g)
    x_ref = x_cpu.clone().requires_grad_(True)
    s_ref, v_ref = O.lif_multi_step_train(x_ref, None, 2.0, 1.0, vr, decay, detach, 2.0)
    ((s_ref * w_cpu).sum() + (v_ref * 0.3).sum()).backward()
    n = neuron.LIFNode(tau=2.0, decay_input=decay, v_threshold=1.0, v_reset=vr, surrogate_function=surrogate.ATan(),
                       detach_reset=detach, step_mode="m").train()
    x = x_cpu.cuda().requires_grad_(True)
    s = n(x)
    ((s * w_cpu.cuda()).sum() + (n.v * 0.3).sum()).backward()
    assert torch.equal(s.detach().cpu(), s_ref.detach())
    assert torch.allclose(x.grad.cpu(), x_ref.grad, rtol=1e-5, atol=1e-7)
    # a second call continues from (and back-propagates into) the stored state, like the reference
    s2 = n(x)
    assert s2.requires_grad and n.v.requires_grad
