"""(a7) vector quantiser on the GPU vs the CPU oracle: integer output, bit-exact wherever the distance gap exceeds
1e-4 (north_star), exact-tie behaviour = lowest index."""
import numpy as np
import pytest
import torch

from conftest import SPIKE_MARGIN, golden, make_vqvae
from oracle import snn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K", [(3136, 128), (12544, 512), (7, 1024), (1, 128)])
def test_lookup_vs_oracle(M, K):
    m, sd = make_vqvae(4, K=K)
    g = torch.Generator().manual_seed(M + K)
    z = torch.rand((M, 16), generator=g) * (torch.rand((M, 16), generator=g) < 0.3) * 2.0
    idx = m.vq_layer.get_code_indices(z.cuda()).cpu()
    cb = sd["vq_layer.embeddings.weight"]
    ref = O.vq_code_indices(z, cb)
    margin = O.vq_margin(z, cb)
    assert idx.dtype == torch.int64 and idx.shape == (M,)
    bad = (idx != ref) & (margin > SPIKE_MARGIN)
    assert int(bad.sum()) == 0
    assert float((idx != ref).float().mean()) <= 1e-4


def test_exact_tie_returns_lowest_index():
    g = golden("kat.npz")
    m, _ = make_vqvae(4, K=128)
    from spiking_diffusion_b200.snn_model.vae_model import VectorQuantizer
    vq = VectorQuantizer(16, 8, 0.25, T=4).cuda().eval()
    with torch.no_grad():
        vq.embeddings.weight.copy_(torch.from_numpy(g["tie_codebook"]))
    z = torch.from_numpy(g["tie_codebook"][3:4]).cuda()
    assert int(vq.get_code_indices(z)[0]) == int(g["tie_idx"][0]) == 3
    # all-equal codebook: every distance ties -> index 0
    with torch.no_grad():
        vq.embeddings.weight.fill_(0.25)
    assert int(vq.get_code_indices(torch.rand(5, 16).cuda()).max()) == 0


def test_quantize_gather_and_errors():
    m, sd = make_vqvae(4)
    idx = torch.randint(0, 128, (3, 7, 7))
    q = m.vq_layer.quantize(idx.cuda())
    assert q.shape == (3, 7, 7, 16)
    assert torch.equal(q.cpu(), sd["vq_layer.embeddings.weight"][idx])
    with pytest.raises(IndexError):
        m.vq_layer.quantize(torch.tensor([128]).cuda())


@pytest.mark.parametrize("T", [4, 16])
def test_forward_feature_and_indices_vs_oracle(T):
    m, sd = make_vqvae(T)
    g = torch.Generator().manual_seed(T)
    x = (torch.rand((T, 5, 16, 7, 7), generator=g) < 0.15).float()
    spikes, idx = m.vq_layer(x.cuda())
    feat = m.vq_layer.feature(x.cuda()).cpu()
    e_ref, idx_ref, feat_ref = O.vq_forward_eval(x, sd)
    assert float((feat - feat_ref).abs().max()) <= 1e-6
    margin = O.vq_margin(feat_ref.reshape(-1, 16), sd["vq_layer.embeddings.weight"])
    assert int(((idx.cpu() != idx_ref) & (margin > SPIKE_MARGIN)).sum()) == 0
    assert spikes.shape == (T, 5, 16, 7, 7) and idx.shape == (5 * 49,)
    if torch.equal(idx.cpu(), idx_ref):
        assert float((spikes.cpu() != e_ref).float().mean()) <= 1e-4
