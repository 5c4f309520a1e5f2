"""Generate tests/golden/*.npz from the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):  python oracle/gen_golden.py
The reference is imported as-is (oracle/ref_loader.py) at its hard-coded T=16 / 7x7 / K=128, loaded with
the seeded synthetic parameters of ``spiking-diffusion_b200/synth.py`` (regenerable from the seed, so the
12 MB denoiser weights are not committed), and its outputs are stored compactly:
spikes as packed bits, indices as int16, images / logits as fp32.  The same run asserts that
``oracle/snn_oracle.py`` reproduces every stored tensor bit for bit.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader, snn_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def load_synth():
    spec = importlib.util.spec_from_file_location("sd_synth", os.path.join(ROOT, "spiking-diffusion_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def pack(x: torch.Tensor) -> np.ndarray:
    return np.packbits(x.detach().numpy().astype(np.uint8).reshape(-1))


def hook_lif_outputs(R, model):
    outs = []
    hs = []
    for m in model.modules():
        if isinstance(m, R.neuron.LIFNode):
            hs.append(m.register_forward_hook(lambda mod, i, o: outs.append(o.detach().clone())))
    return outs, hs


def kat(R):
    d = {}
    T = 8
    lif = R.neuron.LIFNode(tau=2.0, v_threshold=1.0, v_reset=0.0, step_mode="m").eval()
    xs = torch.tensor([0.5, 1.0, 1.5, 2.0, 3.0, -1.0, 0.999999, 1.9999999, 2.0000002])
    x_seq = xs[None, :].repeat(T, 1)
    d["lif_x"] = xs.numpy()
    d["lif_spikes"] = lif(x_seq).numpy()
    d["lif_v"] = lif.v.numpy()
    # random drive, hard and soft reset, second call continues from the stored state
    g = torch.Generator().manual_seed(7)
    xr = (torch.rand(6, 257, generator=g) - 0.3) * 3
    for name, vr in (("hard", 0.0), ("soft", None), ("hard_vr", -0.25)):
        n = R.neuron.LIFNode(tau=2.0, v_threshold=1.0, v_reset=vr, step_mode="m").eval()
        s1 = n(xr)
        s2 = n(xr.flip(0))
        d[f"lifr_{name}_s1"], d[f"lifr_{name}_s2"], d[f"lifr_{name}_v"] = s1.numpy(), s2.numpy(), n.v.numpy()
        s_o, v_o = O.lif_multi_step(xr, None, 2.0, 1.0, vr)
        s_o2, v_o2 = O.lif_multi_step(xr.flip(0), v_o, 2.0, 1.0, vr)
        assert torch.equal(s_o, s1) and torch.equal(s_o2, s2) and torch.equal(v_o2, n.v), name
    n = R.neuron.LIFNode(tau=3.0, v_threshold=0.7, v_reset=0.0, decay_input=False, step_mode="m").eval()
    d["lifr_tau3_s"] = n(xr).numpy()
    d["lifr_tau3_v"] = n.v.numpy()
    s_o, v_o = O.lif_multi_step(xr, None, 3.0, 0.7, 0.0, decay_input=False)
    assert torch.equal(s_o, torch.from_numpy(d["lifr_tau3_s"])) and torch.equal(v_o, n.v)
    d["lifr_x"] = xr.numpy()
    d["psp_in"] = np.array([1, 0, 0, 1], dtype=np.float32)
    d["psp_out"] = R.PSP()(torch.tensor([1.0, 0, 0, 1])[:, None]).reshape(-1).numpy()
    assert torch.equal(O.psp(torch.tensor([1.0, 0, 0, 1])[:, None]).reshape(-1), torch.from_numpy(d["psp_out"]))
    d["memout_coef16"] = R.MembraneOutputLayer().coef.reshape(-1).numpy()
    assert torch.equal(O.memout_coef(16).reshape(-1), torch.from_numpy(d["memout_coef16"]))
    # VQ tie-break: rows 3 and 5 of the codebook are equal and nearest -> index 3
    vq = R.VectorQuantizer(16, 8, 0.25)
    with torch.no_grad():
        vq.embeddings.weight.copy_(torch.arange(8 * 16).reshape(8, 16).float() / 10)
        vq.embeddings.weight[5] = vq.embeddings.weight[3]
    z = vq.embeddings.weight[3:4].detach().clone()
    d["tie_codebook"] = vq.embeddings.weight.detach().numpy()
    d["tie_idx"] = vq.get_code_indices(z).numpy()
    assert int(d["tie_idx"][0]) == 3
    np.savez_compressed(os.path.join(GOLD, "kat.npz"), **d)
    print("kat.npz written")


def vqvae(R, S, B=4, seed=0, K=128):
    T = 16
    sd = S.synth_vqvae_state(seed, num_embeddings=K, T=T)
    m = R.SNN_VQVAE(1, 16, K, torch.tensor(1.0))
    R.functional.set_step_mode(m, "m")
    m.load_state_dict(sd)
    m.eval()
    img = S.synth_images(seed, B)
    xs = img.unsqueeze(0).repeat(T, 1, 1, 1, 1)
    outs, hs = hook_lif_outputs(R, m)
    with torch.inference_mode():
        e, rec, idx = m(xs, img)
    for h in hs:
        h.remove()
    R.functional.reset_net(m)
    names = ["enc1", "enc2", "enc3", "gen", "dec1", "dec2"]
    assert len(outs) == 6
    tr = O.Trace()
    e_o, rec_o, idx_o = O.vqvae_forward_eval(xs, sd, trace=tr)
    assert torch.equal(e, e_o) and torch.equal(idx, idx_o) and torch.equal(rec, rec_o), "oracle != reference (vqvae)"
    d = {"seed": seed, "B": B, "T": T, "K": K, "idx": idx.numpy().astype(np.int16), "recon": rec.numpy()}
    for n, o in zip(names, outs):
        assert torch.equal(o, tr[n][0]), n
        d["spk_" + n] = pack(o)
        d["shape_" + n] = np.array(o.shape)
        d["near_" + n] = pack(O.spike_margin(tr[n][1]) < 1e-4)
        print(f"  {n}: rate {float(o.mean()):.4f}  near-threshold {int((O.spike_margin(tr[n][1]) < 1e-4).sum())}")
    flat = tr["feat"].reshape(-1, 16)
    d["vq_near"] = pack(O.vq_margin(flat, sd["vq_layer.embeddings.weight"]) < 1e-4)
    np.savez_compressed(os.path.join(GOLD, f"vqvae_T16_seed{seed}.npz"), **d)
    print("vqvae golden written; distinct codes", idx.unique().numel())


def denoiser(R, S, b=2, seed=0, K=128):
    T = 16
    sd = S.synth_denoiser_state(seed, num_embeddings=K)
    m = R.DummyModel(1, K)
    R.functional.set_step_mode(m, "m")
    m.load_state_dict(sd)
    m.eval()
    g = torch.Generator().manual_seed(4000 + seed)
    x = torch.randint(0, K, (b, 1, 7, 7), generator=g).float()
    x[torch.rand(b, 1, 7, 7, generator=g) < 0.5] = K
    t = torch.randint(1, 50, (b,), generator=g)
    outs, hs = hook_lif_outputs(R, m)
    with torch.inference_mode():
        lg = m(x, t)
    for h in hs:
        h.remove()
    R.functional.reset_net(m)
    tr = O.Trace()
    lg_o = O.denoiser_forward(x, t, sd, T, trace=tr)
    assert torch.equal(lg, lg_o), "oracle != reference (denoiser)"
    d = {"seed": seed, "b": b, "T": T, "K": K, "x": x.numpy(), "t": t.numpy(), "logits": lg.numpy()}
    for i, o in enumerate(outs):
        n = f"den{i + 1}"
        assert torch.equal(o, tr[n][0]), n
        d["spk_" + n] = pack(o)
        d["shape_" + n] = np.array(o.shape)
        d["near_" + n] = pack(O.spike_margin(tr[n][1]) < 1e-4)
        print(f"  {n}: rate {float(o.mean()):.4f}")
    np.savez_compressed(os.path.join(GOLD, f"denoiser_T16_seed{seed}.npz"), **d)
    print("denoiser golden written")


def decode(R, S, b=4, seed=0, K=128):
    """R/main.py:388-401 on fixed indices: quantize -> poisson -> decoder -> tanh(memout) -> uint8."""
    T = 16
    sd = S.synth_vqvae_state(seed, num_embeddings=K, T=T)
    m = R.SNN_VQVAE(1, 16, K, torch.tensor(1.0))
    R.functional.set_step_mode(m, "m")
    m.load_state_dict(sd)
    m.eval()
    g = torch.Generator().manual_seed(5000 + seed)
    sample = torch.randint(0, K, (b, 7, 7), generator=g)
    with torch.inference_mode():
        z = m.vq_layer.quantize(sample).permute(0, 3, 1, 2).contiguous()
        q = torch.unsqueeze(z, dim=0).repeat(16, 1, 1, 1, 1)
        q = m.vq_layer.poisson(q)
        pred = torch.tanh(m.memout(m.decoder(q)))
    R.functional.reset_net(m)
    u8 = np.array(np.clip((pred + 0.5).numpy(), 0., 1.) * 255, dtype=np.uint8)
    pred_o = O.decode_indices(sample, sd, T)
    assert torch.equal(pred, pred_o) and np.array_equal(O.to_uint8(pred_o).numpy(), u8)
    np.savez_compressed(os.path.join(GOLD, f"decode_T16_seed{seed}.npz"), seed=seed, b=b, T=T, K=K,
                        sample=sample.numpy().astype(np.int16), pred=pred.numpy(), u8=u8)
    print("decode golden written")


if __name__ == "__main__":
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    R = ref_loader.load()
    S = load_synth()
    os.makedirs(GOLD, exist_ok=True)
    kat(R)
    vqvae(R, S)
    denoiser(R, S)
    decode(R, S)
    os.system(f"ls -la {GOLD}")
