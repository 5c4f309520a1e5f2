"""Root-flip diagnosis: DummyModel.forward vs the oracle for nsplit 2 / 3 (GPU box): where is the first flipped layer,
how close to the threshold were the flipped neurons, how large is the pre-activation error implied."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_denoiser
from oracle import snn_oracle as O
from spiking_diffusion_b200.activation_based import functional

for (T, b, K, hw) in ((4, 5, 128, 8), (4, 8, 128, 7), (8, 4, 512, 7), (4, 64, 128, 7)):
    for ns in (2, 3):
        m, sd = make_denoiser(T, K, seed=2)
        m.nsplit = ns
        g = torch.Generator().manual_seed(b)
        x = torch.randint(0, K, (b, 1, hw, hw), generator=g).float()
        x[torch.rand(b, 1, hw, hw, generator=g) < 0.5] = K
        t = torch.randint(1, hw * hw + 1, (b,), generator=g)
        tr = O.Trace()
        lg_ref = O.denoiser_forward(x, t, sd, T, trace=tr)
        lg = m(x.cuda(), t.cuda()).cpu()
        plan = m.plan(b, hw, hw)
        bufs = {"den1": (plan.x1, plan.l1), "den2": (plan.x2, plan.l2), "den3": (plan.x3, plan.l3), "den4": (plan.x4, plan.l4), "den5": (plan.x5, plan.l5)}
        functional.reset_net(m)
        msg = f"T={T} b={b} K={K} hw={hw} nsplit={ns}: "
        for n, (bf, l) in bufs.items():
            got = plan.spikes_nchw(bf, l).cpu()
            s, h = tr[n]
            d = got != s
            if bool(d.any()):
                mg = (h - 1.0).abs()[d]
                msg += f"first flipped layer {n}: {int(d.sum())} flips, |h-1| of flipped: min {float(mg.min()):.3e} max {float(mg.max()):.3e}; "
                # smallest margins in this layer overall (how many neurons sit within 1e-6 / 1e-5 / 1e-4)
                allm = (h - 1.0).abs()
                msg += f"neurons within 1e-6/1e-5/1e-4: {int((allm<=1e-6).sum())}/{int((allm<=1e-5).sum())}/{int((allm<=1e-4).sum())} of {allm.numel()}"
                break
        else:
            msg += f"no flips; logits err {float((lg-lg_ref).abs().max()):.2e}"
        print(msg, flush=True)
