import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_vqvae
from oracle import snn_oracle as O
from spiking_diffusion_b200 import engine, synth

T, B, K = 8, 32, 512
m, sd = make_vqvae(T, K, seed=1)
img = synth.synth_images(1, B)
xs = img.unsqueeze(0).repeat(T, 1, 1, 1, 1)
tr = O.Trace()
O.vqvae_forward_eval(xs, sd, trace=tr)
plan = m.plan(T, B, 28, 28)
plan.forward(img.cuda(), const_over_T=True)
gen_ref = tr["gen"][0]
# run dec1 alone on the oracle's generator spikes
stf = engine.stf_from_nchw(gen_ref.cuda())
out = plan.d1.alloc_out()
plan.d1.run(stf, out)
got = engine.stf_to_nchw(out, T, B, 64, 14, 14).cpu()
s_ref, h = tr["dec1"]
diff = got != s_ref
marg = O.spike_margin(h)
print("generic" if os.environ.get("SD_SIMT_GENERIC") else "spike8", "flips", int(diff.sum()), "margins of flips:", marg[diff].tolist()[:10])
idx = diff.nonzero()[:6]
for i in idx:
    t, b, c, y, x = [int(v) for v in i]
    print("  at t,b,c,y,x", t, b, c, y, x, "h_ref over time:", [round(float(h[tt, b, c, y, x]), 6) for tt in range(T)],
          "ref spikes", s_ref[:, b, c, y, x].tolist(), "ours", got[:, b, c, y, x].tolist())
# BN scale of the offending channels
q = "decoder.snn_convs."
sc = sd[q + "1.weight"] / torch.sqrt(sd[q + "1.running_var"] + 1e-5)
print("BN scale range", float(sc.min()), float(sc.max()), "scale of flipped channels", [round(float(sc[int(i[2])]), 2) for i in idx])
