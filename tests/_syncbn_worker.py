"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU): a conv -> train-mode BN
block on a batch sharded over the ranks with statistics shared over NCCL must reproduce the single-GPU result on the
whole batch - outputs, running statistics, input gradient and (rank-averaged, as DDP does) parameter gradients."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spiking_diffusion_b200.activation_based import functional, layer  # noqa: E402


def block(seed):
    g = torch.Generator().manual_seed(seed)
    conv = layer.Conv2d(8, 16, 3, stride=1, padding=1, step_mode="m")
    bn = layer.BatchNorm2d(16, step_mode="m")
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.2)
        conv.bias.copy_(torch.randn(16, generator=g) * 0.1)
        bn.weight.copy_(torch.rand(16, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(16, generator=g) * 0.1)
    return torch.nn.Sequential(conv, bn).cuda().train()


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    T, B = 2, 4 * world
    g = torch.Generator().manual_seed(7)
    x_all = torch.randn(T, B, 8, 6, 6, generator=g)
    # a fixed random read-out: sum(y^2) would be (nearly) constant behind a BatchNorm and its gradient pure rounding noise
    r_all = torch.randn(T, B, 16, 6, 6, generator=g)
    lo, hi = rank * B // world, (rank + 1) * B // world
    # sharded, statistics over all ranks
    m = functional.convert_sync_batchnorm(block(3))
    xs = x_all[:, lo:hi].cuda().requires_grad_(True)
    y = m(xs)
    ((y * r_all[:, lo:hi].cuda()).mean()).backward()
    grads = {n: p.grad.clone() for n, p in m.named_parameters()}
    for v in grads.values():
        dist.all_reduce(v)
        v /= world
    # single GPU, whole batch
    ref = block(3)
    xr = x_all.cuda().requires_grad_(True)
    yr = ref(xr)
    ((yr * r_all.cuda()).mean()).backward()
    errs = {
        "y": rel(y.detach(), yr.detach()[:, lo:hi]),
        "gx": rel(xs.grad / world, xr.grad[:, lo:hi]),
        "running_mean": rel(m[1].running_mean, ref[1].running_mean),
        "running_var": rel(m[1].running_var, ref[1].running_var),
    }
    for n, p in ref.named_parameters():
        if n != "0.bias":   # a conv bias in front of a train-mode BN has no effect: its gradient is rounding noise
            errs["grad " + n] = rel(grads[n], p.grad)
    # DistributedDataParallel over the spiking denoiser in training mode (custom autograd functions, SyncBN inside):
    # one optimiser step; the averaged gradients must be finite and identical on every rank
    from spiking_diffusion_b200 import synth
    from spiking_diffusion_b200.snn_model.vq_diffusion import DummyModel
    den = DummyModel(1, 32, T=2)
    functional.set_step_mode(den, "m")
    den.load_state_dict(synth.synth_denoiser_state(0, n_channel=1, num_embeddings=32, num_timesteps=49))
    den = functional.convert_sync_batchnorm(den.cuda().train())
    ddp = torch.nn.parallel.DistributedDataParallel(den, device_ids=[int(os.environ["LOCAL_RANK"])])
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-3)
    gd = torch.Generator().manual_seed(100 + rank)
    xd = torch.randint(0, 33, (4, 1, 7, 7), generator=gd).float().cuda()
    td = torch.randint(1, 50, (4,), generator=gd).cuda()
    tgt = torch.randint(0, 32, (4, 7, 7), generator=gd).cuda()
    before = [p.detach().clone() for p in ddp.parameters()]
    loss = torch.nn.functional.cross_entropy(ddp(xd, td), tgt)
    opt.zero_grad(); loss.backward()
    flat = torch.cat([p.grad.flatten() for p in ddp.parameters()])
    other = flat.clone()
    dist.broadcast(other, src=0)
    errs["ddp grads differ across ranks"] = float((flat - other).abs().max())
    errs["ddp grads not finite"] = 0.0 if bool(torch.isfinite(flat).all()) and float(flat.abs().max()) > 0 else 1.0
    opt.step(); functional.reset_net(den)
    errs["ddp step left parameters unchanged"] = 0.0 if any(not torch.equal(a, b) for a, b in zip(before, ddp.parameters())) else 1.0
    bad = {k: v for k, v in errs.items() if not v <= 2e-5}
    t = torch.tensor([len(bad)], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("SYNCBN", "OK" if int(t) == 0 else "FAIL", {k: f"{v:.2e}" for k, v in errs.items()}, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 0 else 1)


if __name__ == "__main__":
    main()
