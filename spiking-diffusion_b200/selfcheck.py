"""Product self-checks that need several GPUs (run by bench.py under --gpus >= 2 and by tests/_syncbn_worker.py).

The sampling path has no exchange step; the optional training path has one: train-mode BatchNorm statistics (and the
parameter gradients, through DistributedDataParallel) travel over NCCL (SURVEY.md section 8(e),
functional.convert_sync_batchnorm).  ``syncbn_over_nccl`` proves on the ranks of an initialised process group that a
batch sharded over the ranks reproduces the single-GPU whole-batch result."""
import torch
import torch.distributed as dist

from .activation_based import functional, layer


def _block(seed):
    g = torch.Generator().manual_seed(seed)
    conv = layer.Conv2d(8, 16, 3, stride=1, padding=1, step_mode="m")
    bn = layer.BatchNorm2d(16, step_mode="m")
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.2)
        conv.bias.copy_(torch.randn(16, generator=g) * 0.1)
        bn.weight.copy_(torch.rand(16, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(16, generator=g) * 0.1)
    return torch.nn.Sequential(conv, bn).cuda().train()


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def syncbn_over_nccl(tol: float = 2e-5, with_ddp: bool = True):
    """Needs torch.distributed initialised with the nccl backend and the rank's device current.  Returns
    (ok over all ranks, {quantity: relative error}) -- conv -> train-mode SyncBN on a sharded batch vs one GPU on the
    whole batch: outputs, running statistics, input gradient, rank-averaged parameter gradients; then one
    DistributedDataParallel step of the spiking denoiser in training mode."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.cuda.current_device()
    T, B = 2, 4 * world
    g = torch.Generator().manual_seed(7)
    x_all = torch.randn(T, B, 8, 6, 6, generator=g)
    # a fixed random read-out: sum(y^2) would be (nearly) constant behind a BatchNorm and its gradient pure rounding noise
    r_all = torch.randn(T, B, 16, 6, 6, generator=g)
    lo, hi = rank * B // world, (rank + 1) * B // world
    m = functional.convert_sync_batchnorm(_block(3))
    xs = x_all[:, lo:hi].cuda().requires_grad_(True)
    y = m(xs)
    ((y * r_all[:, lo:hi].cuda()).mean()).backward()
    grads = {n: p.grad.clone() for n, p in m.named_parameters()}
    for v in grads.values():
        dist.all_reduce(v)
        v /= world
    ref = _block(3)
    xr = x_all.cuda().requires_grad_(True)
    yr = ref(xr)
    ((yr * r_all.cuda()).mean()).backward()
    errs = {
        "y": _rel(y.detach(), yr.detach()[:, lo:hi]),
        "gx": _rel(xs.grad / world, xr.grad[:, lo:hi]),
        "running_mean": _rel(m[1].running_mean, ref[1].running_mean),
        "running_var": _rel(m[1].running_var, ref[1].running_var),
    }
    for n, p in ref.named_parameters():
        if n != "0.bias":   # a conv bias in front of a train-mode BN has no effect: its gradient is rounding noise
            errs["grad " + n] = _rel(grads[n], p.grad)
    if with_ddp:
        # DistributedDataParallel over the spiking denoiser in training mode (custom autograd functions, SyncBN inside):
        # one optimiser step; the averaged gradients must be finite and identical on every rank
        from . import synth
        from .snn_model.vq_diffusion import DummyModel
        den = DummyModel(1, 32, T=2)
        functional.set_step_mode(den, "m")
        den.load_state_dict(synth.synth_denoiser_state(0, n_channel=1, num_embeddings=32, num_timesteps=49))
        den = functional.convert_sync_batchnorm(den.cuda().train())
        ddp = torch.nn.parallel.DistributedDataParallel(den, device_ids=[dev])
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
    bad = {k: v for k, v in errs.items() if not v <= tol}
    t = torch.tensor([len(bad)], device="cuda")
    dist.all_reduce(t)
    return int(t) == 0, errs
