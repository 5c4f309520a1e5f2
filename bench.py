#!/usr/bin/env python
"""Benchmark of the Spiking-Diffusion hot path: generated images/sec (BASELINE.json:metric).

A *step* is one pass of the whole path over one batch: AbsorbingDiffusion.sample (h*w reverse-diffusion steps, each
a full T-timestep spiking denoiser forward + categorical draw + unmask update) followed by the decode of
R/main.py:388-401 (quantize -> spike generator -> spiking decoder -> tanh(memout) -> uint8).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|ref16] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Our arm.  Headline = BASELINE.json configs[1] (cfg2: MNIST-shape sampling, b=256 per GPU, T=4, K=128), weak-scaled over
the ranks: `value` = images/s with everything resident in HBM, device-timed (CUDA events, max over ranks); `e2e` = the
same through the public API with host buffers (pinned H2D of the initial token grid, D2H of the uint8 images inside
the timed region).  The same JSON line carries `secondary` records for the other BASELINE configs: cfg3 (CIFAR-shape,
b=1024 per GPU), cfg4 (T=8, K=512, a GLOBAL batch of 4096 split over the N ranks) and the batch sweep
{64, 1024, 4096, 16384, 65536} images per GPU at cfg2's shapes, each with value / ms / clocks / dominant-kernel
fraction; `roofline` holds the dominant kernel timed in isolation against the BURST peak, the whole step against the
SUSTAINED peak, and HBM entries for the stand-alone LIF and sampling-step kernels.  Under --gpus >= 2 the ranks also run
the SyncBN-over-NCCL self-check of the optional training path (`syncbn_check`).

`--impl reference` (rank 0 only) times the reference's algorithm on the host CPU: `value` = the oracle port at the
headline workload's T / K on a bounded sample per step; the line also reports one full-batch step of the port and the
UNMODIFIED reference from baseline/_ref (oracle/install_ref.py) at its hard-coded T=16 / 16 images (`reference_as_shipped`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1..3]
    "cfg2": dict(desc="MNIST-shape 28x28 sampling, b=256 per GPU, T=4, K=128, 49 steps + decode", b=256, T=4, K=128,
                 hw=7, in_dim=1),
    "cfg3": dict(desc="CIFAR-10-shape 3x32x32 sampling, b=1024 per GPU, T=4, K=128, 64 steps + decode", b=1024, T=4,
                 K=128, hw=8, in_dim=3),
    "cfg4": dict(desc="KMNIST/Letters-shape 28x28 sampling, T=8, K=512, 49 steps + decode, global batch 4096 split over "
                      "the ranks", b=4096, T=8, K=512, hw=7, in_dim=1, strong=True),
    # the reference as shipped: T=16 (hard-coded), 16 samples per sample() call x 2 (R/main.py:383-387)
    "ref16": dict(desc="reference as shipped: 28x28 sampling, b=32, T=16, K=128, 49 steps + decode", b=32, T=16, K=128,
                  hw=7, in_dim=1),
}
SWEEP = (64, 1024, 4096, 16384, 65536)     # BASELINE.json configs[4]: images per GPU at cfg2's shapes
CHUNK = 4096                               # larger batches run as consecutive chunks of one plan (one global stream)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML every 10 ms in a thread; nvidia-smi as a
    fallback).  `reasons` lists the throttle reasons seen active in any sample."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.sm, self.mask, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._nvml = pynvml
        except Exception:  # noqa: BLE001
            self._nvml = None
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        return self

    def _sample(self):
        if self._nvml is not None:
            n = self._nvml
            self.sm.append(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
            try:
                self.mask |= n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
            except Exception:  # noqa: BLE001
                self.mask |= n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        else:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
            r = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5).stdout.split(",")
            self.sm.append(int(r[0])); self.max_mhz = int(r[1]); self.mask |= int(r[2].strip(), 16)

    def _loop(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(v for k, v in self.REASONS.items() if self.mask & k), "samples": len(sm)}


def base_config(args, wl, world):
    """The workload description shared verbatim by our arm and the reference arm (arm-specific details go elsewhere)."""
    per_gpu = wl["b"] // world if wl.get("strong") else wl["b"]
    return {"workload": f"{args.workload}: {wl['desc']}", "global_batch": per_gpu * world, "temp": args.temp,
            "parallelism": f"batch-shard x{world}, no collective on the sampling path"}


# ----------------------------------------------------------------------------------------------------------
def build_models(wl, device):
    import torch
    from spiking_diffusion_b200 import synth
    from spiking_diffusion_b200.activation_based import functional
    from spiking_diffusion_b200.snn_model import SNN_VQVAE, DummyModel, AbsorbingDiffusion
    T, K = wl["T"], wl["K"]
    vsd = synth.synth_vqvae_state(0, in_dim=wl["in_dim"], num_embeddings=K, T=T)
    dsd = synth.synth_denoiser_state(0, n_channel=1, num_embeddings=K, num_timesteps=wl["hw"] ** 2)
    vae = SNN_VQVAE(wl["in_dim"], 16, K, torch.tensor(1.0), T=T)
    den = DummyModel(1, K, T=T)
    functional.set_step_mode(vae, "m"); functional.set_step_mode(den, "m")
    vae.load_state_dict(vsd); den.load_state_dict(dsd)
    vae, den = vae.eval().to(device), den.eval().to(device)
    ab = AbsorbingDiffusion(den, mask_id=K, shape=(wl["hw"], wl["hw"]), n_samples=wl["b"])
    return vae, den, ab, vsd, dsd


# ---- CPU legs (the only places that execute oracle/) ------------------------------------------------------
def cpu_port_images_per_s(wl, b_cpu, steps=1, warm_steps=3):
    """The reference's algorithm (oracle port, torch CPU fp32, all host threads) on a bounded sample of the workload:
    sample() for `b_cpu` images + decode.  Returns (images/s, cores, step times, description)."""
    import torch
    from oracle import philox, snn_oracle as O
    from spiking_diffusion_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T, K, hw = wl["T"], wl["K"], wl["hw"]
    vsd = synth.synth_vqvae_state(0, in_dim=wl["in_dim"], num_embeddings=K, T=T)
    dsd = synth.synth_denoiser_state(0, n_channel=1, num_embeddings=K, num_timesteps=hw * hw)
    uni = lambda step, n: torch.from_numpy(philox.uniform(0, step * 12, n, 148, 2048))
    expo = lambda step, rows, k: torch.from_numpy(philox.exponential(0, step * 12 + 4, rows * k, 148, 2048)).reshape(rows, k)
    with torch.inference_mode():
        x = torch.full((min(b_cpu, 16), 1, hw, hw), float(K))
        for _ in range(warm_steps):
            O.denoiser_forward(x, torch.full((x.shape[0],), 1), dsd, T)
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            tok = O.sample(dsd, T, b_cpu, (hw, hw), K, K, 1.0, hw * hw, uni, expo)
            O.to_uint8(O.decode_indices(tok.reshape(b_cpu, hw, hw), vsd, T))
            times.append(time.perf_counter() - t0)
    return b_cpu / min(times), cores, times, (f"oracle port of the reference (torch {torch.__version__} CPU fp32), sample()+decode of "
                                              f"{b_cpu} images per step, {hw * hw} diffusion steps, T={T}, K={K}, best of {steps}")


def reference_as_shipped(device="cpu", n_calls=1):
    """The UNMODIFIED reference (baseline/_ref): DummyModel + AbsorbingDiffusion.sample at its hard-coded T=16, 7x7,
    16 images per call (R/snn_model/vq_diffusion.py:51,103-142), then the caller-side decode of R/main.py:388-401
    (restated here because main.py itself cannot be imported: matplotlib / torchmetrics are absent).  On the CPU its
    'cuda' literal is redirected (oracle/ref_loader.py:redirect_cuda_to_cpu); on 'cuda' it runs exactly as shipped
    (eager PyTorch, cuDNN/cuBLAS).  Returns a record or None when no reference install is present."""
    import numpy as np
    import torch
    from oracle import ref_loader
    from spiking_diffusion_b200 import synth
    if not ref_loader.available():
        return None
    ns = ref_loader.load()
    cores = os.cpu_count() or 1
    if device == "cpu":
        torch.set_num_threads(cores)
    vae = ns.SNN_VQVAE(1, 16, 128, torch.tensor(1.0))
    den = ns.DummyModel(1, 128)
    ns.functional.set_step_mode(vae, "m"); ns.functional.set_step_mode(den, "m")
    vae.load_state_dict(synth.synth_vqvae_state(0, num_embeddings=128, T=16))
    den.load_state_dict(synth.synth_denoiser_state(0, n_channel=1, num_embeddings=128, num_timesteps=49))
    vae, den = vae.eval().to(device), den.eval().to(device)
    ab = ns.AbsorbingDiffusion(den, mask_id=128)          # n_samples = 16, 7x7, as shipped

    def one_call():
        sample = ab.sample(temp=1.0, sample_steps=49).reshape(-1, 7, 7)                    # main.py:384-387
        quantized = vae.vq_layer.quantize(sample).permute(0, 3, 1, 2)                      # main.py:389-391
        quantized = vae.vq_layer.poisson(torch.unsqueeze(quantized, dim=0).repeat(16, 1, 1, 1, 1))
        pred = torch.tanh(vae.memout(vae.decoder(quantized)))                              # main.py:398-399
        ns.functional.reset_net(vae)
        img = (np.clip(pred.cpu().numpy() + 0.5, 0, 1) * 255).astype(np.uint8)             # main.py:401
        return img

    import contextlib
    ctx = ref_loader.redirect_cuda_to_cpu(ns.vq_diffusion) if device == "cpu" else contextlib.nullcontext()
    times = []
    with torch.inference_mode(), ctx:
        if device != "cpu":
            one_call()                                    # warm-up: cuDNN autotune, lazy init
            torch.cuda.synchronize()
        for _ in range(n_calls):
            t0 = time.perf_counter()
            img = one_call()
            if device != "cpu":
                torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    assert img.shape == (16, 1, 28, 28)
    return {"value": round(16 / min(times), 4), "unit": "images/s", "kind": "reference", "device": device,
            "cores": cores if device == "cpu" else None, "source": ref_loader.source(),
            "sample": f"unmodified reference: AbsorbingDiffusion.sample(temp=1, 49 steps) + decode, T=16 hard-coded, "
                      f"16 images per call, best of {n_calls} call(s), torch {torch.__version__} on {device}"}


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    t0 = time.perf_counter()
    b_cpu = args.cpu_batch
    ips, cores, times, sample = cpu_port_images_per_s(wl, b_cpu, steps=max(1, args.steps), warm_steps=max(1, args.warmup))
    cpu = {"value": round(ips, 4), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
    if not args.quick:
        # one step on the FULL per-GPU batch of the workload (same configuration as our arm, not a sample of it)
        full_b = wl["b"] // world if wl.get("strong") else wl["b"]
        full_b = min(full_b, 256)
        fi, _, ft, fs = cpu_port_images_per_s(wl, full_b, steps=1, warm_steps=0)
        cpu["full_batch"] = {"value": round(fi, 4), "images": full_b, "seconds": round(ft[0], 2), "sample": fs}
    shipped = None
    if not args.quick:
        try:
            shipped = reference_as_shipped("cpu", n_calls=1)
        except Exception as e:  # noqa: BLE001
            shipped = {"unavailable": f"{type(e).__name__}: {e}"}
    line = {
        "impl": "reference", "metric": "generated images/sec", "value": round(ips, 4), "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * min(times), 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args, wl, world),
        "cpu_baseline": cpu,
        "reference_as_shipped": shipped,
        "e2e": {"value": round(ips, 4), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 1),
    }
    print(json.dumps(line), flush=True)


# ---- our arm ------------------------------------------------------------------------------------------------
class Runner:
    """One workload on one rank: models, sampler plan (chunked above CHUNK images), decode plan, timing helpers."""

    def __init__(self, wl, b, n_global, shard_base, dev, nsplit, temp):
        import torch
        self.torch, self.wl, self.b, self.dev, self.temp = torch, wl, b, dev, temp
        self.n_global, self.shard_base = n_global, shard_base
        self.vae, self.den, self.ab, _, _ = build_models(dict(wl, b=min(b, CHUNK)), dev)
        self.den.nsplit = nsplit
        self.T, self.K, self.hw = wl["T"], wl["K"], wl["hw"]
        self.steps_diff = self.hw * self.hw
        self.chunk = min(b, CHUNK)
        self.n_chunks = -(-b // self.chunk)
        if b % self.chunk:
            raise ValueError("bench batches are multiples of the chunk size")
        self.splan = self.ab.plan(self.chunk, n_global, shard_base)
        self.vplan = self.vae.plan(self.T, self.chunk, 4 * self.hw, 4 * self.hw)
        self.img8 = torch.empty((self.chunk, wl["in_dim"], 4 * self.hw, 4 * self.hw), dtype=torch.uint8, device=dev)

    def one_pass(self, seed):
        from spiking_diffusion_b200 import _lib
        for c in range(self.n_chunks):
            tok = self.splan.sample(self.temp, self.steps_diff, seed, 0, extra_base=c * self.chunk)
            pred = self.vplan.decode_indices(tok)
            _lib.check(_lib.lib().sd_to_uint8(pred.data_ptr(), self.img8.data_ptr(), pred.numel(), _lib.stream_ptr()))

    def launches_per_pass(self):
        return self.n_chunks * (self.steps_diff * self.splan.kernel_launches_per_step + 5 + 1)

    def flops_per_image(self, executed=False):
        """Dense un-split algorithmic FLOPs per generated image; `executed`: the read-out layer counted
        as executed (once on the T-summed spikes instead of T times)."""
        sp = self.splan
        per = 0
        for dp, _, _ in sp.subs:
            for l in dp.layers:
                f = l.flops()
                if executed and l is dp.l6:
                    f //= dp.T
                per += f
        return per * self.steps_diff // sp.b + self.vplan.flops() // self.chunk

    def timed(self, steps, warmup, flush, barrier, seed0=0):
        """Device-timed passes (CUDA events per step on the launching stream, L2 flushed between steps)."""
        torch = self.torch
        for i in range(warmup):
            self.one_pass(1000 + i)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for i in range(steps):
            flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
            ev[i][0].record()
            self.one_pass(seed0 + i)
            ev[i][1].record()
        barrier()
        return sum(a.elapsed_time(b_) for a, b_ in ev)

    def layer_groups(self, flush, reps=14, skip=3):
        """Every tcgen05 layer timed as it runs in the step: one 'launch group' = the layer's launches of all
        sub-batches, forked onto their streams from one start event and joined into one stop event, ALONE on the GPU
        with a flushed L2 (so: an isolated-kernel figure, to be read against the burst peak)."""
        torch, splan = self.torch, self.splan
        names = [("den.conv2", "l2", "x1", "x2", None, None), ("den.conv3", "l3", "x2", "x3", None, None),
                 ("den.conv4", "l4", "x3", "x4", None, None), ("den.conv5", "l5", "x4", "x5", "x5s", None),
                 ("den.conv6", "l6", "x5s", "logits", None, "x1s")]
        # each group is captured once into a CUDA graph (fork onto the sub-batch streams, join) and replayed between two
        # events, so the figure is GPU time, not the Python launch path
        graphs = {}
        for n, ln, xi, xo, xs, x2 in names:
            def enqueue(ln=ln, xi=xi, xo=xo, xs=xs, x2=x2):
                cur = torch.cuda.current_stream()
                start = torch.cuda.Event()
                start.record(cur)
                for (dp, _, _), st in zip(splan.subs, splan.streams):
                    run = lambda: getattr(dp, ln).run(getattr(dp, xi), getattr(dp, xo), x2=getattr(dp, x2) if x2 else None,
                                                      out_sum=getattr(dp, xs) if xs else None)
                    if st is None:
                        run()
                    else:
                        st.wait_event(start)
                        with torch.cuda.stream(st):
                            run()
                        done = torch.cuda.Event()
                        done.record(st)
                        cur.wait_event(done)
            enqueue()                      # eager once: lazy one-time initialisation must not happen under capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                enqueue()
            graphs[n] = g
        acc = {n: [] for n, *_ in names}
        for rep in range(reps):
            flush.zero_()
            for n, *_ in names:
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                graphs[n].replay()
                c.record()
                acc[n].append((a, c))
        torch.cuda.synchronize()
        per = {}
        for n, ln, *_ in names:
            t = sorted(a.elapsed_time(c) for a, c in acc[n][skip:])
            mean_ms = sum(t) / len(t)
            lyrs = [getattr(dp, ln) for dp, _, _ in splan.subs]
            fl = sum(l.flops() for l in lyrs)
            ex = sum(l.executed_tensor_ops() for l in lyrs)
            per[n] = dict(ms=round(mean_ms, 4), tflops=round(fl / mean_ms / 1e9, 1),
                          executed_tensor_tops=round(ex / mean_ms / 1e9, 1), mma_kind=lyrs[0].mma_kind(),
                          impl=lyrs[0].impl, launches_per_group=len(splan.subs))
        return per


def hbm_kernels(dev, flush, pk):
    """HBM-bound kernels of the path, timed alone: stand-alone LIFNode.forward on cfg2's largest layer output
    ([4,256,512,7,7], 8 B per neuron-timestep + 8 B per neuron) and one sampling step (logits read + token/mask
    read-write; the kernel is Philox-ALU bound, the fraction says how far from the HBM bound)."""
    import torch
    from spiking_diffusion_b200 import _lib
    L = _lib.lib()
    out = {}
    T, N = 4, 256 * 512 * 49
    x = torch.randn((T, N), device=dev) * 1.5
    v = torch.zeros(N, device=dev)
    s = torch.empty_like(x)
    ts = []
    for _ in range(8):
        flush.zero_(); v.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(L.sd_lif_forward(x.data_ptr(), v.data_ptr(), s.data_ptr(), None, T, N, 2.0, 1.0, 0.0, 1, 1, _lib.stream_ptr()))
        c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    by = 8 * T * N + 8 * N
    out["lif_vec4_kernel"] = {"bound": "hbm", "shape": "[4,256,512,7,7]", "ms": round(ms, 4), "achieved": round(by / ms / 1e6, 1),
                              "peak": pk["hbm"], "unit": "GB/s", "frac": round(by / ms / 1e6 / pk["hbm"], 4),
                              "algorithmic_bytes": by}
    n, K = 256 * 49, 128
    logits = torch.randn((n, K), device=dev)
    xt = torch.full((n,), K, dtype=torch.int64, device=dev)
    um = torch.zeros(n, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(8):
        flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(L.sd_sample_step(logits.data_ptr(), xt.data_ptr(), um.data_ptr(), None, n, K, 5, 1.0, 1, 0, 12, 0, n, _lib.stream_ptr()))
        c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    by = n * K * 4 + n * (8 + 1) * 2
    out["sample_step_kernel"] = {"bound": "hbm", "shape": f"{n} tokens x K={K}", "ms": round(ms, 4),
                                 "achieved": round(by / ms / 1e6, 1), "peak": pk["hbm"], "unit": "GB/s",
                                 "frac": round(by / ms / 1e6 / pk["hbm"], 4), "algorithmic_bytes": by,
                                 "note": "K Philox4x32-10 evaluations per token: ALU-bound, not HBM-bound"}
    return out


def run_ours(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from spiking_diffusion_b200 import engine, _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    pk = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def shard(w):
        """(per-rank batch, global batch, first image of this rank) of a workload: weak scaling unless it says strong."""
        per = w["b"] // world if w.get("strong") else w["b"]
        return per, per * world, per * rank

    # ---- headline --------------------------------------------------------------------------------------------
    b, n_global, shard_base = shard(wl)
    run = Runner(wl, b, n_global, shard_base, dev, args.nsplit, args.temp)
    hw, K = run.hw, run.K
    clocks = ClockSampler(local_rank).start()
    total_ms = max_over_ranks(run.timed(args.steps, args.warmup, flush, barrier))
    value = n_global * args.steps / (total_ms / 1e3)

    # ---- end to end through the public API with host buffers --------------------------------------------------
    ab, vae = run.ab, run.vae
    ab.n_samples = run.chunk
    x0_host = torch.full((run.chunk, 1, hw, hw), K, dtype=torch.int64).pin_memory()
    um_host = torch.zeros((run.chunk, 1, hw, hw), dtype=torch.uint8).pin_memory()
    out_host = torch.empty((run.chunk, wl["in_dim"], 4 * hw, 4 * hw), dtype=torch.uint8).pin_memory()

    def e2e_pass(seed):
        for c in range(run.n_chunks):
            # H2D inside sample(): the initial (fully masked) token grid and the unmask map, from pinned host memory
            tok = ab.sample(temp=args.temp, sample_steps=run.steps_diff, seed=seed, n_global=n_global,
                            shard_base=shard_base + c * run.chunk, x_init=x0_host, unmasked_init=um_host)
            pred = vae.decode_indices(tok.reshape(run.chunk, hw, hw))
            out_host.copy_(engine.to_uint8(pred), non_blocking=True)      # D2H: the generated uint8 images
            torch.cuda.synchronize()

    e2e_pass(7)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_pass(100 + i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clk = clocks.stop()
    e2e = n_global * args.steps / e2e_s

    # ---- roofline ---------------------------------------------------------------------------------------------
    roof = None
    if rank == 0:
        per = run.layer_groups(flush)
        lif_layers = {k: v for k, v in per.items() if k != "den.conv6"}
        dom = max(lif_layers, key=lambda k: per[k]["ms"])
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            traffic = json.load(open(tf)).get(f"{args.workload}:{dom}:streams{len(run.splan.subs)}")
        ms_step = total_ms / args.steps
        step_alg = run.flops_per_image() * b / ms_step / 1e9
        step_exe = run.flops_per_image(executed=True) * b / ms_step / 1e9
        roof = {"bound": "tensor", "kernel": f"conv3x3_tc_kernel ({dom})", "achieved": per[dom]["tflops"],
                "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": round(per[dom]["tflops"] / pk["tf_burst"], 4),
                "traffic": traffic,
                "peak_source": pk["source"] + " bf16_tflops (BURST: the launch group is timed alone, L2 flushed, 11 reps)",
                "mma_kind": per[dom]["mma_kind"],
                "flops_definition": "achieved = dense un-split 2*MAC*B*T of the layer (SURVEY.md 8(d)) / group time; exact "
                                    "weights cost more than one tensor pass per algorithmic FLOP (mma_kind), which is overhead, "
                                    "not work: executed_tensor_tops gives what the tensor pipe actually ran",
                "isolated_vs_step": "the isolated groups leave their last wave partly empty; inside the step the other "
                                    "sub-batches' layers fill it, so the sum of isolated group times exceeds ms_per_step",
                "whole_step": {"tflops_algorithmic": round(step_alg, 1),
                               "tflops_readout_as_executed": round(step_exe, 1),
                               "peak": pk["tf_sustained"], "frac": round(step_exe / pk["tf_sustained"], 4),
                               "peak_source": pk["source"] + " bf16_tflops_sustained (the step runs for many ms under the power cap)",
                               "note": "tflops_readout_as_executed counts den.conv6 once on the T-summed spikes (as run) "
                                       "instead of T times (as the reference computes it)"} if world == 1 else None,
                "layers": per, "hbm_kernels": hbm_kernels(dev, flush, pk)}

    head = dict(streams=len(run.splan.subs), flop_per_image=int(run.flops_per_image()), n_chunks=run.n_chunks,
                launches_per_pass=run.launches_per_pass(), mma_kind=run.splan.dp.l4.mma_kind())

    # ---- secondary records: the other BASELINE configs and the batch sweep ---------------------------------------
    secondary = []
    if not args.no_secondary:
        del run, ab, vae
        torch.cuda.empty_cache()
        jobs = [(n, dict(WORKLOADS[n])) for n in ("cfg3", "cfg4") if n != args.workload]
        jobs += [(f"sweep_b{sb}", dict(WORKLOADS["cfg2"], b=sb,
                                       desc=f"batch sweep: cfg2 shapes (28x28, T=4, K=128), {sb} images per GPU")) for sb in SWEEP]
        for name, w in jobs:
            sb_, ng_, base_ = shard(w)
            rec = {"name": name, "workload": w["desc"], "per_gpu_batch": sb_, "global_batch": ng_,
                   "scaling": "strong" if w.get("strong") else "weak"}
            try:
                r2 = Runner(w, sb_, ng_, base_, dev, args.nsplit, args.temp)
                steps2 = 1 if sb_ >= 16384 else min(args.steps, args.secondary_steps)
                warm2 = 1 if sb_ >= 4096 else 2
                cs = ClockSampler(local_rank).start()
                ms2 = max_over_ranks(r2.timed(steps2, warm2, flush, barrier, seed0=50))
                c2 = cs.stop()
                rec.update({"value": round(ng_ * steps2 / (ms2 / 1e3), 2), "unit": "images/s", "steps": steps2,
                            "warmup": warm2, "ms_per_step": round(ms2 / steps2, 3), "clocks": c2,
                            "chunks_per_step": r2.n_chunks, "sampler_streams": len(r2.splan.subs),
                            "whole_step_tflops_readout_as_executed": round(r2.flops_per_image(executed=True) * sb_ / (ms2 / steps2) / 1e9, 1),
                            "whole_step_frac_of_sustained": round(r2.flops_per_image(executed=True) * sb_ / (ms2 / steps2) / 1e9 / pk["tf_sustained"], 4)})
                if rank == 0:
                    per2 = r2.layer_groups(flush, reps=8, skip=2)
                    lif2 = {k: v for k, v in per2.items() if k != "den.conv6"}
                    d2 = max(lif2, key=lambda k: per2[k]["ms"])
                    rec["dominant_kernel"] = {"kernel": f"conv3x3_tc_kernel ({d2})", "ms": per2[d2]["ms"],
                                              "tflops": per2[d2]["tflops"], "frac_of_burst": round(per2[d2]["tflops"] / pk["tf_burst"], 4),
                                              "mma_kind": per2[d2]["mma_kind"]}
                del r2
            except Exception as e:  # noqa: BLE001   (a failing secondary must not take the headline line with it)
                rec["error"] = f"{type(e).__name__}: {e}"
            torch.cuda.empty_cache()
            secondary.append(rec)

    # ---- SyncBN over NCCL: the one exchange step of the optional training path, checked where >= 2 GPUs are present ---
    syncbn = None
    if world > 1:
        from spiking_diffusion_b200 import selfcheck
        try:
            ok, errs = selfcheck.syncbn_over_nccl()
            syncbn = {"ok": bool(ok), "ranks": world, "max_rel_err": max(errs.values()), "tolerance": 2e-5,
                      "what": "conv -> train-mode SyncBN on a sharded batch vs one GPU on the whole batch (outputs, running "
                              "stats, gradients) + one DDP step of the spiking denoiser"}
        except Exception as e:  # noqa: BLE001
            syncbn = {"ok": False, "error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # the reference as shipped on THIS GPU (eager PyTorch, T=16, 16 images per call): the like-for-like competitor of ref16
    shipped_gpu = None
    if world == 1 and not args.no_secondary:
        try:
            shipped_gpu = reference_as_shipped("cuda", n_calls=2)
        except Exception as e:  # noqa: BLE001
            shipped_gpu = {"unavailable": f"{type(e).__name__}: {e}"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ips, cores, times, sample = cpu_port_images_per_s(wl, args.cpu_batch, steps=1, warm_steps=2)
        cpu = {"value": round(ips, 4), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
    line = {
        "metric": "generated images/sec", "value": round(value, 2), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 3),
        "higher_is_better": True, "scaling": "strong" if wl.get("strong") else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": base_config(args, wl, world),
        "arm": {"weight_terms": args.nsplit, "mma_kind": head["mma_kind"], "sampler_streams": head["streams"],
                "flop_per_image": head["flop_per_image"],
                "timing": "CUDA events per step, max over ranks; L2 flushed (256 MiB write) between timed steps"},
        "e2e": {"value": round(e2e, 2), "unit": "images/s", "h2d_bytes_per_step": int(head["n_chunks"] * (x0_host.numel() * 8 + um_host.numel())),
                "d2h_bytes_per_step": int(head["n_chunks"] * out_host.numel())},
        "gpu_launches": args.steps * head["launches_per_pass"], "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        "secondary": secondary, "syncbn_check": syncbn, "reference_as_shipped_on_gpu": shipped_gpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--temp", type=float, default=1.0)
    ap.add_argument("--nsplit", type=int, default=3, choices=[1, 2, 3],
                    help="weight representation of the tcgen05 layers: 3 int8 digits (default), 2 fp16 terms, 1 fp16 term")
    ap.add_argument("--cpu-batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="headline only (quick experiments)")
    ap.add_argument("--secondary-steps", type=int, default=3)
    ap.add_argument("--quick", action="store_true", help="reference arm: skip the full-batch and as-shipped runs")
    ap.add_argument("--batch", type=int, default=0, help="experiment: override the workload's per-GPU batch")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    if args.batch > 0:
        wl["desc"] = wl["desc"] + f" (batch overridden: {args.batch})"
        wl["b"] = args.batch
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torch.distributed.run with {args.gpus} ranks", file=sys.stderr)
        sys.exit(2)
    run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
