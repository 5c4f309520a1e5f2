#!/usr/bin/env python
"""Benchmark of the Spiking-Diffusion hot path: generated images/sec (BASELINE.json:metric).

A *step* is one pass of the whole path over one batch: AbsorbingDiffusion.sample (h*w reverse-diffusion steps, each
a full T-timestep spiking denoiser forward + categorical draw + unmask update) followed by the decode of
R/main.py:388-401 (quantize -> spike generator -> spiking decoder -> tanh(memout) -> uint8).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Our arm: `value` = images/s with everything resident in HBM, device-timed (CUDA events, max over ranks);
`e2e` = the same through the public API with host buffers (pinned H2D of the initial token grid, D2H of the uint8
images inside the timed region).  `--impl reference` times the reference's algorithm on the host CPU (the oracle
port: the Python reference itself cannot travel to the GPU box) on a bounded sample of the same workload.
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
    "cfg4": dict(desc="KMNIST/Letters-shape 28x28 sampling, b=512 per GPU, T=8, K=512, 49 steps + decode", b=512, T=8,
                 K=512, hw=7, in_dim=1),
    # the reference as shipped: T=16 (hard-coded), 16 samples per sample() call x 2 (R/main.py:383-387)
    "ref16": dict(desc="reference as shipped: 28x28 sampling, b=32, T=16, K=128, 49 steps + decode", b=32, T=16, K=128,
                  hw=7, in_dim=1),
}


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


def cpu_reference_images_per_s(wl, b_cpu, steps=1, warm_steps=3):
    """The reference's algorithm (oracle port, torch CPU fp32, all host threads) on a bounded sample of the workload:
    sample() for `b_cpu` images + decode.  Returns (images/s, cores, description)."""
    import numpy as np
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
        x = torch.full((b_cpu, 1, hw, hw), float(K))
        for _ in range(warm_steps):
            O.denoiser_forward(x, torch.full((b_cpu,), 1), dsd, T)
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            tok = O.sample(dsd, T, b_cpu, (hw, hw), K, K, 1.0, hw * hw, uni, expo)
            O.to_uint8(O.decode_indices(tok.reshape(b_cpu, hw, hw), vsd, T))
            times.append(time.perf_counter() - t0)
    return b_cpu / min(times), cores, times, (f"oracle port of the reference (torch {torch.__version__} CPU fp32), sample()+decode of "
                                              f"{b_cpu} images, {hw * hw} diffusion steps, T={T}, K={K}, best of {steps}")


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    b_cpu = args.cpu_batch
    t0 = time.perf_counter()
    for _ in range(0):
        pass
    ips, cores, times, sample = cpu_reference_images_per_s(wl, b_cpu, steps=max(1, args.steps), warm_steps=max(1, args.warmup))
    line = {
        "impl": "reference", "metric": "generated images/sec", "value": round(ips, 4), "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * min(times), 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']} (CPU arm: bounded sample of {b_cpu} images per step)"},
        "cpu_baseline": {"value": round(ips, 4), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(ips, 4), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 1),
    }
    print(json.dumps(line), flush=True)


def run_ours(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import spiking_diffusion_b200 as sd
    from spiking_diffusion_b200 import engine, _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    vae, den, ab, vsd, dsd = build_models(wl, dev)
    den.nsplit = args.nsplit
    b, T, K, hw = wl["b"], wl["T"], wl["K"], wl["hw"]
    steps_diff = hw * hw
    n_global, shard_base = b * world, b * rank          # weak scaling: per-GPU work fixed, shards of one global stream
    splan = ab.plan(b, n_global, shard_base)
    vplan = vae.plan(T, b, 4 * hw, 4 * hw)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    img8 = torch.empty((b, wl["in_dim"], 4 * hw, 4 * hw), dtype=torch.uint8, device=dev)

    def one_pass(seed):
        tok = splan.sample(args.temp, steps_diff, seed, 0)
        pred = vplan.decode_indices(tok)
        _lib.check(_lib.lib().sd_to_uint8(pred.data_ptr(), img8.data_ptr(), pred.numel(), _lib.stream_ptr()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        one_pass(1000 + i)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        one_pass(i)
        ev[i][1].record()
    barrier()
    ms = [a.elapsed_time(b_) for a, b_ in ev]
    total_ms = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    value = n_global * args.steps / (total_ms / 1e3)

    # ---- end to end through the public API with host buffers --------------------------------------------------
    x0_host = torch.full((b, 1, hw, hw), K, dtype=torch.int64).pin_memory()
    um_host = torch.zeros((b, 1, hw, hw), dtype=torch.uint8).pin_memory()
    out_host = torch.empty((b, wl["in_dim"], 4 * hw, 4 * hw), dtype=torch.uint8).pin_memory()

    def e2e_pass(seed):
        # H2D inside sample(): the initial (fully masked) token grid and the unmask map, from pinned host memory
        tok = ab.sample(temp=args.temp, sample_steps=steps_diff, seed=seed, n_global=n_global, shard_base=shard_base,
                        x_init=x0_host, unmasked_init=um_host)
        pred = vae.decode_indices(tok.reshape(b, hw, hw))
        out_host.copy_(engine.to_uint8(pred), non_blocking=True)      # D2H: the generated uint8 images
        torch.cuda.synchronize()

    e2e_pass(7)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_pass(100 + i)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    clk = clocks.stop()
    e2e = n_global * args.steps / float(e2e_s)

    # ---- roofline of the dominant kernel --------------------------------------------------------------------------
    # The timed region launches every layer once per sub-batch, the sub-batches on separate streams.  Each layer is
    # therefore measured as it runs there: one "launch group" = the layer's launches of all sub-batches, forked onto
    # their streams from one start event and joined into one stop event.  FLOPs are those of the whole group.
    roof = None
    if rank == 0:
        names = [("den.conv2", "l2", "x1", "x2", None), ("den.conv3", "l3", "x2", "x3", None),
                 ("den.conv4", "l4", "x3", "x4", None), ("den.conv5", "l5", "x4", "x5", "x5s")]
        cur = torch.cuda.current_stream()
        acc = {n: [] for n, *_ in names}
        for rep in range(14):
            flush.zero_()
            for n, ln, xi, xo, xs in names:
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(cur)
                for (dp, _, _), st in zip(splan.subs, splan.streams):
                    run = lambda: getattr(dp, ln).run(getattr(dp, xi), getattr(dp, xo), out_sum=getattr(dp, xs) if xs else None)
                    if st is None:
                        run()
                    else:
                        st.wait_event(a)
                        with torch.cuda.stream(st):
                            run()
                        done = torch.cuda.Event()
                        done.record(st)
                        cur.wait_event(done)
                c.record(cur)
                acc[n].append((a, c))
        torch.cuda.synchronize()
        pk = peaks()
        per = {}
        for n, ln, *_ in names:
            t = sorted(a.elapsed_time(c) for a, c in acc[n][3:])
            mean_ms = sum(t) / len(t)
            fl = sum(getattr(dp, ln).flops() for dp, _, _ in splan.subs)
            per[n] = dict(ms=round(mean_ms, 4), tflops=round(fl / mean_ms / 1e9, 1), impl=getattr(splan.dp, ln).impl,
                          launches_per_group=len(splan.subs))
        dom = max(per, key=lambda k: per[k]["ms"])
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            traffic = json.load(open(tf)).get(f"{args.workload}:{dom}:streams{len(splan.subs)}")
        step_flops = splan.flops_per_image(steps_diff) * b + vplan.flops()
        roof = {"bound": "tensor", "kernel": f"conv3x3_tc_kernel ({dom})", "achieved": per[dom]["tflops"],
                "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": round(per[dom]["tflops"] / pk["tf_sustained"], 4),
                "traffic": traffic, "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                "nsplit": args.nsplit,
                "flops_definition": "dense un-split 2*MAC*B*T (SURVEY.md 8(d)); the two exact fp16 weight terms are overhead, "
                                    "so the ceiling of this fraction is 0.5",
                "whole_step_tflops": round(step_flops * n_global / b / (total_ms / args.steps) / 1e9, 1) if world == 1 else None,
                "layers": per}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ips, cores, times, sample = cpu_reference_images_per_s(wl, args.cpu_batch, steps=1, warm_steps=2)
        cpu = {"value": round(ips, 4), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
    launches = args.steps * (steps_diff * splan.kernel_launches_per_step + 5 + 1)
    line = {
        "metric": "generated images/sec", "value": round(value, 2), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "global_batch": n_global, "parallelism": f"batch-shard x{world}, no collective on the sampling path",
                   "temp": args.temp, "weight_split_terms": args.nsplit,
                   "timing": "CUDA events per step, max over ranks; L2 flushed (256 MiB write) between timed steps",
                   "sampler_streams": len(splan.subs),
                   "flop_per_image": int(splan.flops_per_image(steps_diff) + vplan.flops() // b)},
        "e2e": {"value": round(e2e, 2), "unit": "images/s", "h2d_bytes_per_step": int(x0_host.numel() * 8 + um_host.numel()),
                "d2h_bytes_per_step": int(out_host.numel())},
        "gpu_launches": launches, "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
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
    ap.add_argument("--nsplit", type=int, default=2, choices=[1, 2])
    ap.add_argument("--cpu-batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=0, help="experiment: override the workload's per-GPU batch")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    if args.batch > 0:
        wl["desc"] = wl["desc"].replace(f"b={wl['b']}", f"b={args.batch} (overridden)")
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
