"""cfg1 of SURVEY.md 8(d): SNN_VQVAE.forward in eval mode (encoder -> quantiser -> spike generator -> decoder), B=64, T=4,
timed on the GPU through the public module and through the fused plan, per stage, with the oracle port of the
reference timed on the host cores beside it.  Usage: python tools/bench_vqvae.py [B] [T]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import snn_oracle as O  # noqa: E402  (CPU leg only)
from spiking_diffusion_b200 import engine  # noqa: E402
from spiking_diffusion_b200.activation_based import functional  # noqa: E402


def gpu_time(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    wl = dict(bench.WORKLOADS["cfg2"], b=B, T=T)
    dev = torch.device("cuda", 0)
    vae, den, ab, vsd, dsd = bench.build_models(wl, dev)
    torch.manual_seed(0)
    img = torch.rand(B, 1, 28, 28) - 0.5
    x_seq = img.unsqueeze(0).repeat(T, 1, 1, 1, 1)
    xg, ig = x_seq.to(dev), img.to(dev)
    plan = engine.VQVAEPlan(vae, T, B, 28, 28)

    def module_forward():
        functional.reset_net(vae)
        with torch.no_grad():
            return vae(xg, ig)

    res = {"B": B, "T": T, "flop_per_image": plan.flops() // B}
    res["module_forward_ms"] = round(gpu_time(module_forward), 4)
    res["plan_forward_ms"] = round(gpu_time(lambda: plan.forward(xg)), 4)
    res["plan_forward_const_input_ms"] = round(gpu_time(lambda: plan.forward(ig, const_over_T=True)), 4)
    z = plan.encode(xg)
    idx = plan.quantize_indices(z).clone()
    e = plan.generate(idx)
    res["stages_ms"] = {
        "encode": round(gpu_time(lambda: plan.encode(xg)), 4),
        "encode_const_input": round(gpu_time(lambda: plan.encode(ig, const_over_T=True)), 4),
        "vq_feature+lookup": round(gpu_time(lambda: plan.quantize_indices(z)), 4),
        "gather+generator": round(gpu_time(lambda: plan.generate(idx)), 4),
        "decode": round(gpu_time(lambda: plan.decode(e)), 4),
    }
    if plan.tc_decoder:
        from spiking_diffusion_b200._lib import check, lib, ptr, stream_ptr
        L, d1, d2 = lib(), plan.d1.desc, plan.d2.desc
        res["decode_stages_ms"] = {
            "upsample0": round(gpu_time(lambda: check(L.sd_stf_upsample2x(ptr(e), ptr(plan.up0), T, B, d1.C_in, plan.h, plan.w, stream_ptr()))), 4),
            "convT1 (tc)": round(gpu_time(lambda: plan.d1.run(plan.up0, plan.sd1)), 4),
            "upsample1": round(gpu_time(lambda: check(L.sd_stf_upsample2x(ptr(plan.sd1), ptr(plan.up1), T, B, d2.C_in, 2 * plan.h, 2 * plan.w, stream_ptr()))), 4),
            "convT2 (tc)": round(gpu_time(lambda: plan.d2.run(plan.up1, plan.sd2)), 4),
            "convT3+memout+tanh": round(gpu_time(lambda: plan.d3.run(plan.sd2, plan.recon)), 4),
        }
    else:
        res["decode_stages_ms"] = {
            "convT1": round(gpu_time(lambda: plan.d1.run(e, plan.sd1)), 4),
            "convT2": round(gpu_time(lambda: plan.d2.run(plan.sd1, plan.sd2)), 4),
            "convT3+memout+tanh": round(gpu_time(lambda: plan.d3.run(plan.sd2, plan.recon)), 4),
        }
    res["encode_stages_ms"] = {
        "conv1 (const input)": round(gpu_time(lambda: plan.e1c.run(ig, plan.s1)), 4),
        "conv2 s2": round(gpu_time(lambda: plan.e2.run(plan.s1, plan.s2)), 4),
        "conv3 1x1": round(gpu_time(lambda: plan.e3.run(plan.s2, plan.s3)), 4),
    }
    res["images_per_s_gpu"] = round(B / res["plan_forward_ms"] * 1e3, 1)
    # CPU leg: the oracle port of the reference on all host threads
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    with torch.inference_mode():
        for _ in range(2):
            O.vqvae_forward_eval(x_seq, vsd)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            O.vqvae_forward_eval(x_seq, vsd)
            ts.append(time.perf_counter() - t0)
    res["cpu_oracle_ms"] = round(1e3 * min(ts), 2)
    res["cpu_cores"] = cores
    res["images_per_s_cpu"] = round(B / min(ts), 1)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
