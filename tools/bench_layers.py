"""Per-layer timing of the denoiser's tcgen05 layers under different kernel configurations (GPU box).
Usage: python tools/bench_layers.py [workload] ; configurations are SD_TC_* environment overrides."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from spiking_diffusion_b200 import _lib, engine  # noqa: E402

CONFIGS = [
    {},
    {"SD_TC_TACC": "2"},
    {"SD_TC_TACC": "2", "SD_TC_KBLK": "32"},
]


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
    dev = torch.device("cuda", 0)
    vae, den, ab, _, _ = bench.build_models(wl, dev)
    b, T, hw = wl["b"], wl["T"], wl["hw"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    base = None
    for cfg in CONFIGS:
        for k in list(os.environ):
            if k.startswith("SD_TC_"):
                del os.environ[k]
        os.environ.update(cfg)
        _lib.check(_lib.lib().sd_debug_tc_reload_knobs())   # the library reads the SD_TC_* knobs once per process
        try:
            dp = engine.DenoiserPlan(den, T, b, hw, hw, nsplit=2)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"cfg": cfg, "error": str(e)}))
            continue
        x_t = torch.full((b * hw * hw,), wl["K"], dtype=torch.int64, device=dev)
        x_t[::3] = 5
        dp.run_tokens(x_t, 7)
        torch.cuda.synchronize()
        outs = {n: getattr(dp, n).clone() for n in ("x2", "x3", "x4", "x5")}
        if base is None:
            base = (outs, dp.logits.clone())
            flips = 0.0
        else:
            flips = max(float((outs[n] != base[0][n]).float().mean()) for n in outs)
        layers = [("conv2", dp.l2, dp.x1, dp.x2, None, None), ("conv3", dp.l3, dp.x2, dp.x3, None, None),
                  ("conv4", dp.l4, dp.x3, dp.x4, None, None), ("conv5", dp.l5, dp.x4, dp.x5, dp.x5s, None),
                  ("conv6", dp.l6, dp.x5s, dp.logits, None, dp.x1s)]
        res = {}
        for n, l, xi, xo, xs, x2 in layers:
            ts = []
            for rep in range(12):
                flush.zero_()
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); l.run(xi, xo, x2=x2, out_sum=xs); c.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(c))
            ts = sorted(ts[2:])
            ms = sum(ts) / len(ts)
            res[n] = {"ms": round(ms, 4), "tflops": round(l.flops() / ms / 1e9, 1)}
        tot = sum(v["ms"] for v in res.values())
        print(json.dumps({"cfg": cfg, "total_ms": round(tot, 4), "max_flip_vs_first": flips,
                          "logit_err_vs_first": float((dp.logits - base[1]).abs().max()), "layers": res}), flush=True)


if __name__ == "__main__":
    main()
