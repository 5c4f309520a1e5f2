"""Cycle-stamp trace of the tcgen05 conv kernel's MMA and epilogue warps per layer (GPU box).
Usage: python tools/trace_tc.py [workload] [sub_batch] [nsplit]   -> where a tile's time goes: operand stalls, MMA, epilogue, hand-off."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spiking_diffusion_b200 import _lib  # noqa: E402
# the cycle stamps exist only in the -DSD_TRACE build of the library (build/libsd_b200_trace.so); build it here (on the
# CPU box, before gpurun) or on the GPU box, and make the package load it instead of the shipped library
if not os.path.exists(_lib.TRACE_LIB_PATH) or os.environ.get("SD_TRACE_REBUILD"):
    _lib.build(trace=True)
os.environ["SD_B200_LIB"] = _lib.TRACE_LIB_PATH
import bench  # noqa: E402
from spiking_diffusion_b200 import engine  # noqa: E402


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
    b = int(sys.argv[2]) if len(sys.argv) > 2 else wl["b"]
    dev = torch.device("cuda", 0)
    vae, den, ab, _, _ = bench.build_models(wl, dev)
    T, hw = wl["T"], wl["hw"]
    nsplit = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    dp = engine.DenoiserPlan(den, T, b, hw, hw, nsplit=nsplit)
    print(f"# workload {sys.argv[1] if len(sys.argv) > 1 else 'cfg2'}, b={b}, nsplit={nsplit} ({dp.l4.mma_kind()})")
    x_t = torch.full((b * hw * hw,), wl["K"], dtype=torch.int64, device=dev)
    x_t[::3] = 5
    for _ in range(3):
        dp.run_tokens(x_t, 7)
    torch.cuda.synchronize()
    lib = _lib.lib()
    trace = torch.zeros(296 * 128, dtype=torch.int64, device=dev)
    layers = [("conv2", dp.l2, dp.x1, dp.x2, None, None), ("conv3", dp.l3, dp.x2, dp.x3, None, None),
              ("conv4", dp.l4, dp.x3, dp.x4, None, None), ("conv5", dp.l5, dp.x4, dp.x5, dp.x5s, None),
              ("conv6", dp.l6, dp.x5s, dp.logits, None, dp.x1s)]
    for n, l, xi, xo, xs, x2 in layers:
        trace.zero_()
        _lib.check(lib.sd_debug_tc_trace(trace.data_ptr()))
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); l.run(xi, xo, x2=x2, out_sum=xs); c.record()
        torch.cuda.synchronize()
        _lib.check(lib.sd_debug_tc_trace(None))
        tr = trace.cpu().numpy().reshape(296, 128)
        used = tr[:, 0] != 0
        g = int(used.sum())
        tr = tr[used]
        print(f"== {n}: {a.elapsed_time(c) * 1e3:.1f} us (traced), grid {g}")
        life = tr[:, 2] - tr[:, 0]
        print(f"   CTA lifetime cycles: min {life.min()} median {int(np.median(life))} max {life.max()};  set-up {int(np.median(tr[:, 1] - tr[:, 0]))}")
        # leaders own the MMA stamps (all CTAs when not paired); every CTA has epilogue stamps
        for it in range(7):
            base = 3 + 8 * it
            m = tr[:, base + 0] != 0
            e = tr[:, base + 4] != 0
            if not e.any():
                break
            msg = f"   pass {it}: "
            if m.any():
                t = tr[m]
                msg += (f"mma CTAs {int(m.sum())}: acc->first operands {int(np.median(t[:, base + 1] - t[:, base + 0]))}, "
                        f"issue span {int(np.median(t[:, base + 2] - t[:, base + 0]))}, operand stall {int(np.median(t[:, base + 3]))} "
                        f"(max {int(t[:, base + 3].max())}); ")
                both = m & e
                t = tr[both]
                msg += f"acc ready after {int(np.median(t[:, base + 4] - t[:, base + 0]))} (max {int((t[:, base + 4] - t[:, base + 0]).max())}); "
                if it > 0:
                    pb = base - 8
                    msg += f"hand-off epi release -> mma start {int(np.median(t[:, base + 0] - t[:, pb + 5]))}; "
            t = tr[e]
            msg += f"epilogue {int(np.median(t[:, base + 5] - t[:, base + 4]))} (max {int((t[:, base + 5] - t[:, base + 4]).max())}) on {int(e.sum())} CTAs"
            g = t[t[:, base + 7] != 0]
            if len(g):   # LIF layers: first 16-column group of warp 0
                msg += (f"; group 0: state loaded after {int(np.median(g[:, base + 6] - g[:, base + 4]))}, "
                        f"LIF over the pass's timesteps {int(np.median(g[:, base + 7] - g[:, base + 6]))}")
            print(msg)
        # warp 0's 16-column groups of pass 1: start -> state loaded -> LIF done -> end (and the gap to the next group)
        gp = tr[tr[:, 64] != 0][:, 64:64 + 16].reshape(-1, 4, 4)
        if len(gp):
            gp = gp[(gp[:, :, 3] != 0).all(axis=1)]
        if len(gp):
            med = lambda a: np.median(a, axis=0).astype(int).tolist()   # noqa: E731
            print(f"   pass 1, warp 0, per 16-column group: load state {med(gp[:, :, 1] - gp[:, :, 0])}, LIF {med(gp[:, :, 2] - gp[:, :, 1])}, "
                  f"store {med(gp[:, :, 3] - gp[:, :, 2])}, between groups {med(gp[:, 1:, 0] - gp[:, :-1, 3])}; "
                  f"acc ready -> first group {int(np.median(gp[:, 0, 0] - tr[tr[:, 64] != 0][:len(gp), 3 + 8 + 4]))}, "
                  f"last group -> release {int(np.median(tr[tr[:, 64] != 0][:len(gp), 3 + 8 + 5] - gp[:, 3, 3]))}")
        # tail: last epilogue release -> exit
        last = np.zeros(len(tr), dtype=np.int64)
        for it in range(7):
            v = tr[:, 3 + 8 * it + 5]
            last = np.where(v != 0, v, last)
        print(f"   exit - last epilogue release: median {int(np.median(tr[:, 2] - last))} max {int((tr[:, 2] - last).max())}")


if __name__ == "__main__":
    main()
