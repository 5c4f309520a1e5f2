"""Turn ncu artefacts under gpurun_out/ into small text summaries under profiles/ (read here, on the CPU box).
Usage: python tools/ncu_summary.py <round-tag>"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
]


def raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    rows = [x for x in rows if len(x) > 10]
    return rows[0], rows[1], rows[2:]


def summarize_rep(rep, out, title):
    hdr, units, rows = raw(rep)
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: ncu --set full --clock-control none (cold cache, serialised); file {os.path.basename(rep)}\n")
        for r in rows:
            f.write("\n== " + r[hdr.index("Kernel Name")] + "  grid " + r[hdr.index("Grid Size")] + " block " + r[hdr.index("Block Size")] + "\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:92s} {r[i]:>16s} {units[i]}\n")
            if "dram__bytes_read.sum" in hdr:
                def val(k):
                    i = hdr.index(k)
                    v = float(r[i].replace(",", ""))
                    u = units[i].lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                f.write(f"{'traffic = dram read + write (bytes)':92s} {val('dram__bytes_read.sum') + val('dram__bytes_write.sum'):16.0f}\n")
    print("wrote", out)


def summarize_launches(csvf, out, title):
    rows = [r for r in csv.reader(open(csvf)) if len(r) > 8]
    hdr = rows[0]
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    seq = []
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        name = r[ki].split("(")[0].replace("void ", "").replace("sd::", "")
        agg.setdefault(name, []).append(v)
        seq.append((name, r[gi], v))
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n# per-launch device times from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold cache, "
                "serialised: compare SHARES, not absolutes)\n\n")
        f.write(f"{'kernel':44s} {'launches':>8s} {'total us':>11s} {'mean us':>9s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:44s} {len(v):8d} {sum(v):11.1f} {sum(v) / len(v):9.2f} {100 * sum(v) / tot:6.1f}%\n")
        f.write(f"{'TOTAL':44s} {sum(len(v) for v in agg.values()):8d} {tot:11.1f}\n\n# one diffusion step, in launch order:\n")
        start = next((i for i, s in enumerate(seq) if s[0].startswith("conv_real_const_lif")), 0)
        f.write("# (sub-batch 0 of 5; the other sub-batches follow in the same order)\n")
        for name, grid, v in seq[start:start + 7]:
            f.write(f"  {name:44s} grid {grid:14s} {v:9.2f} us\n")
        first_dec = next((i for i, s in enumerate(seq) if s[0].startswith("vq_gather")), None)
        f.write("\n# decode after the last diffusion step, in launch order:\n")
        for name, grid, v in ([] if first_dec is None else seq[first_dec:first_dec + 8]):
            f.write(f"  {name:44s} grid {grid:14s} {v:9.2f} us\n")
    print("wrote", out)


def write_traffic(rep, n_subs=5, workload="cfg2", tag="r01"):
    """profiles/traffic.json: DRAM bytes per LAUNCH GROUP (the layer's launches of all sub-batches of one diffusion
    step).  The capture holds 5 tcgen05 layers x n_subs sub-batches in launch order (sub-batch major)."""
    import json
    hdr, units, rows = raw(rep)
    if len(rows) < 5 * n_subs:
        print("traffic: capture too short", len(rows))
        return

    def val(r, k):
        i = hdr.index(k)
        v = float(r[i].replace(",", ""))
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[i].lower(), 1)

    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per LAUNCH GROUP (the layer's launches of all "
                       f"{n_subs} sub-batches of one diffusion step), summed from profiles/{tag}_conv3x3_tc.txt "
                       "(ncu --set full, cold cache, serialised)"}
    for li, name in enumerate(("conv2", "conv3", "conv4", "conv5", "conv6")):
        tot = sum(val(rows[s * 5 + li], "dram__bytes_read.sum") + val(rows[s * 5 + li], "dram__bytes_write.sum")
                  for s in range(n_subs))
        out[f"{workload}:den.{name}:streams{n_subs}"] = int(tot)
    json.dump(out, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    print("wrote traffic.json", out)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    if os.path.exists(os.path.join(OUT, "p_launches.csv")):
        summarize_launches(os.path.join(OUT, "p_launches.csv"), os.path.join(PROF, f"{tag}_launch_list.txt"),
                           "bench.py --steps 1 --warmup 1 (cfg2: b=256, T=4, K=128, 49 steps + decode), 5 sub-batches (4 x 52 + 48 images), SD_SAMPLER_GRAPH=0")
    for rep, name, title in (("p_conv_tc.ncu-rep", "conv3x3_tc", "fused conv+BN+LIF tcgen05 kernel: den.conv2..conv6 of one diffusion step, sub-batch by sub-batch (cfg2: 4 x 52 + 48 images)"),
                             ("p_sample.ncu-rep", "sample_step", "fused sampling-step kernel (cfg2: 12544 tokens, K=128)"),
                             ("p_conv1.ncu-rep", "conv_real_const_lif", "den.conv1: real-input conv + BN + LIF (cfg2)")):
        if os.path.exists(os.path.join(OUT, rep)):
            if name == "conv3x3_tc":
                write_traffic(os.path.join(OUT, rep), tag=tag)
            summarize_rep(os.path.join(OUT, rep), os.path.join(PROF, f"{tag}_{name}.txt"), title)
