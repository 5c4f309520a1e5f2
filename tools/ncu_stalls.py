"""Warp-state samples of the lean int8 LIF epilogue's group loop, from a full ncu capture with source counters
(gpurun_out/p_conv_tc.ncu-rep, tools/gpu_ncu_tc25.sh).  Usage: python tools/ncu_stalls.py [report] [launch index]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/p_conv_tc.ncu-rep"
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 3          # conv5 of the first sub-batch (launch order conv2..conv6)
    r = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::conv3x3_tc_kernel:{idx}"],
                       capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    h = rows[1]
    col = {n: i for i, n in enumerate(h)}
    data = [x for x in rows[2:] if len(x) >= len(h) - 2 and x[0].startswith("0x")]
    half = len(data) // 2
    if [x[col["Source"]] for x in data[:half]] == [x[col["Source"]] for x in data[half:]]:
        data = data[:half]                                       # the page lists the function twice
    S, SRC, IE = col["# Samples"], col["Source"], col["Instructions Executed"]
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    ld = [i for i, x in enumerate(data) if "LDTM" in x[SRC] and int(x[IE]) > 0]
    lo, hi = ld[0] - 80, ld[1] + 260                             # the group loop around the executed TMEM loads
    total = sum(int(x[S]) for x in data)
    region = data[lo:hi]
    n_reg = sum(int(x[S]) for x in region)
    agg = collections.Counter()
    for x in region:
        for c in stalls:
            agg[c] += int(x[col[c]])
    print(f"# {name}, launch {idx} of {rep}")
    print(f"# {len(data)} SASS instructions, {total} warp-state samples; epilogue group loop = instructions {lo}..{hi}: {n_reg} samples")
    for c, v in agg.most_common():
        if v:
            print(f"{c:28s} {v:6d}  {100.0 * v / max(n_reg, 1):5.1f} %")
    ops = collections.Counter()
    for x in region:
        t = x[SRC].strip().split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += int(x[IE])
    tot_i = sum(ops.values())
    print(f"# warp instructions executed in the loop: {tot_i}; by opcode:")
    print("  " + ", ".join(f"{o} {100.0 * v / tot_i:.1f}%" for o, v in ops.most_common(14)))


if __name__ == "__main__":
    main()
