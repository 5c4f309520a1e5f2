"""SASS evidence for profiles/: which Blackwell-native instructions each kernel of libsd_b200.so contains (cuobjdump -sass,
read on the CPU box).  tcgen05.mma -> UTCHMMA (kind::f16) / UTCIMMA (kind::i8), tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP,
tcgen05.commit -> UTCBAR, tcgen05.alloc -> UTCATOMSWS ...; HMMA (legacy mma.sync) must not appear.
Usage: python tools/sass_summary.py > profiles/r02_sass.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spiking-diffusion_b200", "libsd_b200.so")
PAT = re.compile(r"\b(UTC[A-Z0-9]*MMA(?:\.2CTA)?|LDTM|STTM|UBLKCP(?:\.[A-Z.]+)?|UTMALDG|UTCBAR(?:\.[A-Z0-9.]+)?|UTCATOMSWS[A-Z.]*|SYNCS\.[A-Z.0-9]+|UCGABAR_[A-Z]+|HMMA[A-Z0-9.]*|IMMA[A-Z0-9.]*)")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, per = None, collections.OrderedDict()
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("sd::", "")
            per[cur] = collections.Counter()
            continue
        if cur:
            for tok in PAT.findall(ln):
                per[cur][tok.split(".")[0] + (".2CTA" if ".2CTA" in tok else "")] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} ({os.path.getsize(LIB)} bytes): counts of tensor-core / TMEM / bulk-copy /")
    print("# mbarrier instructions per kernel.  UTCHMMA = tcgen05.mma kind::f16, UTCIMMA = kind::i8, .2CTA = cta_group::2,")
    print("# LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, UCGABAR = cluster barrier.  No HMMA/IMMA (mma.sync).\n")
    for k, c in per.items():
        if any(t.startswith(("UTC", "LDTM", "UBLKCP", "HMMA", "IMMA")) for t in c):
            print(f"{k}\n    " + ", ".join(f"{t} {n}" for t, n in sorted(c.items())))
    others = [k for k, c in per.items() if not any(t.startswith(("UTC", "LDTM", "UBLKCP", "HMMA", "IMMA")) for t in c)]
    print(f"\n# {len(others)} other kernels (CUDA-core / HBM-bound: LIF, VQ, sampling step, SIMT convs, training, metrics) contain none of these.")
    legacy = [k for k, c in per.items() if any(t.startswith(("HMMA", "IMMA")) for t in c)]
    print(f"# kernels with legacy mma.sync (HMMA/IMMA): {legacy if legacy else 'none'}")


if __name__ == "__main__":
    main()
