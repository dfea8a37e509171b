#!/usr/bin/env python
"""SASS opcode histogram of the chain kernels in the built library (cuobjdump -sass; no GPU needed):

    python tools/sass_hist.py [kernel-name-regex ...] > profiles/rNN_sass_chain_kernels.txt

Default kernels: the three tcgen05 MLP kernels (forward chain, dgrad chain, wgrad GEMM).  Prints, per kernel, the
instruction count, the counts of the Blackwell tensor / TMA / TMEM mnemonics (UTCHMMA = tcgen05.mma, UTCBAR =
tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops) and the 25 most frequent
opcodes."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("LUSH_LIB_PATH") or os.path.join(ROOT, "lush_nerf_b200", "liblushnerf_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "FFMA", "F2FP",
       "MUFU", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR", "FMNMX", "SHF", "PRMT")


def main():
    pats = [re.compile(p) for p in (sys.argv[1:] or ["mlp_tc_fwd_kernel", "mlp_tc_dgrad_kernel", "mlp_tc_wgrad_kernel"])]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = name if any(p.search(name) for p in pats) else None
            if cur:
                hist[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Za-z0-9_]+)*)", line)
        if m:
            hist[cur][m.group(1)] += 1
    print(f"# {os.path.relpath(LIB, ROOT)} -- cuobjdump -sass, opcode histogram per kernel (sm_100a)")
    for name, h in hist.items():
        total = sum(h.values())
        print(f"\n## {name.split('(')[0]}   [{total} instructions]")
        print("   " + "  ".join(f"{k}={h[k]}" for k in KEY if h.get(k)))
        print("   top: " + ", ".join(f"{k} {v}" for k, v in h.most_common(25)))


if __name__ == "__main__":
    main()
