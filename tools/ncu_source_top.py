#!/usr/bin/env python3
"""Source-level (SASS) summary of an `ncu --set full --import-source on` report, per kernel:
  * the instructions with the most stall samples and their dominant stall reasons,
  * the stall-reason totals of the kernel,
  * every global access whose warp instruction touches more 32-byte sectors than it needs
    ("L2 Theoretical Sectors Global Excessive": partial-sector or strided accesses -- the forward chain's 4-byte ReLU-mask
    store at a 32-byte stride was found this way: a third of the kernel's L2 write transactions for 1/16 of its bytes).

    python tools/ncu_source_top.py gpurun_out/r2_ncu_allstep.ncu-rep [kernel-regex] [--top 25]
"""
import argparse
import collections
import csv
import io
import re
import subprocess


def blocks_of(report, regex):
    cmd = ["ncu", "-i", report, "--page", "source", "--csv"]
    if regex:
        cmd += ["--kernel-name", "regex:" + regex]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    cur = None
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            if cur:
                yield cur
            cur = {"name": r[1], "hdr": None, "rows": []}
        elif cur is not None:
            if cur["hdr"] is None:
                cur["hdr"] = r
            else:
                cur["rows"].append(r)
    if cur:
        yield cur


def num(x):
    try:
        return int(x.replace(",", ""))
    except ValueError:
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("regex", nargs="?", default="")
    ap.add_argument("--top", type=int, default=20)
    a = ap.parse_args()
    seen = collections.Counter()
    for b in blocks_of(a.report, a.regex):
        ix = {h: i for i, h in enumerate(b["hdr"])}
        if "# Samples" not in ix:
            continue
        short = re.sub(r"\(.*", "", b["name"])[-60:]
        seen[short] += 1
        stalls = [h for h in b["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(num(r[ix["# Samples"]]) for r in b["rows"])
        by = collections.Counter()
        for r in b["rows"]:
            for h in stalls:
                by[h[6:]] += num(r[ix[h]])
        print(f"== {short}  (launch {seen[short]} of this kernel in the report)   {len(b['rows'])} instructions, {tot} samples")
        print("   stalls: " + ", ".join(f"{k} {100 * v / max(1, sum(by.values())):.1f}%" for k, v in by.most_common(8)))
        for r in sorted(b["rows"], key=lambda r: -num(r[ix["# Samples"]]))[:a.top]:
            n = num(r[ix["# Samples"]])
            if n * 200 < tot:
                break
            st = sorted(((num(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
            print(f"   {n:7d} {100 * n / tot:5.1f}%  {r[ix['Source']][:64]:64s} " + " ".join(f"{h}:{v}" for v, h in st if v))
        if "L2 Theoretical Sectors Global Excessive" in ix:
            T = sum(num(r[ix["L2 Theoretical Sectors Global"]]) for r in b["rows"])
            for r in b["rows"]:
                ex, t = num(r[ix["L2 Theoretical Sectors Global Excessive"]]), num(r[ix["L2 Theoretical Sectors Global"]])
                if ex * 50 > T:     # more than 2 % of the kernel's sector transactions are waste of this instruction
                    print(f"   sectors: {r[ix['Source']][:60]:60s} {ex} excessive of {t} ({100 * t / T:.1f}% of the kernel's)")


if __name__ == "__main__":
    main()
