#!/usr/bin/env python3
"""Static check of the store hazard in the chain kernels' epilogues.

A global store holds its data registers until the load/store unit has read them (a write-after-read hazard on the long
scoreboard: ncu shows it as stall_long_sb on the instruction that next WRITES one of those registers; profiles/README.md,
round 2).  For every 128-bit global store of a kernel this prints the distance, in instructions, to the first later
instruction that overwrites one of its four data registers -- short distances are stalls.

    python tools/sass_war.py lush_nerf_b200/liblushnerf_b200.so mlp_tc_fwd_kernelILb1
"""
import re
import subprocess
import sys


def kernel_sass(lib, name):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, keep = [], None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if keep:
                return cur
            keep = name in m.group(1)
            cur = []
            continue
        if keep:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                cur.append(m.group(2).strip())
    return cur if keep else []


def dest_regs(ins):
    """registers written by a SASS instruction (first operand; wide ops write consecutive registers)"""
    body = re.sub(r"^@!?U?P\d+\s+", "", ins)
    op, _, rest = body.partition(" ")
    if op.startswith(("ST", "BRA", "BAR", "SYNCS", "EXIT", "NOP", "UTC", "RED", "ATOM", "WARPSYNC", "BSYNC", "BSSY",
                      "YIELD", "MEMBAR", "FENCE", "ISETP", "FSETP", "R2UR", "PLOP3", "CS2R" * 0 or "ZZZ")):
        return set()
    m = re.match(r"\s*R(\d+)", rest)
    if not m:
        return set()
    r = int(m.group(1))
    n = 1
    if ".128" in op or ".x4" in op:
        n = 4
    elif ".64" in op or "WIDE" in op or "DADD" in op:
        n = 2
    m2 = re.search(r"\.x(\d+)", op)
    if op.startswith("LDTM") and m2:
        n = int(m2.group(1))
    return set(range(r, r + n))


def main():
    lib, name = sys.argv[1], sys.argv[2]
    sass = kernel_sass(lib, name)
    print(f"{name}: {len(sass)} instructions")
    for i, ins in enumerate(sass):
        m = re.search(r"STG\.E(?:\.EF)?\.128 .*\], R(\d+)", ins)
        if not m:
            continue
        data = set(range(int(m.group(1)), int(m.group(1)) + 4))
        dist = None
        for j in range(i + 1, min(i + 400, len(sass))):
            if dest_regs(sass[j]) & data:
                dist = j - i
                break
            if sass[j].startswith("BRA") and "@" not in sass[j]:
                break
        print(f"{i:6d}  {ins[:60]:60s} next write of data regs: +{dist}" + (f"   [{sass[i + dist][:50]}]" if dist else ""))


if __name__ == "__main__":
    main()
