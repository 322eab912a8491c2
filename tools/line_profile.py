#!/usr/bin/env python
"""Attribute ncu per-instruction executed counts to source lines through the inline chain.

usage: python tools/line_profile.py <lib.so> <report.ncu-rep> <kernel regex for ncu> <mangled substring> [depth]
Needs the .so the profile was taken with (compiled with -lineinfo). Runs here, no GPU."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    so, rep, kre, mangled = sys.argv[1:5]
    depth = int(sys.argv[5]) if len(sys.argv) > 5 else 2
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # locate function
    start = [i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l][0]
    frames, insts = [], []
    pending = []
    for l in dis[start + 1:]:
        if l.startswith(".text.") or l.startswith("//-----"):
            if insts:
                break
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*)", l)
        if m:
            if pending:
                frames_cur = pending
                pending = []
            insts.append((int(m.group(1), 16), m.group(2), list(frames_cur) if 'frames_cur' in dir() else []))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
    h = rows[hi]
    iI, iN = h.index("Instructions Executed"), h.index("# Samples")
    data = [r for r in rows[hi + 1:] if len(r) == len(h)]
    # the csv may contain the kernel several times (several launches): keep the first len(insts)
    data = data[:len(insts)]
    assert len(data) == len(insts), (len(data), len(insts))
    agg = collections.Counter()
    smp = collections.Counter()
    tot = 0.0
    for (addr, text, fr), r in zip(insts, data):
        ie = float(r[iI] or 0)
        tot += ie
        # frames are innermost first; the LAST is the line in the kernel body
        chain = list(reversed(fr))[:depth]
        key = " <- ".join("%s:%d" % f for f in reversed(chain))
        agg[key] += ie
        smp[key] += float(r[iN] or 0)
    nw = float(os.environ.get("WARP_STEPS", "0"))
    print("total executed warp-instructions %.4g%s" % (tot, ("  = %.1f per warp-substep" % (tot / nw)) if nw else ""))
    ts = sum(smp.values()) or 1
    for k, v in agg.most_common(int(os.environ.get("TOP", "60"))):
        print("%6.2f%% inst %6.2f%% samples %s  %s" % (100 * v / tot, 100 * smp[k] / ts,
                                                      ("%7.1f/step" % (v / nw)) if nw else "", k))


if __name__ == "__main__":
    main()
