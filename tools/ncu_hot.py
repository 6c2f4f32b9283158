#!/usr/bin/env python
"""Summarise the SASS source page of an .ncu-rep: samples per instruction, grouped in
windows, so that the hot region of a kernel can be located without a GUI.

    python tools/ncu_hot.py report.ncu-rep [--win 40] [--top 15] [--cuda]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    win = int(sys.argv[sys.argv.index("--win") + 1]) if "--win" in sys.argv else 40
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 15
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
    if "--cuda" in sys.argv:
        cmd += ["--print-source", "cuda"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # several kernels may follow each other: split on "Kernel Name" rows
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    for blk in blocks[:1] if "--all" not in sys.argv else blocks:
        h = blk["hdr"]
        i_src = h.index("Source")
        i_smp = h.index("# Samples")
        i_ex = h.index("Instructions Executed")
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        total = sum(int(r[i_smp] or 0) for r in blk["rows"])
        print("==", blk["name"], "instructions", len(blk["rows"]), "samples", total)
        wins = []
        for s in range(0, len(blk["rows"]), win):
            seg = blk["rows"][s:s + win]
            n = sum(int(r[i_smp] or 0) for r in seg)
            st = {}
            for i, c in stall_cols:
                st[c] = sum(int(r[i] or 0) for r in seg)
            wins.append((n, s, seg, st))
        for n, s, seg, st in sorted(wins, key=lambda w: -w[0])[:top]:
            tops = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            hot = max(seg, key=lambda r: int(r[i_smp] or 0))
            print("  [%5d..%5d] %5.1f%%  exec/instr~%s  %s | hottest: %s (%s)" % (
                s, s + len(seg), 100.0 * n / max(total, 1), seg[len(seg) // 2][i_ex],
                " ".join("%s=%d" % (k[6:], v) for k, v in tops), hot[i_src].strip()[:60], hot[i_smp]))


if __name__ == "__main__":
    main()
