"""Static evidence for profiles/: per-kernel resources from `ptxas -v` (registers, spills, static shared memory) and
SASS mnemonic counts from `cuobjdump -sass` of the built library (bulk-copy / mbarrier / async-copy / fp64 / MUFU
instructions per kernel).  Runs without a GPU.

  python tools/static_summary.py > profiles/r2_static_resources.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "asr_dfcnn_transformer_b200", "csrc")
LIB = os.path.join(ROOT, "asr_dfcnn_transformer_b200", "libasrk.so")
SOURCES = ["spectrogram.cu", "ctc.cu", "noise.cu", "color_noise.cu", "logfbank.cu", "post.cu"]
WATCH = ["UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DADD", "DMUL", "MUFU", "REDUX", "MATCH", "BAR", "SHFL", "LDS", "STS",
         "LDG", "STG", "ATOM", "RED", "FENCE", "MEMBAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(sig):
    sig = re.sub(r"^void ", "", sig)
    return re.sub(r"\(.*$", "", sig).replace("asrk::", "")


def ptxas():
    rows = []
    for src in SOURCES:
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xptxas", "-v",
               "-diag-suppress", "177", "-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, src), "-o", os.devnull]
        err = subprocess.run(cmd, capture_output=True, text=True).stderr
        cur = None
        for line in err.splitlines():
            m = re.search(r"Compiling entry function '([^']+)'", line)
            if m:
                cur = dict(name=m.group(1), src=src, spill="0/0", stack=0, regs=None, smem=0)
                rows.append(cur)
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur["regs"] is None:
                cur["stack"], cur["spill"] = int(m.group(1)), "%s/%s" % (m.group(2), m.group(3))
            m = re.search(r"Used (\d+) registers", line)
            if m:
                cur["regs"] = int(m.group(1))
                m2 = re.search(r"(\d+) bytes smem", line)
                cur["smem"] = int(m2.group(1)) if m2 else 0
                m3 = re.search(r"used (\d+) barriers", line)
                cur["bars"] = int(m3.group(1)) if m3 else 0
    names = demangle([r["name"] for r in rows])
    print("== ptxas -v (sm_100a): registers per thread, static shared memory, spills ==")
    print("%-58s %-15s %5s %8s %12s %5s" % ("kernel", "source", "regs", "smem(B)", "spill st/ld", "bars"))
    for r in rows:
        print("%-58s %-15s %5s %8d %12s %5s" % (short(names[r["name"]])[:58], r["src"], r["regs"], r["smem"], r["spill"],
                                                r.get("bars", 0)))


def sass():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w):
                    counts[cur][w] += 1
                    break
    names = demangle(list(counts))
    print()
    print("== cuobjdump -sass of libasrk.so: instruction counts per kernel (static) ==")
    cols = ["total"] + WATCH
    print("%-46s " % "kernel" + " ".join("%6s" % c[:6] for c in cols))
    for k, c in counts.items():
        print("%-46s " % short(names[k])[:46] + " ".join("%6d" % c[x] for x in cols))
    print()
    print("UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier arrive/try_wait, LDGSTS = cp.async, "
          "D* = fp64 pipe, MUFU = special-function unit.")


if __name__ == "__main__":
    ptxas()
    if os.path.isfile(LIB):
        sass()
    else:
        sys.exit("build the library first: python -m asr_dfcnn_transformer_b200._build")
