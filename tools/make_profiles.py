#!/usr/bin/env python
"""Turn the ncu captures under gpurun_out/ into the tracked summaries under profiles/:
    profiles/r1_launches_step.csv, r1_launch_shares.md, r1_ncu_full_summary.txt, traffic.json
    python tools/make_profiles.py
"""
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def main():
    lines = [l for l in open(os.path.join(G, "launches_r1.csv")) if not l.startswith("==")]
    open(os.path.join(P, "r1_launches_step.csv"), "w").writelines(lines)
    rows = list(csv.DictReader(lines))
    ours = [r for r in rows if ("spec::" in r["Kernel Name"] or "ctc::" in r["Kernel Name"])]
    last = ours[-4:]      # spectrogram, stats, z-score, fused CTC
    tot = sum(float(r["Metric Value"]) for r in last)
    md = ["# One step of the hot path (C2 batch, 256 utterances), ncu --metrics gpu__time_duration.sum --clock-control none",
          "# (cold-cache, serialised launches: compare SHARES, not absolutes). Source: profiles/r1_launches_step.csv",
          "", "| kernel | grid | block | ns | share |", "|---|---|---|---|---|"]
    for r in last:
        md.append("| %s | %s | %s | %s | %.1f %% |" % (r["Kernel Name"].replace("|", "/"), r["Grid Size"], r["Block Size"],
                                                        r["Metric Value"], 100 * float(r["Metric Value"]) / tot))
    md.append("| total | | | %d | |" % tot)
    open(os.path.join(P, "r1_launch_shares.md"), "w").write("\n".join(md) + "\n")
    rep = os.path.join(G, "prof_step_r1.ncu-rep")
    summ = subprocess.run(["python3", os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(P, "r1_ncu_full_summary.txt"), "w").write(
        "# ncu --set full --clock-control none summaries of the step's kernels (C2 batch), tools/ncu_summary.py\n" + summ)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    h, units = rr[0], rr[1]
    i_name, i_r, i_w, i_t = (h.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                                   "gpu__time_duration.sum"))
    i_f = h.index("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    key = {"spectrogram_kernel": "spec_main", "fused_small_kernel": "ctc_fused", "normalize_kernel": "spec_normalize",
           "stats_kernel": "spec_stats"}
    res = {}
    for r in rr[2:]:
        for k, name in key.items():
            if k in r[i_name]:
                rd, wr = float(r[i_r]) * mult[units[i_r]], float(r[i_w]) * mult[units[i_w]]
                res[name] = {"kernel": r[i_name], "dram_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
                             "duration_us_under_ncu": float(r[i_t]), "fp64_pipe_pct": float(r[i_f]),
                             "source": "profiles/r1_ncu_full_summary.txt (gpurun_out/prof_step_r1.ncu-rep)"}
    json.dump(res, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print("\n".join(md))
    print(json.dumps({k: (v["dram_bytes"], v["duration_us_under_ncu"], v["fp64_pipe_pct"]) for k, v in res.items()}))


if __name__ == "__main__":
    main()
