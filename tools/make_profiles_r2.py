#!/usr/bin/env python
"""Round 2: turn the ncu captures under gpurun_out/ (tools/r2_run16.sh) into the tracked summaries under profiles/:
    r2_launches_<workload>.csv   launch lists of three steps (ncu --metrics gpu__time_duration.sum --clock-control none)
    r2_launch_shares.md          last step of each list: kernel, grid, block, ns, share
    r2_ncu_full_summary.txt      --set full summaries (tools/ncu_summary.py) of the C2, C2 keras, C3 and C4 kernels
    traffic.json                 per-launch DRAM bytes / duration / fp64 pipe of every kernel (bench.py reads it)
    python tools/make_profiles_r2.py
"""
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

WORKLOADS = [("c2", "C2 (256 utterances, 3-7 s; logits surface)"), ("c2keras", "C2, Keras surface (probabilities in)"),
             ("c2merged", "C2 with the merged tail (--merged-tail: z-score as co-work of the fused CTC kernel)"),
             ("c3", "C3 (64 utterances of 20 s, T = 1998, L ~ 300: generic CTC kernels)"),
             ("c4", "C4 (512 float32 utterances + noise, SNR2K + fused mix, features only)")]
OURS = ("spec::", "ctc::", "noise::")
REPS = [("prof_step_r2", "c2"), ("prof_keras_r2", "c2keras"), ("prof_merged_r2", "c2merged"), ("prof_c3_r2", "c3"),
        ("prof_c4_r2", "c4")]
KEYS = {"spectrogram_kernel<0>": "spec_main", "spectrogram_kernel<1>": "spec_main_f32", "fused_small_kernel": "ctc_fused",
        "normalize_kernel": "spec_normalize", "stats_kernel": "spec_stats", "rows_kernel": "ctc_rows",
        "lattice_kernel": "ctc_lattice", "grad_kernel": "ctc_grad", "snr2k_kernel": "snr2k"}


def last_step(rows):
    """kernels of the last step in the list: walk back from the end to the last launch of the step's first kernel"""
    ours = [r for r in rows if any(o in r["Kernel Name"] for o in OURS)]
    names = [r["Kernel Name"] for r in ours]
    seen, out = set(), []
    for r in reversed(ours):
        if r["Kernel Name"] in seen:
            break
        seen.add(r["Kernel Name"])
        out.append(r)
    return list(reversed(out)), len(names)


def main():
    md = ["# One step of the hot path per workload: ncu --metrics gpu__time_duration.sum --clock-control none",
          "# (cold-cache, serialised launches: compare SHARES, not absolutes).  Sources: profiles/r2_launches_<workload>.csv,",
          "# produced by tools/r2_run16.sh on a B200 after the same command had exited 0 without ncu.", ""]
    for w, title in WORKLOADS:
        lines = [l for l in open(os.path.join(G, "launches_r2_%s.csv" % w)) if not l.startswith("==")]
        open(os.path.join(P, "r2_launches_%s.csv" % w), "w").writelines(lines)
        step, n = last_step(list(csv.DictReader(lines)))
        tot = sum(float(r["Metric Value"]) for r in step)
        md += ["## " + title, "", "| kernel | grid | block | ns | share |", "|---|---|---|---|---|"]
        for r in step:
            md.append("| %s | %s | %s | %s | %.1f %% |" % (r["Kernel Name"].replace("|", "/")[:80], r["Grid Size"],
                                                            r["Block Size"], r["Metric Value"],
                                                            100 * float(r["Metric Value"]) / tot))
        md += ["| total | | | %d | |" % tot, ""]
    open(os.path.join(P, "r2_launch_shares.md"), "w").write("\n".join(md) + "\n")

    summ = ["# ncu --set full --clock-control none --import-source on, one launch per kernel (tools/ncu_summary.py)"]
    res = {}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tmult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
    for rep, w in REPS:
        path = os.path.join(G, rep + ".ncu-rep")
        s = subprocess.run(["python3", os.path.join(ROOT, "tools", "ncu_summary.py"), path], capture_output=True, text=True).stdout
        summ += ["", "===== %s (gpurun_out/%s.ncu-rep) =====" % (w, rep), s]
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(out.splitlines()))
        h, units = rr[0], rr[1]
        ix = {k: h.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                      "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                                      "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                      "smsp__issue_active.avg.pct_of_peak_sustained_active",
                                      "lts__t_sector_hit_rate.pct")}
        for r in rr[2:]:
            for k, name in KEYS.items():
                if k in r[ix["Kernel Name"]]:
                    key = name if w in ("c2",) or name.startswith(("ctc_rows", "ctc_lattice", "ctc_grad", "snr2k", "spec_main_f32")) \
                        else "%s_%s" % (name, w)
                    if key in res:
                        continue
                    rd = float(r[ix["dram__bytes_read.sum"]]) * mult[units[ix["dram__bytes_read.sum"]]]
                    wr = float(r[ix["dram__bytes_write.sum"]]) * mult[units[ix["dram__bytes_write.sum"]]]
                    res[key] = {"kernel": r[ix["Kernel Name"]], "workload": w, "dram_bytes": rd + wr, "dram_read_bytes": rd,
                                "dram_write_bytes": wr,
                                "duration_us_under_ncu": float(r[ix["gpu__time_duration.sum"]]) *
                                tmult.get(units[ix["gpu__time_duration.sum"]], 1.0),
                                "fp64_pipe_pct": float(r[ix["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]]),
                                "dram_throughput_pct": float(r[ix["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                                "issue_active_pct": float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                                "l2_hit_pct": float(r[ix["lts__t_sector_hit_rate.pct"]]),
                                "source": "profiles/r2_ncu_full_summary.txt (gpurun_out/%s.ncu-rep)" % rep}
    # bench.py looks kernels up by their C++ name too
    for alias, key in (("spectrogram_kernel", "spec_main"), ("fused_small_kernel", "ctc_fused"), ("rows_kernel", "ctc_rows"),
                       ("lattice_kernel", "ctc_lattice"), ("grad_kernel", "ctc_grad")):
        if key in res:
            res[alias] = res[key]
    open(os.path.join(P, "r2_ncu_full_summary.txt"), "w").write("\n".join(summ) + "\n")
    json.dump(res, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print("\n".join(md))
    for k, v in res.items():
        print("%-22s %8.1f MB  %8.1f us  fp64 %5.1f %%  dram %5.1f %%  issue %5.1f %%  L2 hit %5.1f %%" % (
            k, v["dram_bytes"] / 1e6, v["duration_us_under_ncu"], v["fp64_pipe_pct"], v["dram_throughput_pct"],
            v["issue_active_pct"], v["l2_hit_pct"]))


if __name__ == "__main__":
    main()
