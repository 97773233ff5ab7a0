#!/usr/bin/env python3
"""Turn one gpurun round's artefacts (gpurun_out/*_<tag>.*) into the tracked summaries under profiles/.
   python tools/summarize_profiles.py <tag> <round-prefix> [n_signatures_per_profiled_launch]"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.per_cycle_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]


def main():
    tag, pre = sys.argv[1], sys.argv[2]
    nsig = int(sys.argv[3]) if len(sys.argv) > 3 else 303104
    for src, dst in ((f"bench_{tag}.json", f"{pre}_bench.json"), (f"bench_ref_{tag}.json", f"{pre}_bench_reference_arm.json"),
                     (f"launches_{tag}.csv", f"{pre}_ncu_launches.csv"), (f"gpu_{tag}.txt", f"{pre}_gpu_box.txt")):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    rep = os.path.join(G, f"prof_{tag}.ncu-rep")
    # launch-list shares
    rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        agg[r[ki].split("(")[0]][0] += 1
        agg[r[ki].split("(")[0]][1] += float(r[vi].replace(",", "")) / 1e6
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"{pre}_ncu_launch_shares.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3 --no-cpu --pool 32768\n")
        f.write("# launches   total ms   share   kernel\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%6d %10.3f %6.1f%%  %s\n" % (v[0], v[1], 100 * v[1] / tot, k))
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        h, u = rr[0], rr[1]
        idx = {x: i for i, x in enumerate(h)}
        out = [f"# ncu --set full --clock-control none, {nsig} signatures per launch (tools/prof_run.py), one column per captured launch",
               "# kernels: " + " | ".join(r[idx["Kernel Name"]][:44] for r in rr[2:])]
        for m in WANT:
            if m in idx:
                out.append("%-88s %-16s %s" % (m, u[idx[m]], [r[idx[m]] for r in rr[2:]]))
        open(os.path.join(P, f"{pre}_ncu_full_summary.txt"), "w").write("\n".join(out) + "\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
        tmp = "/tmp/_src.csv"
        open(tmp, "w").write(src)
        mix = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_mix.py"), tmp, str(nsig)], capture_output=True, text=True).stdout
        open(os.path.join(P, f"{pre}_ncu_instruction_mix.txt"), "w").write(mix)
        # DRAM traffic of the headline kernel, scaled to the bench's launch size
        k1 = [r for r in rr[2:] if "CurveK1" in r[idx["Kernel Name"]]][0]

        def mb(name):
            v, unit = float(k1[idx[name]]), u[idx[name]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

        per_sig = (mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")) / nsig
        json.dump({"source": f"profiles/{pre}_ncu_full_summary.txt: dram__bytes_read.sum + dram__bytes_write.sum of ecrecover_kernel<CurveK1> at {nsig} signatures per launch",
                   "secp256k1_dram_bytes_per_signature": per_sig, "secp256k1_dram_bytes_per_launch": per_sig * 1048576,
                   "note": "scaled linearly to the bench's 1,048,576-signature launch; algorithmic bytes are 161 B/signature -- the excess is the per-thread affine table {1..8}R (960 B of scratch per signature, written once and read back through L2)"},
                  open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    print(open(os.path.join(P, f"{pre}_ncu_launch_shares.txt")).read())
    if os.path.exists(rep):
        print(open(os.path.join(P, f"{pre}_ncu_full_summary.txt")).read())


if __name__ == "__main__":
    main()
