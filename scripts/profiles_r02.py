"""Regenerate profiles/r02_{chain16,pp,train}_ncu.md and profiles/r02_traffic.json from the captures scripts/ncu_r02.sh left in
gpurun_out/ (ncu --page raw --csv exports)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def table(name):
    rows = list(csv.reader(open(os.path.join(G, f"prof_r02_{name}_raw.csv"))))
    hdr, units = rows[0], rows[1]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    out, traffic = [], []
    for r in rows[2:]:
        out.append("| metric | value | unit |\n|---|---|---|")
        for w, i in cols:
            out.append(f"| {w} | {r[i]} | {units[i]} |")
        out.append("")
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        traffic.append(float(r[rd]) * UNIT[units[rd]] + float(r[wr]) * UNIT[units[wr]])
    with open(os.path.join(P, f"r02_{name}_ncu.md"), "w") as f:
        f.write(f"# ncu --set full, round 2: {name} (per launch; cold-cache, serialised: use for shares and counters, not absolute times)\n\n")
        f.write("\n".join(out))
    return traffic


t16 = table("chain16")
tpp = table("pp")
table("train")
if os.path.exists(os.path.join(G, "prof_r02_stage_raw.csv")):
    table("stage")
traffic = {"bounded16:16777216": t16[0], "two_moons_conditional:1000000": tpp[0]}
if os.path.exists(os.path.join(G, "prof_r02_stage_bench_raw.csv")):   # the bench's own stage leg (d = 8, K = 32, 493,421 events)
    traffic["rqs_stage:493421"] = table("stage_bench")[0]
json.dump(traffic, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, "r02_traffic.json")).read())
