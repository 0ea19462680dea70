"""Writes profiles/ncu_r01_hot_kernels.txt and profiles/ncu_traffic.json from the ncu reports of
profiles/run_gpu_r01_final.sh (gpurun_out/prof_r01_final*.ncu-rep; the reports themselves are scratch).

usage: python profiles/summarize_ncu.py"""
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return int(float(v) * scale)


def main():
    """python profiles/summarize_ncu.py                       -> round 1 files (as committed)
       python profiles/summarize_ncu.py r02 U64.ncu-rep ZIPF.ncu-rep [script name]
                                                              -> profiles/ncu_r02_hot_kernels.txt, ncu_traffic_r02.json"""
    import sys
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    if tag == "r01":
        reps = (("config 2: 1e8 uniform u64 pairs x 16 B (bench.py default)", "prof_r01_final.ncu-rep", "u64"),
                ("config 3 shape at 2e8 Zipf(1.1) string pairs x 32 B (combiner on)", "prof_r01_final_zipf.ncu-rep", "zipf32"))
        script = "profiles/run_gpu_r01_final.sh"
    else:
        reps = (("config 2: 1e8 uniform u64 pairs x 16 B (ride-along block of the bench line)", sys.argv[2], "u64"),
                ("config 3: 1e9 Zipf(1.1) string pairs x 32 B, combiner on (the bench headline, FULL size)", sys.argv[3], "zipf32"))
        script = sys.argv[4] if len(sys.argv) > 4 else "profiles/run_gpu_%s_final.sh" % tag
    lines = ["# ncu --set full --clock-control none, one launch per hot kernel, %s (%s)" % (tag, script),
             "# serialised cold-cache launches: compare shares, not absolutes; never used as bench values", ""]
    traffic, traffic2 = {}, {}
    seen = {}
    for title, rep, wl in reps:
        path = rep if os.path.isabs(rep) or os.path.exists(rep) else os.path.join(ROOT, "gpurun_out", rep)
        if not os.path.exists(path):
            continue
        hdr, units, rows = rows_of(path)
        lines.append("# ==== " + title)
        for r in rows:
            name = r[hdr.index("Kernel Name")].replace("void ", "").split("(")[0]
            grid = r[hdr.index("launch__grid_size")]
            key = (rep, name.split("<")[0] if name.startswith("k_split") else name)
            seen[key] = seen.get(key, 0) + 1
            if name.startswith("k_split") and seen[key] <= 2:
                label = "%s level %d (grid %s)" % (name, seen[key], grid)
            elif seen[key] > 1:
                continue
            else:
                label = name
            lines.append("## " + label)
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    lines.append("%-90s %20s %s" % (m, r[i], units[i]))
            lines.append("")
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            t = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
            short = name.split("<")[0]
            traffic2.setdefault(wl, {})
            if short == "k_split_tma":
                traffic2[wl]["k_split_tma_level%d" % seen[key]] = t
                traffic2[wl].setdefault("k_split_tma", t)
            else:
                traffic2[wl][{"k_sort_reduce_u64": "k_sort_reduce"}.get(short, short)] = t
            if wl == "u64":
                if name.startswith("k_split"):
                    traffic["k_split_level%d" % seen[key]] = t
                elif name.startswith("k_sort_reduce"):
                    traffic["k_sort_reduce"] = t
    open(os.path.join(ROOT, "profiles", "ncu_%s_hot_kernels.txt" % tag), "w").write("\n".join(lines) + "\n")
    if tag == "r01":
        json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    else:
        json.dump(traffic2, open(os.path.join(ROOT, "profiles", "ncu_traffic_%s.json" % tag), "w"), indent=1)
    print(traffic if tag == "r01" else traffic2)


if __name__ == "__main__":
    main()
