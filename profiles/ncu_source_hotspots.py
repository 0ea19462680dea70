"""Per-source-line stall samples of one kernel from an ncu report captured with --import-source on.

usage: python profiles/ncu_source_hotspots.py REPORT.ncu-rep KERNEL_REGEX [TOP_N]
(reads the report with `ncu -i ... --page source --csv --print-source cuda,sass`)"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
    cur, hdr, data, inst = None, None, {}, -1
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "Kernel Name":
            inst += 1
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Function Name":
            continue
        if hdr and r[0] != "" and len(r) == len(hdr):
            try:
                s = int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            data.setdefault(inst, []).append((s, cur, r))
    cols = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_lg", "stall_wait", "stall_math",
            "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_no_inst", "stall_dispatch",
            "stall_membar", "stall_sleep", "stall_drain", "stall_tex", "stall_misc"]
    for k, d in data.items():
        tot = sum(x[0] for x in d)
        print("== launch %d of %s: %d samples" % (k, rx, tot))
        agg = {c: sum(int(x[2][hdr.index(c)] or 0) for x in d) for c in cols}
        print("   by reason: " + ", ".join("%s %.1f%%" % (c[6:], 100.0 * v / tot) for c, v in sorted(agg.items(), key=lambda t: -t[1]) if v * 200 > tot))
        for s, f, r in sorted(d, key=lambda x: -x[0])[:top]:
            why = max(cols, key=lambda c: int(r[hdr.index(c)] or 0))
            print("%6.1f%% %-12s inst=%9s | %s:%s: %s" % (100.0 * s / tot, why[6:], r[hdr.index("Instructions Executed")], f, r[0], r[1].strip()[:110]))


if __name__ == "__main__":
    main()
