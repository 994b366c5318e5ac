"""Turns an `ncu --set full` report into the short text summary kept under profiles/: headline metrics of the launch plus the
instructions with the most warp-stall samples.   python tools/ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "lg_throttle", "mio_throttle", "math_pipe_throttle", "membar", "not_selected",
          "no_instruction", "branch_resolving", "dispatch_stall", "selected"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr = next(i for i, r in enumerate(raw) if r and r[0] == "ID")
    h, u = raw[hdr], raw[hdr + 1]
    for v in raw[hdr + 2:]:
        if len(v) != len(h):
            continue
        d = dict(zip(h, zip(v, u)))
        print("kernel: %s   grid %s block %s" % (d["Kernel Name"][0][:110], d["Grid Size"][0], d["Block Size"][0]))
        for k in KEYS:
            if k in d:
                print("  %-72s %s %s" % (k, d[k][0], d[k][1]))
        print("  warp stalls per issued instruction:")
        for s in STALLS:
            k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            if k in d:
                print("    %-22s %s" % (s, d[k][0]))
    src = page(rep, "source")
    hdr = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    h = src[hdr]
    ix = {n: i for i, n in enumerate(h)}
    rows = [r for r in src[hdr + 1:] if len(r) == len(h)]
    tot = sum(int(r[ix["# Samples"]]) for r in rows)
    print("  top stall sites (of %d samples over %d SASS instructions):" % (tot, len(rows)))
    top = sorted(enumerate(rows), key=lambda t: -int(t[1][ix["# Samples"]]))[:18]
    for i, r in sorted(top):
        reasons = {k[6:]: int(r[ix[k]]) for k in h if k.startswith("stall_") and "(Not" not in k and r[ix[k]].isdigit() and int(r[ix[k]]) > 0}
        main_r = ", ".join("%s %d" % kv for kv in sorted(reasons.items(), key=lambda kv: -kv[1])[:3])
        print("    #%-5d %-58s samples %-4s executed %-7s %s" % (i, r[ix["Source"]].strip()[:58], r[ix["# Samples"]], r[ix["Instructions Executed"]], main_r))


if __name__ == "__main__":
    main()
