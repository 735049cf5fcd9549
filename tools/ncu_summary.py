"""Summarise ncu outputs (read on the CPU box):
    python tools/ncu_summary.py launches <launches.csv>      -> per-kernel time share
    python tools/ncu_summary.py rep <file.ncu-rep> [regex]   -> key metrics per captured launch
    python tools/ncu_summary.py source <file.ncu-rep> [topN] -> hottest source lines (stall samples)
    python tools/ncu_summary.py opcodes <file.ncu-rep> [topN] -> dynamic instruction mix
    python tools/ncu_summary.py traffic <k4.ncu-rep> <k1.ncu-rep> [note]
        -> rewrites profiles/ncu_traffic.json (dram bytes per launch of K4 / K1 from `ncu --set full`
           captures at the headline shape), stamped with the commit and a hash of the kernel sources
           so that bench.py drops the figure when those sources change
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ki])[:90]
        tot[name] += float(r[vi].replace(",", ""))
        cnt[name] += 1
    total = sum(tot.values())
    print("%-90s %6s %12s %10s %7s" % ("kernel", "count", "total_ns", "avg_ns", "share"))
    for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("%-90s %6d %12.0f %10.0f %6.1f%%" % (name, cnt[name], t, t / cnt[name], 100 * t / total))


def rep(path, flt=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if flt and not re.search(flt, name):
            continue
        print("==== %s  (id %s)" % (name[:100], r[0]))
        for k in KEYS:
            if k in hdr:
                print("  %-85s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))


def source(path, top=25):
    """Hottest CUDA source lines by warp-stall samples (needs -lineinfo at compile time)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    acc = collections.defaultdict(lambda: [0.0, 0.0, ""])
    fname, hdr = "", None
    for r in csv.reader(io.StringIO(out)):
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 8 or not r[0].isdigit() or r[2] != "-":
            continue
        key = (fname, int(r[0]))
        acc[key][0] += float(r[hdr.index("# Samples")] or 0)
        acc[key][1] += float(r[hdr.index("Instructions Executed")] or 0)
        acc[key][2] = r[1].strip()
    tot = sum(a[0] for a in acc.values()) or 1
    toti = sum(a[1] for a in acc.values()) or 1
    print("samples%  inst%   file:line  source")
    for (f, ln), (s, i, src) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%6.2f%% %6.2f%%  %s:%d  %s" % (100 * s / tot, 100 * i / toti, f, ln, src[:110]))


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2])
    elif mode == "rep":
        rep(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    elif mode == "source":
        source(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)


def opcodes(path, top=25):
    """Dynamic instruction mix: warp-level executed instructions per SASS opcode."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    hdr, cnt, smp = None, collections.Counter(), collections.Counter()
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].startswith("0x"):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)?)", r[1])
        op = m.group(2) if m else r[1][:12]
        op = op.split(".")[0] if not op.startswith(("LDS", "LDG", "STG", "STS", "MUFU")) else op
        cnt[op] += float(r[hdr.index("Instructions Executed")] or 0)
        smp[op] += float(r[hdr.index("# Samples")] or 0)
    tot, tots = sum(cnt.values()) or 1, sum(smp.values()) or 1
    print("%-14s %14s %7s %9s" % ("opcode", "warp-inst", "share", "samples%"))
    for op, c in cnt.most_common(top):
        print("%-14s %14.0f %6.1f%% %8.1f%%" % (op, c, 100 * c / tot, 100 * smp[op] / tots))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "opcodes":
    opcodes(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)


def traffic(k4_rep, k1_rep, note=""):
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench

    def dram_bytes(path, flt):
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        vals = []
        for r in rows[2:]:
            if not re.search(flt, r[hdr.index("Kernel Name")]):
                continue
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v, u = float(r[hdr.index(k)].replace(",", "")), units[hdr.index(k)].lower()
                tot += v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
            vals.append(tot)
        return sum(vals) / len(vals)
    path = os.path.join(root, "profiles", "ncu_traffic.json")
    try:
        t = json.load(open(path))
    except Exception:
        t = {}
    t["k4_bytes_per_launch"] = int(dram_bytes(k4_rep, "bnn_"))
    t["k1_burn_in_bytes_per_launch"] = int(dram_bytes(k1_rep, "sghmc_update"))
    t["k1_burn_in_algorithmic_bytes"] = 44 * 8192 * 5252
    t["k4_algorithmic_bytes"] = 8 * 8192 * 5252
    t["source"] = ("ncu --set full --clock-control none, 8192 chains x D=5252 (%s, %s): dram__bytes_read.sum + "
                   "dram__bytes_write.sum per launch%s" % (os.path.basename(k4_rep), os.path.basename(k1_rep),
                                                          "; " + note if note else ""))
    t["commit"] = subprocess.run(["git", "-C", root, "rev-parse", "--short", "HEAD"], capture_output=True,
                                 text=True).stdout.strip()
    t["source_hash"] = {k: bench.source_hash(v) for k, v in bench.KERNEL_SOURCES.items()}
    json.dump(t, open(path, "w"), indent=1)
    print(json.dumps(t, indent=1))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "traffic":
    traffic(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
