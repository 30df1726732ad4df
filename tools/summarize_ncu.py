"""Text summaries of ncu captures for profiles/ (run in the build container; ncu can read reports without a GPU).
  python tools/summarize_ncu.py launches <launches.csv> [first_id]     per-kernel share of one sigma
  python tools/summarize_ncu.py report <file.ncu-rep>                  key counters + SASS opcode mix per launch"""
import collections, csv, re, subprocess, sys

KEYS = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum"]


def launches(path, first=None):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]; i_name, i_val, i_id, i_grid, i_unit = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "ID", "Grid Size", "Metric Unit"))
    seq = []
    for r in rows[1:]:
        v = float(r[i_val].replace(",", "")); u = r[i_unit]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        seq.append((int(r[i_id]), re.sub(r"\(.*", "", r[i_name]).replace("void ", ""), r[i_grid], v))
    n = len(seq)
    first = n // 2 if first is None else first
    half = [s for s in seq if s[0] >= first]
    t = sum(s[3] for s in half)
    print("launches in file: %d; last sigma evaluation = ids %d..%d: %d launches, %.1f us (serialised, cold cache)" % (n, first, n - 1, len(half), t))
    for s in half:
        if s[3] > 0.01 * t:
            print("  id %3d %-34s grid %-14s %9.1f us  %5.1f %%" % (s[0], s[1], s[2], s[3], 100 * s[3] / t))
    agg = collections.defaultdict(float)
    for s in half:
        agg[s[1]] += s[3]
    print("per kernel:")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print("  %-34s %9.1f us %5.1f %%" % (k, v, 100 * v / t))


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==== %s" % r[idx["Kernel Name"]])
        for k in KEYS:
            if k in idx:
                print("  %-82s %s %s" % (k, r[idx[k]], units[idx[k]]))
        for h in hdr:
            if "issue_stalled" in h and h.endswith("per_warp_active.pct") and float(r[idx[h]] or 0) > 5:
                print("  stall %-76s %.1f %%" % (h.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", ""), float(r[idx[h]])))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    for si, h in enumerate(starts):
        hdr = rows[h]
        i_src, i_ex, i_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        ex, sa = collections.Counter(), collections.Counter()
        te = ts = 0
        end = starts[si + 1] - 1 if si + 1 < len(starts) else len(rows)
        for r in rows[h + 1:end]:
            if len(r) <= i_samp:
                continue
            op = r[i_src].split()
            if not op:
                continue
            o = (op[1] if op[0].startswith("@") else op[0]).rstrip(";")
            try:
                e, s_ = int(r[i_ex] or 0), int(r[i_samp] or 0)
            except ValueError:
                continue
            ex[o] += e; sa[o] += s_; te += e; ts += s_
        print("---- SASS mix of launch %d: %d warp instructions, %d stall samples" % (si, te, ts))
        for o, e in ex.most_common(12):
            print("  %-24s exec %6.2f %%   samples %6.2f %%" % (o, 100.0 * e / max(te, 1), 100.0 * sa[o] / max(ts, 1)))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
    else:
        report(sys.argv[2])
