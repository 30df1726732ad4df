"""Per-source-line stall samples / executed instructions of one launch of an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py <file.ncu-rep> [launch_index] [top_n]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; k = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the report has one table per source file; keep the first (the kernel's own file) and per-line rows only
samp2, ex2, text, stalls = collections.Counter(), collections.Counter(), {}, {}
hdr = None; fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; i_samp, i_ex = hdr.index("# Samples"), hdr.index("Instructions Executed")
        st_cols = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if hdr is None or len(r) <= i_ex or not r[0]:
        continue
    try:
        ln = (fname, int(r[0])); samp2[ln] = int(r[i_samp] or 0); ex2[ln] = int(r[i_ex] or 0); text[ln] = r[1]
        stalls[ln] = sorted(((int(r[i] or 0), c[6:]) for i, c in st_cols if i < len(r) and (r[i] or "0").isdigit()), reverse=True)[:3]
    except ValueError:
        pass
ts, te = sum(samp2.values()) or 1, sum(ex2.values()) or 1
print("launch %d: %d samples, %d warp instructions" % (k, ts, te))
for ln, s_ in samp2.most_common(top):
    print("%14s:%-4d samp %5.1f%%  exec %5.1f%%  %-34s %s" % (ln[0][-14:], ln[1], 100.0 * s_ / ts, 100.0 * ex2[ln] / te, " ".join("%s:%d" % (c, v) for v, c in stalls[ln] if v), text.get(ln, "").strip()[:90]))
