"""Opcode histogram (executed instructions, stall samples) from `ncu -i X.ncu-rep --page source --csv`."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]
i_src, i_ex, i_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ex, sa = collections.Counter(), collections.Counter()
tot_e = tot_s = 0
body = []
for r in rows[h + 1:]:
    if len(r) <= i_samp: continue
    op = r[i_src].split()
    if not op: continue
    o = op[1] if op[0].startswith("@") else op[0]
    o = o.rstrip(";")
    e, s = int(r[i_ex] or 0), int(r[i_samp] or 0)
    ex[o] += e; sa[o] += s; tot_e += e; tot_s += s
    body.append((s, e, r[i_src].strip()))
print("total executed warp-instructions %d, samples %d" % (tot_e, tot_s))
for o, e in ex.most_common(28):
    print("  %-28s exec %6.2f %%   samples %6.2f %%" % (o, 100.0 * e / tot_e, 100.0 * sa[o] / max(tot_s, 1)))
if len(sys.argv) > 2:
    print("hottest instructions by stall samples:")
    for s, e, src in sorted(body, reverse=True)[:int(sys.argv[2])]:
        print("  %6d samples  %10d exec   %s" % (s, e, src))
