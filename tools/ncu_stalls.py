#!/usr/bin/env python3
"""Per-region stall-sample summary of an ensemble-kernel ncu report (needs --set full): how the warp-stall samples split
between the hot loop and the service section, the stall reasons, and the instructions that collect the most samples.
Usage: tools/ncu_stalls.py report.ncu-rep [top]"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, A, SRC, IE, TH = ix['# Samples'], ix['Address'], ix['Source'], ix['Instructions Executed'], ix['Avg. Threads Executed']
def f(x):
    try: return float(x)
    except ValueError: return 0.0
addr = [int(r[A], 16) for r in data]
tot = sum(f(r[S]) for r in data); it = sum(f(r[IE]) for r in data)
best = None
for i, r in enumerate(data):
    m = re.search(r'BRA[^;]* 0x([0-9a-f]+)', r[SRC])
    if m:
        t = int(m.group(1), 16); a = addr[i]
        if t < a and t - addr[0] > 0x1000 and (best is None or a - t > best[1] - best[0]): best = (t, a)
print("total samples %d, %d SASS instructions; hot loop at +0x%x..+0x%x" % (tot, len(data), best[0] - addr[0], best[1] - addr[0]))
reg = {"service section (before the loop)": [r for i, r in enumerate(data) if addr[i] < best[0]],
       "hot loop": [r for i, r in enumerate(data) if best[0] <= addr[i] <= best[1]],
       "after the loop (out-of-line slow paths)": [r for i, r in enumerate(data) if addr[i] > best[1]]}
for name, rr in reg.items():
    ie = sum(f(r[IE]) for r in rr)
    print("  %-42s samples %5.1f%%  warp instructions %5.1f%%  avg active threads %.1f" % (name, sum(f(r[S]) for r in rr) / tot * 100, ie / it * 100,
          sum(f(r[IE]) * f(r[TH]) for r in rr) / max(ie, 1)))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for name, rr in reg.items():
    agg = collections.Counter()
    for r in rr:
        for h in stalls: agg[h] += f(r[ix[h]])
    t = sum(agg.values())
    print("  %s: " % name + " ".join("%s %.1f%%" % (k[6:], v / max(t, 1) * 100) for k, v in agg.most_common(8)))
print("top instructions by stall samples:")
for r in sorted(data, key=lambda r: -f(r[S]))[:topn]:
    where = "loop" if best[0] <= int(r[A], 16) <= best[1] else "svc " if int(r[A], 16) < best[0] else "post"
    dom = max(stalls, key=lambda h: f(r[ix[h]]))
    print("  %5.2f%% %s +0x%05x %-64s exec=%s thr=%s %s" % (f(r[S]) / tot * 100, where, int(r[A], 16) - addr[0], r[SRC].strip()[:64], r[IE], r[TH], dom[6:]))
