"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` by device function and by line.
usage: python tools/ncu_funcs.py dump.csv [top_lines]"""
import csv, collections, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = open('/root/repo/lcqpow_b200/csrc/lcqp_device.cuh').read().split('\n')
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r'^(?:LCQ_DEVN|LCQ_DEV|LCQ_SYM_INL|LCQ_R1_INL|inline LCQ_HD)\s+[\w:<>\*& ]+?\s+\**(\w+)\(', l)
    if m: starts.append((i, m.group(1)))
def func_of(ln):
    name = '?'
    for s, nm in starts:
        if s <= ln: name = nm
        else: break
    return name
fpath = None; hdr = None
S = collections.Counter(); I = collections.Counter(); LS = collections.Counter(); LI = collections.Counter(); txt = {}
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed'); continue
    if hdr is None or not r[0].strip(): continue
    try: ln = int(r[0]); sm = int(r[iS] or 0); ins = int(r[iI] or 0)
    except ValueError: continue
    key = func_of(ln) if fpath == 'lcqp_device.cuh' else fpath
    S[key] += sm; I[key] += ins; LS[(fpath, ln)] += sm; LI[(fpath, ln)] += ins; txt[(fpath, ln)] = r[1]
tots = sum(S.values()); tot = sum(I.values())
print("samples", tots, "inst", tot)
for k, v in S.most_common(30): print(f"{k:28s} {100*v/tots:5.1f}% samp {100*I[k]/tot:5.1f}% inst")
print()
for k, v in LS.most_common(top): print(f"{k[0]}:{k[1]:>5} {100*v/tots:5.1f}% samp {100*LI[k]/tot:5.1f}% inst  {txt[k].strip()[:110]}")
