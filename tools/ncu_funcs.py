"""Aggregate an ncu source-page CSV by device function (line ranges found by scanning lcqp_device.cuh)."""
import csv, collections, re, sys
path = sys.argv[1]
src = open('/root/repo/lcqpow_b200/csrc/lcqp_device.cuh').read().split('\n')
# function starts: lines beginning with LCQ_DEV / LCQ_DEVN / inline LCQ_HD
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r'^(?:LCQ_DEVN|LCQ_DEV|inline LCQ_HD)\s+[\w:<>\*& ]+?\s+\**(\w+)\(', l)
    if m: starts.append((i, m.group(1)))
def func_of(ln):
    name = 'device:?'
    for s, nm in starts:
        if s <= ln: name = nm
        else: break
    return name
rows = list(csv.reader(open(path)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
S = collections.Counter(); I = collections.Counter(); tot = tots = 0
for n, hi in enumerate(his):
    hdr = rows[hi]; fpath = rows[hi - 2][1]
    iSamp = hdr.index('# Samples'); iInst = hdr.index('Instructions Executed')
    end = his[n + 1] - 2 if n + 1 < len(his) else len(rows)
    for r in rows[hi + 1:end]:
        if len(r) <= iInst or not r[0].strip(): continue
        try: ins = int(r[iInst] or 0); sm = int(r[iSamp] or 0); ln = int(r[0])
        except ValueError: continue
        f = fpath.split('/')[-1]
        key = func_of(ln) if f == 'lcqp_device.cuh' else f
        S[key] += sm; I[key] += ins; tot += ins; tots += sm
print('inst', tot, 'samples', tots)
for k, v in S.most_common(25):
    print(f"{k:28s} {100*v/tots:5.1f}% samp {100*I[k]/tot:5.1f}% inst")
