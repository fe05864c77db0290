"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by source line / file."""
import csv, collections, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
tot = tots = 0
byline = collections.Counter(); sampline = collections.Counter(); srctext = {}
for n, hi in enumerate(his):
    hdr = rows[hi]
    fpath = rows[hi - 2][1] if hi >= 2 else '?'
    iSamp = hdr.index('# Samples'); iInst = hdr.index('Instructions Executed')
    end = his[n + 1] - 2 if n + 1 < len(his) else len(rows)
    for r in rows[hi + 1:end]:
        if len(r) <= iInst: continue
        try:
            ins = int(r[iInst] or 0); sm = int(r[iSamp] or 0)
        except ValueError:
            continue
        key = (fpath.split('/')[-1], r[0])
        byline[key] += ins; sampline[key] += sm; srctext[key] = r[1]
        tot += ins; tots += sm
print('total inst', tot, 'samples', tots)
for k, v in sampline.most_common(top):
    print(f"{k[0]}:{k[1]:>5} {100*v/max(tots,1):5.1f}% samp {100*byline[k]/max(tot,1):5.1f}% inst  {srctext[k].strip()[:120]}")
