import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        k=(d['Kernel Name'][:48], d['Grid Size'])
        agg.setdefault(k,[]).append(float(d['Metric Value'])/1000)
for k,v in agg.items(): print("%-50s %-16s n=%d  avg %.1f us  min %.1f" % (k[0],k[1],len(v),sum(v)/len(v),min(v)))
