import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hi]; ix={h:i for i,h in enumerate(H)}
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<len(H): continue
    k=r[ix['Kernel Name']].split('(')[0].replace('void ','')
    m=r[ix['Metric Name']]; v=float(r[ix['Metric Value']].replace(',',''))
    agg.setdefault(k,{}).setdefault(m,[]).append(v)
for k,a in agg.items():
    t=a.get('gpu__time_duration.sum',[0])
    f=lambda m: sum(a.get(m,[0]))/max(1,len(t))
    print("%-28s n=%4d total %8.2f ms  avg %7.1f us | thr/inst %.1f  warps %.1f%%  issue %.1f%%"%(k,len(t),sum(t)/1e6,sum(t)/len(t)/1e3,
        f('smsp__thread_inst_executed_per_inst_executed.ratio'), f('sm__warps_active.avg.pct_of_peak_sustained_active'), f('smsp__issue_active.avg.pct_of_peak_sustained_active')),
        ("dram rd %.1f MB/launch  L2 hit %.0f%%" % (f('dram__bytes_read.sum')/1e6 if max(a.get('dram__bytes_read.sum',[0]))>1e4 else f('dram__bytes_read.sum'), f('lts__t_sector_hit_rate.pct'))) if 'dram__bytes_read.sum' in a else "")
