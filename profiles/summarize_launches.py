import csv, collections, re, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
per=collections.defaultdict(dict)
for row in csv.DictReader(lines):
    per[(row['ID'],row['Kernel Name'])][row['Metric Name']]=(float(row['Metric Value'].replace(',','')),row['Metric Unit'])
agg=collections.defaultdict(lambda:[0,0.0,0.0])
for (id_,k),d in per.items():
    t,tu=d['gpu__time_duration.sum']; t_us=t/1000 if tu.startswith('n') else (t if tu.startswith('u') else t*1000)
    B=lambda x: d[x][0]*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[d[x][1]] if x in d else 0
    a=agg[re.sub(r'\(.*','',k)]; a[0]+=1; a[1]+=t_us; a[2]+=B('dram__bytes_read.sum')+B('dram__bytes_write.sum')
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda x:-x[1][1]):
    print(f"{a[1]:10.0f} us {100*a[1]/tot:5.1f}% n={a[0]:4d} avg {a[1]/a[0]:8.1f} us  {a[2]/a[1]/1e3:7.1f} GB/s  dram {a[2]/1e9:8.3f} GB  {k[:80]}")
