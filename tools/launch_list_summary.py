"""Share of each kernel in an ncu launch list:  python tools/launch_list_summary.py launches.csv
(the CSV of `ncu --metrics gpu__time_duration.sum --csv --log-file launches.csv <command>`)."""
import csv,sys
from collections import defaultdict
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>14 and r[0].isdigit()]
t=defaultdict(lambda:[0,0.0])
for r in rows:
    k=r[4].split("(")[0][-40:]; t[k][0]+=1; t[k][1]+=float(r[14])/1e6
for k,v in sorted(t.items(), key=lambda kv:-kv[1][1])[:6]: print(v[0], round(v[1],2), k)
