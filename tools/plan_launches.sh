#!/bin/bash
# Per-kernel device time of one plan() call (the 4th of tools/plan_timing.py) from an ncu launch list.  Times under ncu are
# serialised and cold-cache: read them as shares, not as bench values.
#   tools/plan_launches.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
for s in "3D_superquadrics_cluttered_rabbit 0 10" "3D_superquadrics_maze_rabbit 0 10 44 20"; do
  set -- $s
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_plan_launches_$1.csv python tools/plan_timing.py $s > /dev/null 2>&1
  python - "$1" "$TAG" <<PY
import csv,sys,collections
rows=[r for r in csv.reader(open("gpurun_out/%s_plan_launches_%s.csv"%(sys.argv[2],sys.argv[1]))) if len(r)>5 and r[0].isdigit()]
n=len(rows)//4
agg=collections.OrderedDict()
for r in rows[-n:]:
    k=r[4][:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[-1])/1e3
print(sys.argv[1], "launches per call", n, "sum %.1f us" % sum(t for _,t in agg.values()))
for k,(c,t) in agg.items(): print("  %-62s x%-3d %9.1f us"%(k,c,t))
PY
done
