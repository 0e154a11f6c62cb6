#!/bin/bash
# Strip-height sweep of the column kernels on config 2:
# USL_COL_R0 = rows per strip of the largest scale, USL_COL_R = of the others
# (0 = the planner's own choice).  Run on a GPU box: tools/grun.sh 1200 -- 'bash tools/sweep_strips.sh'
for cfg in "0 0" "0 32" "0 24" "0 44" "0 64" "0 96" "64 0" "128 0"; do
set -- $cfg
USL_COL_R0=$1 USL_COL_R=$2 python bench.py --no-extra --no-cpu --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
k=d['kernels_ms']
print('R0=$1 R=$2 step',round(d['ms_per_step'],4),'fused',round(k['loss_fused_main']*1e3,1),'fused+scatter',round(k['loss_fused_plus_scatter']*1e3,1))
"
done
