for r in 0 32; do
USL_COL_R=$r python bench.py --no-cpu --steps 100 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
e=d['extra']
print('R=$r c2',round(d['ms_per_step'],4),'c3',round(e['c3_strong']['ms_per_step'],4),'c4',round(e['c4_adversarial']['plain']['ms_per_step'],4),'c4adv',round(e['c4_adversarial']['adversarial']['ms_per_step'],4))
"
for w in c1; do
USL_COL_R=$r python bench.py --no-cpu --no-extra --workload $w --steps 100 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('R=$r $w',round(d['ms_per_step'],4))
"
done
done
