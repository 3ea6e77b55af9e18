# latency case c1 under the A/B switches (same box, back to back)
for v in "" "SS_F512C_STAGE=0" "SS_FAST512_GEN=2"; do
  for i in 1 2; do
    env $v python bench.py --config c1 --steps 200 --warmup 20 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c1 [$v]', round(d['ms_per_step']*1000,2), 'us')"
  done
done
