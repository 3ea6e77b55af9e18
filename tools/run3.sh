python -m pytest tests -m gpu -x -q -k "stft or mel" 2>&1 | tail -6
for g in 2 1; do SS_STFT2048_GEN=$g python bench.py --config c3 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c3 gen$g', d['config']['kernel'], round(d['ms_per_step'],3), round(d['roofline']['frac'],4))"; done
