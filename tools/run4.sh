python -m pytest tests -m gpu -x -q -k "stft or mel or multi_device" 2>&1 | tail -6
python bench.py --config c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4 gen2', d['config']['kernel'], round(d['ms_per_step'],3), round(d['roofline']['frac'],4))"
SS_STFT400_GEN=1 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4 gen1', d['config']['kernel'], round(d['ms_per_step'],3), round(d['roofline']['frac'],4))"
