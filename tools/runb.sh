python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --config c2 --steps 30 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_c2_g3.json 2>/dev/null
python bench.py --config c5 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_c5_g3.json 2>/dev/null
python -c "
import json
for f in ['c2_g3','c5_g3']:
    d=json.load(open('gpurun_out/b_%s.json'%f)); print(f, round(d['ms_per_step'],3), round(d['roofline']['frac'],4))
"
ncu --set full --import-source on --clock-control none -k regex:k_fast512c -c 1 -f -o gpurun_out/prof_fast512c_r1c python bench.py --config c2 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_c.log 2>&1
