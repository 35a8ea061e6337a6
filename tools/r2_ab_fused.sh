for f in 0 1; do
  TROYB_KS_FUSED=$f python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/ab_fused_$f.json 2> gpurun_out/ab_fused_$f.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_fused_$f.json'))
print('fused=$f', round(d['value']), round(d['ms_per_step'],3), 'rotate', round(d['rotate']['value']), [(k['kernel'], round(k['ms'],3)) for k in d['kernels']])
PY
done
