for mb in 4096 66 132 264; do echo "chunk_mb=$mb $(TROYB_KS_CHUNK_MB=$mb python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sweep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), 'ops/s', ' '.join(k['kernel']+'='+str(round(k['ms'],3)) for k in d['kernels']))")"; done
