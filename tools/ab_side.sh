for rep in 1 2; do for m in 15 31; do
  SYMGPU_SIDE=$m python bench.py --steps 30 --warmup 3 --cpu-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('side $m', round(d['ms_per_step'],4), 'wall/step', round(d['wall_s']/30*1000,4), 'e2e', round(d['e2e']['ms_per_step'],4), 'copy', round(d['e2e']['d2h_copy_ms_per_step'],3))"
done; done
