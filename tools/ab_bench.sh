# A/B of shared-library variants on one box: sh tools/ab_bench.sh variant...   (open-symuvia_b200/libSymuViaB200_<variant>.so; "base" = the default library)
for rep in 1 2; do for v in "$@"; do
  if [ "$v" = base ]; then unset SYMB200_LIB; else export SYMB200_LIB=/root/repo/open-symuvia_b200/libSymuViaB200_$v.so; fi
  python bench.py --steps 30 --warmup 3 --cpu-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms']; print('$v', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'merge(bd)', k['k_merge'], 'ctx', k['k_context'], 'relax', k['k_relax'])"
done; done
