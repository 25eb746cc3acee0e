#!/bin/sh
# usage: run_ab.sh name[@ENV=val]...   (build/ab/name.so, "base" = the in-tree library) -> gpurun_out/ab_<name>.json
for spec in "$@"; do
  n=${spec%%@*}; envs=""; [ "$n" != "$spec" ] && envs=${spec#*@}
  if [ "$n" = base ]; then lib=/root/repo/open-symuvia_b200/libSymuViaB200.so; else lib=/root/repo/build/ab/$n.so; fi
  tag=$(echo "$spec" | tr '@=' '__')
  env SYMB200_LIB=$lib $envs python bench.py --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
    print(n, "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k[2:]:round(v/30,4) for k,v in d["roofline"]["kernel_ms"].items() if v}, "passes", d.get("follow_passes"))
except Exception as e:
    print(n, "FAILED", e)
PY
done
