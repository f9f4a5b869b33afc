# A/B on the GPU box: tools/ab_env.sh <out dir> <ENV_NAME> "<bench args>" <values...>  - device-timed legs only, per value, twice
OUT=gpurun_out/$1; VAR=$2; ARGS=$3; shift 3
mkdir -p $OUT
B="python bench.py --no-e2e --no-cpu-baseline --no-extras --no-config5 $ARGS"
for i in 1 2; do for v in "$@"; do env $VAR=$v $B > $OUT/bench_${VAR}_${v}_$i.json 2>> $OUT/err.txt; done; done
python - $OUT <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+"/bench_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"],4), {k[:18]:round(v,3) for k,v in d["kernels_ms_per_step"].items() if isinstance(v,float)})
PY
