#!/bin/bash
# usage (on the GPU box): tools/bench_matrix.sh N "workload ..."  -> gpurun_out/r02_bench_n{N}[_{workload}].json
N=$1; shift
for wl in $1; do
  suffix=""; [ "$wl" != "dragon4k" ] && suffix="_${wl}"
  out=gpurun_out/r02_bench_n${N}${suffix}.json
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --workload $wl > $out 2> ${out%.json}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 --workload $wl > $out 2> ${out%.json}.err
  fi
  echo "N=$N $wl rc=$?"
  python - "$out" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print({k: d.get(k) for k in ("value", "ms_per_step", "frame_matches_oracle", "gpu_launches")}, "e2e", d["e2e"]["ms_per_frame"], "pipelined", d["e2e"]["pipelined"]["ms_per_frame"], d.get("one_time_ms"))
PY
done
