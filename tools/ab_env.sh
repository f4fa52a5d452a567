#!/bin/bash
# A/B an environment switch on the same box: tools/ab_env.sh VAR steps   (runs VAR=1 and VAR=0 alternately)
V=$1; K=${2:-20}
for r in 1 2; do for v in 1 0; do
  env $V=$v timeout 200 python bench.py --steps $K --warmup 3 --no-cpu-baseline > gpurun_out/bench_${V}${v}_$r.json 2> gpurun_out/bench_${V}${v}_$r.err
  python tools/show_bench.py gpurun_out/bench_${V}${v}_$r.json > gpurun_out/bench_${V}${v}_$r.txt
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${V}${v}_$r.json') if l.startswith('{')][-1])
x=d.get('extra',d)
print('$V=$v run $r: value %.2f e2e %.2f prof_ms %.3f dvec %d multislice %d infer_ms %.3f' % (d['value'], d['e2e']['value'], d['profiled_step_ms'], x['dvectors_per_sec_160f_1gpu'], x['multislice_utt_per_sec_5x64_1gpu'], x['infer_ms_per_960x160_batch']))
PY
done; done
