#!/bin/bash
# A/B two builds of libspkemb.so on the same box (alternating, three rounds):
#   build variant A -> cp libspkemb.so libspkemb_new.so ; build variant B -> cp libspkemb.so libspkemb_ab.so
#   gpurun -- 'bash tools/ab.sh'
for r in 1 2 3; do for v in new ab; do
  cp speaker_embedding_torch_b200/libspkemb_$v.so speaker_embedding_torch_b200/libspkemb.so
  timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_${v}_$r.json 2> gpurun_out/ab_${v}_$r.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/ab_${v}_$r.json') if l.startswith('{')][-1])
print('${v}_$r value %.2f e2e %.2f prof_ms %.3f clocks %s %s dvec %.0f' % (d['value'], d['e2e']['value'], d['profiled_step_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['infer']['dvectors_per_sec_160f']))
PY
done; done
cp speaker_embedding_torch_b200/libspkemb_new.so speaker_embedding_torch_b200/libspkemb.so
