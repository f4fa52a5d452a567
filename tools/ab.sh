#!/bin/bash
# A/B two builds of libspkemb.so on the same box: tools/ab.sh <steps>   (expects libspkemb_new.so / libspkemb_ab.so)
K=${1:-20}
for r in 1 2; do for v in new ab; do
  cp speaker_embedding_torch_b200/libspkemb_$v.so speaker_embedding_torch_b200/libspkemb.so
  timeout 200 python bench.py --steps $K --warmup 3 --no-cpu-baseline > gpurun_out/bench_${v}_$r.json 2> gpurun_out/bench_${v}_$r.err
  python tools/show_bench.py gpurun_out/bench_${v}_$r.json > gpurun_out/bench_${v}_$r.txt
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${v}_$r.json') if l.startswith('{')][-1])
x=d.get('extra',d)
print('${v}_$r value %.2f e2e %.2f prof_ms %.3f dvec %d multislice %d infer_ms %.3f' % (d['value'], d['e2e']['value'], d['profiled_step_ms'], x['dvectors_per_sec_160f_1gpu'], x['multislice_utt_per_sec_5x64_1gpu'], x['infer_ms_per_960x160_batch']))
PY
done; done
cp speaker_embedding_torch_b200/libspkemb_new.so speaker_embedding_torch_b200/libspkemb.so
