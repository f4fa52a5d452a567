import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print('value %.3f %s  ms/step %.3f  e2e %.3f  profiled_step_ms %.3f (T=%s) launches %s' % (
    d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d.get('profiled_step_ms', 0), d.get('profiled_T'), d.get('gpu_launches')))
print('clocks', d.get('clocks'), 'cpu', d.get('cpu_baseline'))
print('extra', d.get('extra'))
print('roofline', d.get('roofline'))
for key in ('breakdown', 'infer_breakdown'):
    if key in d:
        print('---', key)
        tot = 0
        for k, v in d[key].items():
            ms = v.get('ms_per_step', v.get('ms'))
            tot += ms
            print('%-26s %8.3f ms  x%-3d  %7.1f TF/s %8.1f GB/s' % (k, ms, v['launches'], v['tflops'], v['gbs']))
        print('total', round(tot, 3))
