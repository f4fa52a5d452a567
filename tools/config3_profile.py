import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from speaker_embedding_torch_b200 import GE2E, _native as N
from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
dev = torch.device('cuda', 0)
torch.manual_seed(0)
m = GE2E(default_hyper_parameters()).to(dev).eval()
gen = torch.Generator(device=dev).manual_seed(1)
w = bench.synth_mel(gen, 4000, 192, dev).half()
with torch.no_grad():
    for _ in range(3): m.embed_windows(w, 64, 32)
    torch.cuda.synchronize()
    N.prof_enable(True)
    for _ in range(2): m.embed_windows(w, 64, 32)
    torch.cuda.synchronize()
rep = N.prof_report(); N.prof_enable(False)
tot = sum(v['ms'] for v in rep.values()) / 2
print('config3 chunk: %.3f ms' % tot)
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]['ms'])[:16]:
    print('  %-24s %.4f ms x%d' % (k, v['ms'] / 2, v['launches'] // 2))
