import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from speaker_embedding_torch_b200 import GE2E
from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
dev = torch.device('cuda', 0)
torch.manual_seed(0)
m = GE2E(default_hyper_parameters()).to(dev).eval()
gen = torch.Generator(device=dev).manual_seed(1)
mel = bench.synth_mel(gen, 960, 160, dev)
with torch.no_grad():
    for _ in range(3): d = m(mel)
torch.cuda.synchronize()
print(float(d.sum()))
