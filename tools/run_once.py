"""One memorize x3 + readout x3 at the DAVIS-17 shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swem_b200 import SWEMCore
from swem_b200.synthetic import em_inputs
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device('cuda:0')
x, v, masks = (t.to(dev) for t in em_inputs(1, N, 64, 512, 30, 54, seed=0))
core = SWEMCore(n_bases=128, valdim=512, n_iters=4, tau=0.05, topl=64).to(dev).eval()
with torch.no_grad():
    for _ in range(4):
        core.memorize(x, v, masks)
    for _ in range(3):
        core.matching_features(x, v[:, 0])
torch.cuda.synchronize()
print('done')
