"""One memorize x4 + readout x3 + GLU fusion convolution x2 at the DAVIS-17 shape (for ncu captures and compute-sanitizer)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swem_b200 import SWEMCore, _lib
from swem_b200.synthetic import em_inputs
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device('cuda:0')
H, W, Cv, TL = 30, 54, 512, 64
x, v, masks = (t.to(dev) for t in em_inputs(1, N, 64, Cv, H, W, seed=0))
core = SWEMCore(n_bases=128, valdim=Cv, n_iters=4, tau=0.05, topl=TL).to(dev).eval()
lib = _lib.load()
with torch.no_grad():
    for _ in range(4):
        core.memorize(x, v, masks)
    for _ in range(3):
        core.matching_features(x, v[:, 0])
    # the engine's layout: [mem_out | S] channels-last, then the fusion layer on the per-object channels (engine.py: match)
    feats = torch.empty(N, Cv + 2 * TL, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    core.readout_into(x, feats, 0, Cv)
    fl = core.fusion_layer
    w = torch.cat([fl.layer_f.weight, fl.layer_a.weight], 0)
    w_obj = torch.cat([w[:, :Cv], w[:, 2 * Cv:]], 1).detach().float().contiguous()
    scale = 2.0 ** (11 - math.ceil(math.log2(float(w_obj.abs().max()))))
    cin, cout = Cv + 2 * TL, fl.layer_f.out_channels
    wblob = torch.empty(lib.swem_fusion_weight_bytes(cin, cout), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.swem_fusion_prepare_weights(w_obj.data_ptr(), cin, cout, scale, wblob.data_ptr(), st), 'prepare')
    ws = torch.empty(lib.swem_fusion_workspace_bytes(N, H, W, cin), dtype=torch.uint8, device=dev)
    shared = torch.randn(1, H, W, 2 * cout, device=dev)
    bias = torch.zeros(2 * cout, device=dev)
    out = torch.empty(N, H, W, cout, device=dev)
    for _ in range(2):
        _lib.check(lib.swem_fusion_conv_glu(feats.data_ptr(), wblob.data_ptr(), scale, shared.data_ptr(), bias.data_ptr(), N, N, H, W, cin, cout,
                                            ws.data_ptr(), ws.numel(), out.data_ptr(), st), 'fusion')
torch.cuda.synchronize()
assert torch.isfinite(out).all()
print('done')
