"""FrameEngine (swem_b200/engine.py) evaluates the same function as the plain torch modules: BN folding,
stacked heads, linearity splits of the object-independent conv inputs.  CPU here; the CUDA path (fused cuDNN
ops, readout into the narrow buffer, whole sequences) is in test_gpu_parity.py."""
import pytest
import torch

from swem_b200 import SWEM, make_config
from swem_b200.engine import FrameEngine, _fold_bn


def _randomise_bn(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def test_fold_bn_matches_conv_then_bn():
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(5, 7, 3, stride=2, padding=1, bias=True)
    bn = torch.nn.BatchNorm2d(7).eval()
    bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2); bn.weight.data.uniform_(0.5, 2); bn.bias.data.normal_()
    x = torch.randn(2, 5, 9, 11)
    with torch.no_grad():
        w, b = _fold_bn(conv, bn)
        assert _rel(torch.nn.functional.conv2d(x, w, b, stride=2, padding=1), bn(conv(x))) < 1e-5


@pytest.mark.parametrize('backbone,single', [('resnet50', False), ('resnet18', True)])
def test_engine_matches_modules_on_cpu(backbone, single):
    torch.manual_seed(0)
    model = SWEM(make_config(keydim=64, n_bases=16, n_iters=2, topl=8, backbone=backbone, single_obj=single)).eval()
    _randomise_bn(model)
    eng = FrameEngine(model, channels_last=False)
    n = 1 if single else 3
    h, w = 64, 96
    frame = torch.rand(1, 3, h, w)
    masks = torch.softmax(torch.randn(1, n + 1, h, w) * 3, dim=1)
    with torch.no_grad():
        want = model('encode_key', frame)
        got = eng('encode_key', frame)
        for a, b, name in zip(got, want, ('qk16', 'qv16', 'f16', 'f8', 'f4')):
            assert a.shape == b.shape and _rel(a, b) < 2e-5, name
        qk16, qv16, s16, s8, s4 = want
        mv_want = model('encode_value', frame, masks, s16)
        mv_got = eng('encode_value', frame, masks, s16)
        assert mv_got.shape == mv_want.shape and _rel(mv_got, mv_want) < 2e-5
        ctx = torch.randn(n, 512, h // 16, w // 16)
        lg_want, pr_want = model('segment', n, ctx, s8, s4, None, (h, w))
        lg_got, pr_got = eng('segment', n, ctx, s8, s4, None, (h, w))
        assert _rel(lg_got, lg_want) < 2e-5 and _rel(pr_got, pr_want) < 2e-5
        valid = torch.tensor([[1.0] + [1.0] * (n - 1) + [0.0]])
        lg_want, _ = model('segment', n, ctx, s8, s4, valid, (h, w))
        lg_got, _ = eng('segment', n, ctx, s8, s4, valid, (h, w))
        assert _rel(lg_got, lg_want) < 2e-5


def test_engine_is_inference_only():
    model = SWEM(make_config(keydim=64, n_bases=16, n_iters=1, topl=8, backbone='resnet18'))
    with pytest.raises(RuntimeError):
        FrameEngine(model.train()).refresh()
    eng = FrameEngine(model.eval(), channels_last=False)
    with pytest.raises(RuntimeError):
        eng('encode_key', torch.rand(1, 3, 32, 32))        # grad mode on


def test_tf32_split_conv_identity():
    """The algebra behind FrameEngine(split_tf32=True): with x = xh + xl, w = wh + wl (hi parts on the TF32 grid),
    conv(xh, wh) + conv([xh | xl], [wl ; wh]) = conv(x, w) - conv(xl, wl), i.e. fp32-accurate; the hi parts have 13 zero
    low mantissa bits (a TF32 tensor core takes them unchanged) and hi + lo reproduces the input exactly."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 9, 11, generator=g) * 3
    w = torch.randn(6, 8, 3, 3, generator=g) * 0.2
    xh, xl = FrameEngine._tf32_split(x)
    wh, wl = FrameEngine._tf32_split(w)
    assert torch.equal(xh + xl, x) and torch.equal(wh + wl, w)
    assert int((xh.view(torch.int32) & 0x1fff).abs().max()) == 0 and int((wh.view(torch.int32) & 0x1fff).abs().max()) == 0
    assert (xl.abs() <= x.abs() * 2.0 ** -11 + 1e-30).all()
    conv = torch.nn.functional.conv2d
    want = conv(x.double(), w.double(), padding=1)
    got = conv(xh.double(), wh.double(), padding=1) + conv(torch.cat([xh, xl], 1).double(), torch.cat([wl, wh], 1).double(), padding=1)
    assert _rel(got, want) < 2.0 ** -20
    assert _rel(conv(xh.double(), wh.double(), padding=1), want) > 2.0 ** -14      # a single TF32 conv is 1e-4 .. 1e-3 off


def test_fusion_kernel_is_cuda_only_and_opt_in_by_accuracy_mode():
    """FrameEngine uses swem_fusion_conv_glu by default exactly when fp32-accurate convolutions are asked for (split_tf32) and never
    builds its weight images for a CPU model (no CUDA: the layer runs through torch as everywhere else on CPU)."""
    torch.manual_seed(0)
    model = SWEM(make_config(keydim=64, n_bases=16, n_iters=1, topl=8, backbone='resnet18')).eval()
    assert FrameEngine(model).fusion_kernel is False
    assert FrameEngine(model, split_tf32=True).fusion_kernel is True
    assert FrameEngine(model, split_tf32=True, fusion_kernel=False).fusion_kernel is False
    eng = FrameEngine(model, split_tf32=True)
    eng.refresh()
    assert eng.g_fused is None


def test_get_affinity_refuses_cpu_tensors():
    """SWEMCore.get_affinity (the reference's signature, incl. the kernelised-memory branch) has no CPU fallback."""
    from swem_b200 import SWEMCore
    core = SWEMCore(n_bases=8, valdim=16, n_iters=1, tau=0.05, topl=4).eval()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        core.get_affinity(torch.randn(1, 16, 4, 5), torch.randn(1, 2, 2, 16, 16), torch.randn(1, 2, 2, 16, 16), n_kernel=3)
