"""The one-line substitution of INTEGRATION.md, executed against the real reference checkout when it
is present (authoring container); skipped on the GPU box where /root/reference does not exist."""
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason='reference checkout not present')


def test_reference_model_builds_on_our_core_and_keeps_its_checkpoint_format():
    import swem_b200
    RefSWEM, ref_modules = ref_shim.load_swem()
    import methods.SWEM.swem as ref_swem
    cfg = ref_shim.model_config(keydim=64, n_bases=128)
    torch.manual_seed(0)
    stock = RefSWEM(cfg)
    original = ref_swem.SWEMCore
    ref_swem.SWEMCore = swem_b200.SWEMCore                     # <- the substitution
    try:
        model = RefSWEM(cfg).eval()
    finally:
        ref_swem.SWEMCore = original
    assert isinstance(model.swem_core, swem_b200.SWEMCore)
    assert list(model.state_dict()) == list(stock.state_dict())
    model.load_state_dict(stock.state_dict())                  # a reference checkpoint loads unchanged
    # the reference's own call sequence reaches our core; on CPU it must refuse, not compute
    qk, qv = torch.randn(1, 64, 4, 6), torch.randn(1, 512, 4, 6)
    with pytest.raises(RuntimeError, match='memory is empty'):
        model('match', qk, qv)
    mask = torch.zeros(1, 2, 64, 96); mask[:, 0] = 1
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model('init', qk, torch.randn(1, 1, 512, 4, 6), mask)
