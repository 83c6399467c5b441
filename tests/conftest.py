import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    import torch

    def load(name):
        return torch.load(os.path.join(GOLDEN, name + '.pt'), weights_only=False)
    return load


@pytest.fixture(scope='session', autouse=True)
def _gpu_warmup():
    """One small call of every kernel family before the first GPU test: context creation, lazy module loading of the
    library's kernels, cuDNN/cuBLAS handles and the caching allocator's first blocks happen here, not inside a parity
    check (nothing is asserted on these results)."""
    import torch
    if not torch.cuda.is_available():
        yield
        return
    from swem_b200 import SWEMCore, _lib
    from swem_b200.synthetic import em_inputs
    dev = torch.device('cuda:0')
    with torch.no_grad():
        for (ck, cv, L, h, w, path) in ((16, 24, 8, 5, 7, _lib.PATH_GENERIC), (64, 512, 128, 12, 20, _lib.PATH_AUTO)):
            core = SWEMCore(n_bases=L, valdim=cv, n_iters=2, tau=0.05, topl=min(L, 64)).to(dev).eval()
            core.em_path = core.readout_path = path
            x, v, masks = (t.to(dev) for t in em_inputs(1, 2, ck, cv, h, w, seed=0))
            core.memorize(x, v, masks)
            core.memorize(x, v, masks)
            core.matching_features(x, v[:, 0])
    torch.cuda.synchronize()
    yield
