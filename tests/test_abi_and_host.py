"""CPU-side checks: the C-ABI library loads and exports what include/swem_b200.h declares, argument
validation returns the documented status codes (no kernel is launched), and the host-side state
machine (memory banks, mode dispatch, error behaviour) mirrors the reference's."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from swem_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    from swem_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'swem_b200.h')).read()
    declared = set(re.findall(r'^\s*(?:int|size_t|long long|const char\*)\s+(swem_\w+)\s*\(', header, flags=re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.swem_abi_version() == 2


def test_struct_layouts_match_header(lib):
    from swem_b200 import _lib
    assert C.sizeof(_lib.SwemDims) == 40
    assert C.sizeof(_lib.SwemEmArgs) == 40 + 10 * 8 + 8 + 8 + 8      # dims, 10 pointers, ws ptr, size, path(+pad)
    assert _lib.SwemReadArgs.out.offset == 40 + 8 + 16 + 16


def test_argument_validation_without_gpu(lib):
    from swem_b200 import _lib
    assert lib.swem_em_forward(None, None) == 1
    assert b'NULL' in lib.swem_last_error()
    d = _lib.SwemDims(1, 1, 64, 512, 100, 128, 4, 0, 0, 0.0)          # tau = 0
    args = _lib.SwemEmArgs(d)
    assert lib.swem_em_forward(C.byref(args), None) == 1
    assert b'tau' in lib.swem_last_error()
    d = _lib.SwemDims(1, 1, 64, 512, 100, 128, 4, 0, 0, 0.05)
    args = _lib.SwemEmArgs(d)                                          # null pointers
    assert lib.swem_em_forward(C.byref(args), None) == 1
    assert lib.swem_em_workspace_bytes(C.byref(d), _lib.PATH_GENERIC) > 0
    r = _lib.SwemDims(1, 1, 64, 512, 100, 128, 0, 3, 64, 0.05)        # n_banks = 3
    assert lib.swem_readout_workspace_bytes(C.byref(r), _lib.PATH_GENERIC) == 0
    ra = _lib.SwemReadArgs(r)
    assert lib.swem_readout_forward(C.byref(ra), None) == 1
    # forcing the fused family on a shape it does not cover is an error, never a silent switch
    odd = _lib.SwemDims(1, 1, 24, 40, 35, 8, 3, 0, 0, 0.05)
    assert lib.swem_em_fused_supported(C.byref(odd)) == 0
    assert lib.swem_em_workspace_bytes(C.byref(odd), _lib.PATH_FUSED) == 0


def _fake_bases(n, tag):
    return {'kappa': torch.full((1, n, 2, 4, 3), float(tag)), 'nu': torch.full((1, n, 2, 5, 3), float(tag)),
            'zita': torch.full((1, n, 2, 1, 3), float(tag))}


def test_memory_bank_state_machine_matches_oracle():
    from oracle import swem_oracle as O
    from swem_b200 import SWEMCore
    core = SWEMCore(n_bases=3, valdim=5, n_iters=1, tau=0.05, topl=2)
    ref = O.MemoryBanks()

    def commit(b):                     # the bookkeeping half of SWEMCore.memorize (modules.py:189-193)
        if core.memories['first'].bases is None:
            core.memories['first'].update(b)
        else:
            core.memories['first'].update(b)
            core.memories['update'].update(b)
        ref.commit(b)

    for n, tag in ((2, 1), (2, 2), (3, 3), (3, 4)):
        commit(_fake_bases(n, tag))
        k, v = core.get_mem()
        rk, rv = ref.read()
        assert torch.equal(k, rk) and torch.equal(v, rv)
        prior = core.memories['update'].bases or core.memories['first'].bases
        assert prior is ref.prior()
    assert core.memories['first'].n_objs == 3 and core.get_mem()[0].shape[-1] == 6
    # the late object's 'first' entry is the bases of the call that introduced it
    assert core.memories['first'].bases['kappa'][0, 2].unique().item() == 3.0
    core.empty()
    assert core.memories['first'].bases is None and core.memories['update'].bases is None


def test_reference_surface_is_present():
    from swem_b200 import SWEM, SWEMCore, make_config
    core = SWEMCore(n_bases=32, valdim=64, n_iters=2, tau=0.1, topl=64)
    assert core.topl == 32 and core.p_drop == 0.0 and core.n_bases == 32 and core.n_iters == 2
    assert set(core.state_dict()) == {f'fusion_layer.layer_{k}.{p}' for k in 'fa' for p in ('weight', 'bias')}
    assert core.fusion_layer.layer_f.in_channels == 2 * 64 + 2 * 32
    for name in ('empty', 'memorize', 'matching', 'get_mem', 'swem', 'random_init'):
        assert callable(getattr(core, name))
    with pytest.raises(AssertionError):
        SWEMCore(tau=0.0)
    model = SWEM(make_config(backbone='resnet18', n_bases=16))
    with pytest.raises(NotImplementedError):
        model('no_such_mode')
    with pytest.raises(KeyError):
        SWEM(make_config(backbone='resnet101'))


def test_cpu_tensors_are_rejected_not_computed():
    from swem_b200 import SWEMCore
    core = SWEMCore(n_bases=8, valdim=16, n_iters=1, tau=0.05, topl=4)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        core.memorize(torch.randn(1, 8, 3, 3), torch.randn(1, 1, 16, 3, 3), torch.rand(1, 1, 2, 3, 3))


def test_product_never_imports_the_oracle():
    import subprocess
    import sys
    code = ('import sys; import swem_b200, swem_b200.evaluator, swem_b200.synthetic; '
            'bad=[m for m in sys.modules if m.split(".")[0]=="oracle"]; assert not bad, bad')
    subprocess.run([sys.executable, '-c', code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'swem_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                assert 'oracle' not in open(os.path.join(dirpath, f)).read().replace('use oracle/ for CPU checking', '')
