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
    assert lib.swem_abi_version() == 3


def test_struct_layouts_match_header(lib):
    from swem_b200 import _lib
    assert C.sizeof(_lib.SwemDims) == 40
    assert C.sizeof(_lib.SwemEmArgs) == 40 + 10 * 8 + 8 + 8 + 8 + 8 + 8      # dims, 10 pointers, ws ptr, size, path + layout, image ws, image bank + banks
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


def test_fused_family_coverage_predicates(lib):
    """What SWEM_PATH_AUTO runs = the tcgen05 family: Ck in {64, 128}, L in {64, 128, 256, 512}, Cv = 512 at any frame
    size (the EM kernel goes windowed when a unit's clusters cannot be co-resident).  Anything else is refused under AUTO
    (workspace query 0, forward SWEM_ERR_UNSUPPORTED -- SURVEY 8(b): fail loudly); the generic fp32 family runs only when
    SWEM_PATH_GENERIC is asked for by name."""
    from swem_b200 import _lib
    dims = lambda ck, cv, hw, l, banks=2: _lib.SwemDims(1, 5, ck, cv, hw, l, 4, banks, min(l, 64), 0.05)
    for ck in (64, 128):
        for l in (64, 128, 256, 512):
            for hw in (240, 1620, 6480):
                assert lib.swem_em_fused_supported(C.byref(dims(ck, 512, hw, l))) == 1, (ck, l, hw)
                for banks in (1, 2):
                    assert lib.swem_readout_fused_supported(C.byref(dims(ck, 512, hw, l, banks))) == 1, (ck, l, hw, banks)
    for bad in (dims(32, 512, 1620, 128), dims(64, 256, 1620, 128), dims(64, 512, 1620, 384), dims(64, 512, 1620, 1024)):
        assert lib.swem_em_fused_supported(C.byref(bad)) == 0
        assert lib.swem_readout_fused_supported(C.byref(bad)) == 0
        assert lib.swem_em_workspace_bytes(C.byref(bad), _lib.PATH_GENERIC) > 0
        assert lib.swem_em_workspace_bytes(C.byref(bad), _lib.PATH_AUTO) == 0
        assert lib.swem_readout_workspace_bytes(C.byref(bad), _lib.PATH_AUTO) == 0
        assert lib.swem_readout_workspace_bytes(C.byref(bad), _lib.PATH_GENERIC) > 0


def test_pixel_major_values_are_only_taken_when_the_fused_kernels_run():
    from swem_b200 import SWEMCore
    core = SWEMCore(n_bases=128, valdim=512, n_iters=4, tau=0.05, topl=64)
    v = torch.randn(3, 512, 6, 8).contiguous(memory_format=torch.channels_last).view(1, 3, 512, 6, 8)
    assert not core._takes_pixel_major(v, 1, 3, 64, 48)               # CPU tensor: never (and swem() raises on it anyway)
    with pytest.raises(RuntimeError):
        core.swem(torch.randn(1, 64, 6, 8), v, torch.rand(1, 3, 2, 6, 8))


def test_multiscale_flip_evaluation_with_a_stub_model():
    """evaluate_davis_seq_ms (swem_evaluator.py:34-57) on a stub model whose scores are a fixed function of the frame:
    scores are averaged over the flip pair (un-flipped first) and then over the scales, argmax last."""
    import torch.nn.functional as F
    from swem_b200.evaluator import evaluate_davis_seq, evaluate_davis_seq_ms

    class Stub:
        def __call__(self, mode, *a):
            if mode == 'encode_key':
                f = a[0]
                return f, f, f, f, f
            if mode == 'match':
                return a[0], 2
            if mode == 'segment':
                n, ctx, out_size = a[0], a[1], a[5]
                g = F.interpolate(ctx, size=out_size, mode='bilinear', align_corners=False)
                ramp = torch.linspace(0, 1, out_size[1]).view(1, 1, 1, -1)       # not flip-symmetric
                prob = torch.softmax(torch.cat([g[:, :1] * 0 + 0.3, g[:, :2] * (1 + ramp)], 1) * 5, dim=1)
                return None, prob
            return a[0] if a else None                                           # encode_value / init / memorize

    frames = torch.rand(1, 4, 3, 48, 80, generator=torch.Generator().manual_seed(0))
    init = [torch.zeros(1, 3, 48, 80)] + [None] * 3
    out = (48, 80)
    scales = (240, 480)
    want = [0] * 3
    for s_ in scales:
        fr = F.interpolate(frames[0], size=(s_, int(s_ / 480 * 864)), mode='bicubic', align_corners=False)[None]
        _, a = evaluate_davis_seq(Stub(), fr, init, out)
        _, b = evaluate_davis_seq(Stub(), torch.flip(fr, dims=[-1]), [torch.flip(init[0], dims=[-1])], out)
        want = [acc + (x + torch.flip(y, dims=[-1])) / 2 / len(scales) for acc, x, y in zip(want, a, b)]
    got = evaluate_davis_seq_ms(Stub(), frames, init, out, scales=scales, is_flip=True)
    assert len(got) == 3 and all(torch.equal(g, torch.argmax(w, dim=1)) for g, w in zip(got, want))
    plain, _ = evaluate_davis_seq(Stub(), F.interpolate(frames[0], size=(480, 864), mode='bicubic', align_corners=False)[None], init, out)
    assert all(torch.equal(a, b) for a, b in zip(plain, evaluate_davis_seq_ms(Stub(), frames, init, out)))
