"""Parity of the CUDA hot path (through the C ABI) with the CPU oracle and the reference goldens.

Tolerances (north_star): readout features max-rel error <= 1e-2; masks >= 99.9 % argmax agreement.
max-rel error here = max|a-b| / max|b| per tensor; kappa / nu are compared on live bases
(zita > 1e-3) because dead bases are ill-conditioned (hard part 4).
Both kernel families are exercised: GENERIC on every shape, FUSED on the shapes it covers.
"""
import ctypes as C

import pytest
import torch

from oracle import swem_oracle as O

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'
TOL = dict(generic=dict(bases=2e-4, feat=2e-4), fused=dict(bases=1e-2, feat=1e-2))


def maxrel(a, b, mask=None):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    if mask is not None:
        mask = mask.expand_as(b)
        if not mask.any():
            return 0.0
        a, b = a[mask], b[mask]
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _core(cfg, family):
    from swem_b200 import SWEMCore, _lib
    core = SWEMCore(n_bases=cfg['L'], valdim=cfg['Cv'], n_iters=cfg['n_iters'], tau=cfg['tau'], topl=cfg['topl'])
    core.kernel_path = _lib.PATH_GENERIC if family == 'generic' else _lib.PATH_FUSED
    return core.to(DEV).eval()


def _fused_covers(B, N, Ck, Cv, HW, L, n_iters=4, n_banks=2, topl=64, tau=0.05, what='em'):
    from swem_b200 import _lib
    lib = _lib.load()
    d = _lib.SwemDims(B, N, Ck, Cv, HW, L, n_iters, n_banks, topl, tau)
    fn = lib.swem_em_fused_supported if what == 'em' else lib.swem_readout_fused_supported
    return bool(fn(C.byref(d)))


def _skip_unless_covered(family, **kw):
    if family == 'fused' and not _fused_covers(**kw):
        pytest.skip('shape not covered by the fused kernels')


def _to(bases, dev):
    return None if bases is None else {k: v.to(dev) for k, v in bases.items()}


def _cpu_random_init(core):
    """random_init that draws on the CPU (seedable identically to the oracle) then moves to the GPU."""
    def init(size, norm_dim=-2, dtype=None, device=None):
        B, N, _, Ck, L = size
        k, n, z = O.random_init(B, N, Ck, L, core.valdim)
        return k.to(device), n.to(device), z.to(device)
    return init


@pytest.mark.parametrize('family', ['generic', 'fused'])
@pytest.mark.parametrize('name', ['core_small', 'core_prod'])
def test_golden_sequences_teacher_forced(golden, name, family):
    """Each memorize call starts from the REFERENCE's previous bases (teacher forcing), readout too."""
    fx = golden(name)
    cfg = fx['cfg']
    HW = cfg['H'] * cfg['W']
    _skip_unless_covered(family, B=cfg['B'], N=fx['calls'][0]['masks'].shape[1], Ck=cfg['Ck'], Cv=cfg['Cv'], HW=HW,
                         L=cfg['L'], n_iters=cfg['n_iters'], topl=min(cfg['L'], cfg['topl']))
    core = _core(cfg, family)
    core.random_init = _cpu_random_init(core)
    ref = O.OracleSWEMCore(n_bases=cfg['L'], valdim=cfg['Cv'], n_iters=cfg['n_iters'], tau=cfg['tau'], topl=cfg['topl'])
    tol = TOL[family]
    with torch.no_grad():
        for call in fx['calls']:
            prior = ref.banks.prior()
            torch.manual_seed(call['rng_seed'])
            got = core.swem(call['x'].to(DEV), call['v'].to(DEV), call['masks'].to(DEV), _to(prior, DEV))
            torch.manual_seed(call['rng_seed'])
            ref.memorize(call['x'], call['v'], call['masks'])
            live = call['zita'] > 1e-3
            assert maxrel(got['zita'], call['zita']) < tol['bases']
            assert maxrel(got['kappa'], call['kappa'], live) < tol['bases']
            assert maxrel(got['nu'], call['nu'], live) < tol['bases']
            assert torch.isfinite(got['kappa']).all() and torch.isfinite(got['nu']).all()
            # readout from the reference's memory
            core.memories['first'].bases = _to(ref.banks.first, DEV)
            core.memories['first'].n_objs = ref.banks.first_n
            core.memories['update'].bases = _to(ref.banks.update, DEV)
            qv = torch.zeros(cfg['B'], cfg['Cv'], cfg['H'], cfg['W'])
            feats, n = core.matching_features(call['q'].to(DEV), qv.to(DEV))
            Cv, tl = cfg['Cv'], core.topl
            mem_out = feats[:, :Cv].reshape(call['mem_out'].shape)
            S = feats[:, 2 * Cv:]
            assert n == call['masks'].shape[1]
            assert maxrel(mem_out, call['mem_out']) < tol['feat']
            assert maxrel(S, call['S']) < tol['feat']
            assert torch.equal(feats[:, Cv:2 * Cv].cpu(), qv.repeat_interleave(n, 0))


@pytest.mark.parametrize('family', ['generic', 'fused'])
def test_last_responsibilities(golden, family):
    fx = golden('steps_small')
    cfg = dict(fx['cfg'], topl=4)
    _skip_unless_covered(family, B=cfg['B'], N=cfg['N'], Ck=cfg['Ck'], Cv=cfg['Cv'], HW=cfg['H'] * cfg['W'], L=cfg['L'],
                         n_iters=cfg['n_iters'], topl=4)
    core = _core(cfg, family)
    prior = {k: fx[k + '_prior'].to(DEV) for k in ('kappa', 'nu', 'zita')}
    with torch.no_grad():
        got = core.swem(fx['x'].to(DEV), fx['v'].to(DEV), fx['masks'].to(DEV), prior, return_z=True)
    assert maxrel(got['z'].view_as(fx['z'][-1]), fx['z'][-1]) < TOL[family]['feat']
    assert maxrel(got['kappa'], fx['kappa'], fx['zita'] > 1e-3) < TOL[family]['bases']


SHAPES = [
    # B, N, Ck,  Cv,  L,   H,  W, iters   (HW: 1620 DAVIS, 1590 YTVOS, ragged tiles, tiny)
    (1, 5, 64, 512, 128, 30, 54, 4),
    (1, 3, 64, 512, 128, 30, 53, 4),
    (1, 1, 64, 512, 64, 24, 24, 2),
    (2, 2, 64, 512, 128, 24, 24, 3),
    (1, 6, 64, 512, 256, 30, 54, 4),
    (1, 2, 128, 512, 256, 17, 29, 1),
    (1, 1, 64, 512, 512, 12, 20, 4),
]


@pytest.mark.parametrize('family', ['generic', 'fused'])
@pytest.mark.parametrize('shape', SHAPES, ids=lambda s: 'x'.join(map(str, s)))
def test_memorize_and_readout_vs_oracle(shape, family):
    """Two chained memorize calls + readout (Lt = 2L) against the fp32 oracle, teacher-forced."""
    from swem_b200.synthetic import em_inputs
    B, N, Ck, Cv, L, H, W, I = shape
    topl = min(L, 64)
    _skip_unless_covered(family, B=B, N=N, Ck=Ck, Cv=Cv, HW=H * W, L=L, n_iters=I, topl=topl)
    cfg = dict(L=L, Cv=Cv, n_iters=I, tau=0.05, topl=64)
    core = _core(cfg, family)
    ref = O.OracleSWEMCore(n_bases=L, valdim=Cv, n_iters=I, tau=0.05, topl=64)
    tol = TOL[family]
    gen = torch.Generator().manual_seed(123)
    with torch.no_grad():
        for call in range(2):
            x, v, masks = em_inputs(B, N, Ck, Cv, H, W, seed=10 + call)
            if N > 1:
                masks[0, N - 1, 1] = 0                        # an empty object
            prior = ref.banks.prior()
            if prior is None:
                prior = dict(zip(('kappa', 'nu', 'zita'), O.random_init(B, N, Ck, L, Cv, generator=gen)))
            want = O.em_memorize(x, v, masks, prior, L, I, 0.05)
            ref.banks.commit(want)
            got = core.swem(x.to(DEV), v.to(DEV), masks.to(DEV), _to(prior, DEV))
            live = want['zita'] > 1e-3
            assert maxrel(got['zita'], want['zita']) < tol['bases'], 'zita'
            assert maxrel(got['kappa'], want['kappa'], live) < tol['bases'], 'kappa'
            assert maxrel(got['nu'], want['nu'], live) < tol['bases'], 'nu'
            assert torch.isfinite(got['kappa']).all() and torch.isfinite(got['nu']).all()
        core.memories['first'].bases = _to(ref.banks.first, DEV)
        core.memories['update'].bases = _to(ref.banks.update, DEV)
        q, qv, _ = em_inputs(B, 1, Ck, Cv, H, W, seed=99)
        feats, n = core.matching_features(q.to(DEV), qv[:, 0].to(DEV))
        want_feats, wn = ref.matching_features(q, qv[:, 0])
        assert n == wn == N
        assert maxrel(feats[:, :Cv], want_feats[:, :Cv]) < tol['feat'], 'mem_out'
        assert maxrel(feats[:, 2 * Cv:], want_feats[:, 2 * Cv:]) < tol['feat'], 'S'
        assert torch.equal(feats[:, Cv:2 * Cv].cpu(), want_feats[:, Cv:2 * Cv])


@pytest.mark.parametrize('family', ['generic', 'fused'])
def test_readout_properties_full_size(family):
    """Size-independent properties at the DAVIS-17 shape: rows of P sum to one (constant values are
    reproduced), S ranks pair up to one, the readout ignores the scale of the query key."""
    from swem_b200.synthetic import em_inputs
    B, N, Ck, Cv, L, H, W = 1, 5, 64, 512, 128, 30, 54
    _skip_unless_covered(family, B=B, N=N, Ck=Ck, Cv=Cv, HW=H * W, L=L, what='readout')
    core = _core(dict(L=L, Cv=Cv, n_iters=4, tau=0.05, topl=64), family)
    g = torch.Generator().manual_seed(5)
    mk = lambda: dict(kappa=torch.randn(B, N, 2, Ck, L, generator=g).to(DEV),
                      nu=torch.full((B, N, 2, Cv, L), 0.75, device=DEV),
                      zita=torch.ones(B, N, 2, 1, L, device=DEV))
    core.memories['first'].bases, core.memories['update'].bases = mk(), mk()
    q, qv, _ = em_inputs(B, 1, Ck, Cv, H, W, seed=3)
    with torch.no_grad():
        f1, _ = core.matching_features(q.to(DEV), qv[:, 0].to(DEV))
        f2, _ = core.matching_features(3.0 * q.to(DEV), qv[:, 0].to(DEV))
    assert (f1[:, :Cv] - 0.75).abs().max().item() < 2e-3          # convex combination of a constant
    S = f1[:, 2 * Cv:]
    assert S.min().item() >= 0 and S.max().item() <= 1
    assert (S[:, :64] + S[:, 64:] - 1).abs().max().item() < 1e-5
    assert maxrel(f2[:, 2 * Cv:], S) < TOL[family]['feat']


@pytest.mark.parametrize('family', ['generic', 'fused'])
def test_em_pixel_permutation_invariance(family):
    """The bases are sums over pixels: shuffling the pixel order must not change them."""
    from swem_b200.synthetic import em_inputs
    B, N, Ck, Cv, L, H, W, I = 1, 2, 64, 512, 128, 30, 54, 4
    _skip_unless_covered(family, B=B, N=N, Ck=Ck, Cv=Cv, HW=H * W, L=L)
    core = _core(dict(L=L, Cv=Cv, n_iters=I, tau=0.05, topl=64), family)
    x, v, masks = em_inputs(B, N, Ck, Cv, H, W, seed=1)
    prior = dict(zip(('kappa', 'nu', 'zita'), O.random_init(B, N, Ck, L, Cv, generator=torch.Generator().manual_seed(2))))
    perm = torch.randperm(H * W, generator=torch.Generator().manual_seed(3))
    shuf = lambda t: t.flatten(-2)[..., perm].view_as(t)
    with torch.no_grad():
        a = core.swem(x.to(DEV), v.to(DEV), masks.to(DEV), _to(prior, DEV))
        b = core.swem(shuf(x).to(DEV), shuf(v).to(DEV), shuf(masks).to(DEV), _to(prior, DEV))
    live = a['zita'].cpu() > 1e-3
    tol = 1e-4 if family == 'generic' else 1e-2
    assert maxrel(b['kappa'], a['kappa'], live) < tol
    assert maxrel(b['nu'], a['nu'], live) < tol


def test_mask_prep_kernel_matches_torch():
    from swem_b200 import SWEM, make_config
    g = torch.Generator().manual_seed(0)
    B, N, Hm, Wm, h16, w16 = 1, 3, 480, 854, 30, 54
    labels = torch.randint(0, N + 1, (B, Hm // 8, Wm // 8), generator=g)
    labels = labels.repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :Hm, :Wm]
    labels = torch.nn.functional.pad(labels, (0, Wm - labels.shape[2], 0, Hm - labels.shape[1]))
    hard = torch.nn.functional.one_hot(labels, N + 1).permute(0, 3, 1, 2).contiguous()
    soft = torch.rand(B, N + 1, 480, 864, generator=g)
    want = O.build_em_masks(hard, soft, h16, w16)
    got = SWEM._em_masks(SWEM, hard.to(DEV), soft.to(DEV), h16, w16)
    assert (got.cpu() - want).abs().max().item() < 1e-6


def test_errors_are_loud():
    from swem_b200 import SWEMCore
    core = SWEMCore(n_bases=16, valdim=32, n_iters=2, tau=0.05, topl=4)
    x = torch.randn(1, 16, 4, 4)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        core.swem(x, torch.randn(1, 1, 32, 4, 4), torch.rand(1, 1, 2, 4, 4))
    core = core.to(DEV)
    with pytest.raises(RuntimeError, match='memory is empty'):
        core.matching(x.to(DEV), torch.randn(1, 32, 4, 4, device=DEV))


def test_free_running_masks_vs_oracle():
    """Whole model, free-running (its own masks feed the next memorize), vs the CPU oracle with the
    same weights: >= 99.9 % pixel agreement per frame (north_star)."""
    from swem_b200 import SWEM, make_config
    from swem_b200.evaluator import evaluate_davis_seq
    from swem_b200.synthetic import davis_sequence
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False          # isolate the hot path: torch convs in full fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(0)
        cfg = make_config(keydim=64, n_bases=128, n_iters=4, topl=64)
        nets_cpu = SWEM(cfg).eval()
        model = SWEM(cfg).eval()
        model.load_state_dict(nets_cpu.state_dict())
        model = model.to(DEV)
        T, N, h, w = 6, 3, 240, 432
        frames, init = davis_sequence(T, N, seed=1, size=(h, w))
        prior = dict(zip(('kappa', 'nu', 'zita'), O.random_init(1, N, 64, 128, 512, generator=torch.Generator().manual_seed(4))))
        oracle = O.OracleSWEM(nets_cpu, 128, 4, 0.05, 64)
        # same initial bases on both sides
        model.swem_core.random_init = lambda size, norm_dim=-2, dtype=None, device=None: tuple(t.to(device) for t in (prior['kappa'], prior['nu'], prior['zita']))
        real_init = O.random_init
        O.random_init = lambda *a, **k: (prior['kappa'], prior['nu'], prior['zita'])
        try:
            want = torch.stack(O.run_davis_sequence(oracle, frames, init, (h, w)))
        finally:
            O.random_init = real_init
        got, _ = evaluate_davis_seq(model, frames.to(DEV), [init.to(DEV)] + [None] * (T - 1), (h, w))
        got = torch.stack(got).cpu()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    per_frame = (got == want).flatten(1).float().mean(dim=1)
    assert per_frame.min().item() >= 0.999, per_frame.tolist()
