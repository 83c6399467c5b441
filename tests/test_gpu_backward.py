"""Gradients through the memory (training path, BASELINE configs[4]) against the CPU oracle's autograd in fp64.

Reference contract: memorize -> gradient to v and to the prior nu only; matching -> gradient
to the raw query key, the nu of both banks, and (through torch) qv.  Tolerance: max-rel 1e-3 on every gradient."""
import pytest
import torch

from oracle import swem_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

from test_gpu_parity import _core, _skip_unless_covered, _to, check, maxrel      # noqa: E402

SHAPES = {   # B, N, Ck, Cv, L, H, W, topl
    'small': (2, 2, 16, 24, 8, 5, 7, 4),
    'train': (2, 2, 64, 512, 128, 24, 24, 64),       # HW = 576: the reference's 384x384 training crops
}


def _problem(shape, seed=0):
    B, N, Ck, Cv, L, H, W, topl = SHAPES[shape]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Ck, H, W, generator=g) * 2.3
    v = torch.randn(B, N, Cv, H, W, generator=g) * 1.8
    fg = (torch.rand(B, N, H, W, generator=g) < 0.3).float()
    fg[:, -1, : H // 2] = 0
    masks = torch.stack([1 - fg, fg], dim=2)
    prior = dict(zip(('kappa', 'nu', 'zita'), O.random_init(B, N, Ck, L, Cv, generator=g)))
    prior['nu'] = torch.randn(B, N, 2, Cv, L, generator=g)
    prior['zita'] = prior['zita'] + torch.rand(B, N, 2, 1, L, generator=g) * 3
    return x, v, masks, prior


@pytest.mark.parametrize('family', ['generic', 'fused'])
@pytest.mark.parametrize('shape', ['small', 'train'])
def test_em_backward(shape, family):
    B, N, Ck, Cv, L, H, W, topl = SHAPES[shape]
    _skip_unless_covered(family, B=B, N=N, Ck=Ck, Cv=Cv, HW=H * W, L=L, n_iters=1, topl=topl)
    x, v, masks, prior = _problem(shape)
    G = torch.randn(B, N, 2, Cv, L, generator=torch.Generator().manual_seed(5))
    # oracle, fp64 autograd, one EM iteration (well conditioned: the responsibilities agree to ~1e-6)
    d = lambda t: t.double()
    v64 = d(v).requires_grad_()
    p64 = {k: d(t) for k, t in prior.items()}
    p64['nu'].requires_grad_()
    want = O.em_memorize(d(x), v64, d(masks), p64, L, 1, 0.05)
    (want['nu'] * d(G)).sum().backward()

    core = _core(dict(L=L, Cv=Cv, n_iters=1, tau=0.05, topl=topl), family).train()
    vg = v.to(DEV).requires_grad_()
    pg = _to(prior, DEV)
    pg['nu'].requires_grad_()
    got = core.swem(x.to(DEV), vg, masks.to(DEV), pg)
    assert got['nu'].requires_grad and not got['kappa'].requires_grad and not got['zita'].requires_grad
    (got['nu'] * G.to(DEV)).sum().backward()
    tol = 2e-4 if family == 'generic' else 1e-3
    check('nu', maxrel(got['nu'], want['nu']), tol)
    check('grad_v', maxrel(vg.grad, v64.grad), tol)
    check('grad_nu_prior', maxrel(pg['nu'].grad, p64['nu'].grad), tol)
    # multi-iteration: the backward must be the exact linear map of ITS OWN saved responsibilities
    core.n_iters = 3
    vg.grad = None
    bases = core.swem(x.to(DEV), vg.detach(), masks.to(DEV), _to(prior, DEV), return_z=True)
    got = core.swem(x.to(DEV), vg, masks.to(DEV), _to(prior, DEV))
    (got['nu'] * G.to(DEV)).sum().backward()
    w = (G.to(DEV) / bases['zita']).double()
    ref = torch.einsum('bnsdl,bnspl->bndp', w, bases['z'].double()).view_as(vg)
    check('grad_v_linear', maxrel(vg.grad, ref), 2e-3 if family == 'fused' else 1e-5)   # fused: z differs run to run (reduce order)


@pytest.mark.parametrize('family', ['generic', 'fused'])
@pytest.mark.parametrize('shape', ['small', 'train'])
def test_readout_backward(shape, family):
    B, N, Ck, Cv, L, H, W, topl = SHAPES[shape]
    _skip_unless_covered(family, B=B, N=N, Ck=Ck, Cv=Cv, HW=H * W, L=L, n_iters=1, topl=topl, what='readout')
    g = torch.Generator().manual_seed(9)
    qk = torch.randn(B, Ck, H, W, generator=g) * 2.3
    qv = torch.randn(B, Cv, H, W, generator=g)
    banks = []
    for k in range(2):
        kap, _, _ = O.random_init(B, N, Ck, L, Cv, generator=g)
        banks.append({'kappa': kap, 'nu': torch.randn(B, N, 2, Cv, L, generator=g), 'zita': torch.ones(B, N, 2, 1, L)})
    chans = 2 * Cv + 2 * topl
    G = torch.randn(B * N, chans, H, W, generator=g)

    d = lambda t: t.double()
    q64, qv64 = d(qk).requires_grad_(), d(qv).requires_grad_()
    nu64 = [d(b['nu']).requires_grad_() for b in banks]
    mk = torch.cat([d(b['kappa']) for b in banks], dim=-1)
    S, mem = O.readout(O.l2norm(q64, 1), O.l2norm(mk, -2), torch.cat(nu64, dim=-1), 0.05, topl)
    want = torch.cat([mem.flatten(end_dim=1), qv64.unsqueeze(1).expand(-1, N, -1, -1, -1).flatten(end_dim=1), S], dim=1)
    (want * d(G)).sum().backward()

    core = _core(dict(L=L, Cv=Cv, n_iters=1, tau=0.05, topl=topl), family).train()
    qg, qvg = qk.to(DEV).requires_grad_(), qv.to(DEV).requires_grad_()
    bg = [_to(b, DEV) for b in banks]
    for b in bg:
        b['nu'].requires_grad_()
    core.memories['first'].bases, core.memories['first'].n_objs = bg[0], N
    core.memories['update'].bases = bg[1]
    feats, n = core.matching_features(qg, qvg)
    assert n == N and feats.shape == want.shape
    (feats * G.to(DEV)).sum().backward()
    tol = 2e-4 if family == 'generic' else 1e-2
    check('feats', maxrel(feats, want), tol)
    check('grad_qk', maxrel(qg.grad, q64.grad), 1e-3)
    check('grad_qv', maxrel(qvg.grad, qv64.grad), 1e-5)
    for k in range(2):
        check(f'grad_nu{k}', maxrel(bg[k]['nu'].grad, nu64[k].grad), 1e-3)
    # the torch re-evaluation of the backward must tell the same story as the kernels
    from swem_b200.autograd import ReadoutFunction
    native = (qg.grad.clone(), [b['nu'].grad.clone() for b in bg])
    qg.grad = None
    for b in bg:
        b['nu'].grad = None
    ReadoutFunction.native_backward = False
    try:
        feats2, _ = core.matching_features(qg, qvg)
        (feats2 * G.to(DEV)).sum().backward()
    finally:
        ReadoutFunction.native_backward = True
    check('grad_qk_vs_torch', maxrel(native[0], qg.grad), 1e-3)
    for k in range(2):
        check(f'grad_nu{k}_vs_torch', maxrel(native[1][k], bg[k]['nu'].grad), 1e-3)


def test_training_step_runs_end_to_end():
    """3-frame clip, 2 objects (the reference's one_step, swem_trainer.py:59-108, restated): loss.backward()
    reaches encoders, fusion conv and decoder through the CUDA memory; BN stays in eval like the reference (:37-39)."""
    from swem_b200 import SWEM, make_config
    from swem_b200.synthetic import davis_sequence
    torch.manual_seed(0)
    model = SWEM(make_config(keydim=64, n_bases=128, n_iters=4, topl=64)).to(DEV).train()
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    T, N, h, w = 3, 2, 96, 96
    frames, init = davis_sequence(T, N, seed=3, size=(h, w))
    frames, init = frames.to(DEV), init.to(DEV)
    label = init.argmax(dim=1)
    mk16, _, s16, _, _ = model('encode_key', frames[:, 0])
    mv16 = model('encode_value', frames[:, 0], init, s16)
    model('init', mk16, mv16, init.long())
    loss = 0
    for i in range(1, T):
        qk16, qv16, s16, s8, s4 = model('encode_key', frames[:, i])
        ctx, n = model('match', qk16, qv16)
        logits, prob = model('segment', n, ctx, s8, s4, None, (h, w))
        loss = loss + torch.nn.functional.cross_entropy(logits, label)
        if i < T - 1:
            hard = torch.nn.functional.one_hot(prob.argmax(1), N + 1).permute(0, 3, 1, 2)
            mv16 = model('encode_value', frames[:, i], prob, s16)
            model('memorize', qk16, mv16, hard, prob)
    loss.backward()
    assert torch.isfinite(loss)
    for name in ('key_encoder.conv1.weight', 'value_encoder.conv1.weight', 'key_proj.key_proj.weight',
                 'swem_core.fusion_layer.layer_f.weight', 'decoder.pred.weight'):
        gr = dict(model.named_parameters())[name].grad
        assert gr is not None and torch.isfinite(gr).all() and gr.abs().max() > 0, name


def test_training_step_loss_and_grads_match_the_reference_autograd():
    """BASELINE configs[4] / SURVEY section 8(d) config 5: ONE full training step (3-frame clip, 2 objects, the second one empty on
    the first frame, swem_trainer.py:59-108 restated) through the CUDA memory -- forward kernels + swem_em_backward /
    swem_readout_backward behind autograd -- against the same step with the oracle core under torch autograd on the same
    device and parameters.  The oracle core gets the reference's gradient structure: E / M / W steps under no_grad
    (modules.py:93,112,122), only nu = (zita_ nu_ + v z) / zita differentiable (:164-165).  Loss and the gradients of key_proj,
    key_comp, the fusion layer and the encoders / decoder ends must agree."""
    import torch.nn.functional as F
    from oracle import swem_oracle as O
    from swem_b200 import SWEM, make_config
    from swem_b200.synthetic import davis_sequence
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(0)
        model = SWEM(make_config(keydim=64, n_bases=128, n_iters=4, topl=64)).to(DEV).train()
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
        T, N, h, w = 3, 2, 96, 96
        frames, init = davis_sequence(T, N, seed=3, size=(h, w))
        init[:, 0] += init[:, N]                              # the last object is empty (padded sample, video_dataset.py:334-335)
        init[:, N] = 0
        frames, init = frames.to(DEV), init.to(DEV)
        label = init.argmax(dim=1)

        def ref_memorize(core, qk, qv, masks):               # OracleSWEMCore.memorize with the reference's no_grad structure
            prior = core.banks.prior()
            B, Ck, H, W = qk.shape
            Nn, Cv, L = masks.shape[1], qv.shape[2], core.n_bases
            if prior is None:
                kp, nup, zp = (t.to(qk.device) for t in O.random_init(B, Nn, Ck, L, Cv, generator=torch.Generator().manual_seed(11)))
            else:
                kp, nup, zp = prior['kappa'], prior['nu'], prior['zita']
            with torch.no_grad():
                xf = qk.flatten(start_dim=-2)[:, None, None]
                x_t = xf.transpose(-2, -1)
                mk = masks.flatten(start_dim=-2).unsqueeze(-1)
                wts, kappa = mk.clone(), kp.clone()
                for it in range(core.n_iters):
                    z = O.e_step(x_t, kappa, wts, core.tau)
                    kappa, zita = O.m_step(z, xf, kp, zp)
                    if it < core.n_iters - 1:
                        wts = O.w_step(kappa, x_t, mk, core.tau)
            nu = (zp * nup + torch.matmul(qv.flatten(start_dim=-2).unsqueeze(2), z)) / zita
            core.banks.commit({'kappa': kappa, 'nu': nu, 'zita': zita})

        def step(use_kernels):
            if use_kernels:
                torch.manual_seed(11)
                ri = model.swem_core.random_init
                model.swem_core.random_init = lambda size, norm_dim=-2, dtype=None, device=None: tuple(
                    t.to(device) for t in O.random_init(size[0], size[1], size[3], size[4], 512, generator=torch.Generator().manual_seed(11)))
                enc_k = lambda f: model('encode_key', f)
                enc_v = lambda f, m, s: model('encode_value', f, m, s)
                init_f = lambda k, v, m: model('init', k, v, m.long())
                mem_f = lambda k, v, hd, sf: model('memorize', k, v, hd, sf)
                match = lambda k, v: model('match', k, v)
                seg = lambda n, c, s8, s4: model('segment', n, c, s8, s4, None, (h, w))
            else:
                om = O.OracleSWEM(model, 128, 4, 0.05, 64)
                enc_k, enc_v = om.encode_key, om.encode_value
                def mem_f(k, v, hd, sf):
                    ref_memorize(om.core, k, v, O.build_em_masks(hd, sf, *k.shape[-2:]))
                def init_f(k, v, m):
                    om.core.empty()
                    mem_f(k, v, m, m.float())
                match = om.match
                seg = lambda n, c, s8, s4: om.segment(n, c, s8, s4, (h, w))
            model.zero_grad(set_to_none=True)
            mk16, _, s16, _, _ = enc_k(frames[:, 0])
            init_f(mk16, enc_v(frames[:, 0], init, s16), init)
            loss = 0
            for i in range(1, T):
                qk16, qv16, s16, s8, s4 = enc_k(frames[:, i])
                ctx, n = match(qk16, qv16)
                logits, prob = seg(n, ctx, s8, s4)
                loss = loss + F.cross_entropy(logits, label)
                if i < T - 1:
                    hard = F.one_hot(prob.argmax(1), N + 1).permute(0, 3, 1, 2)
                    mem_f(qk16, enc_v(frames[:, i], prob, s16), hard, prob)
            loss.backward()
            if use_kernels:
                model.swem_core.random_init = ri
            return loss.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

        loss_k, grads_k = step(True)
        loss_o, grads_o = step(False)
    finally:
        torch.backends.cudnn.allow_tf32 = True
    check('train_loss', abs(loss_k.item() - loss_o.item()) / abs(loss_o.item()), 1e-4)
    names = ['key_proj.key_proj.weight', 'key_comp.weight', 'swem_core.fusion_layer.layer_f.weight',
             'swem_core.fusion_layer.layer_a.weight', 'key_encoder.conv1.weight', 'value_encoder.conv1.weight', 'decoder.pred.weight']
    for name in names:
        assert name in grads_k and name in grads_o, (name, sorted(grads_k)[:8])
        check('grad ' + name, maxrel(grads_k[name], grads_o[name]), 2e-3)


def test_memory_dropout_forward_and_backward_vs_reference_golden(golden):
    """SWEMCore.matching_features in training mode with p_drop > 0 (the reference's memory dropout, modules.py:258-263; hard-wired
    to 0.0 by the reference's constructor, kept for the API: SURVEY 8b) against the unmodified reference's outputs AND its
    autograd's gradients (tests/golden/drop_small.pt): the mask is drawn like the reference draws it (CPU global generator), the
    attention is renormalised with + 1e-6, S ignores the mask; gradients to the raw query key and to the memory values of both
    banks through swem_readout_backward.  In eval mode the same core ignores p_drop."""
    from swem_b200 import SWEMCore, _lib
    fx = golden('drop_small')
    c = fx['cfg']
    B, N, Cv, Lt, H, W = c['B'], c['N'], c['Cv'], c['Lt'], c['H'], c['W']
    L = Lt // 2
    core = SWEMCore(n_bases=L, valdim=Cv, n_iters=1, tau=c['tau'], topl=c['topl']).to(DEV).train()
    core.em_path = core.readout_path = _lib.PATH_GENERIC
    core.p_drop = c['p_drop']
    nus = []
    for name, sl in (('first', slice(0, L)), ('update', slice(L, Lt))):
        nu = fx['mv'][..., sl].contiguous().to(DEV).requires_grad_()
        nus.append(nu)
        core.memories[name].bases = dict(kappa=fx['mk'][..., sl].contiguous().to(DEV), nu=nu, zita=torch.ones(B, N, 2, 1, L, device=DEV))
        core.memories[name].n_objs = N
    q = fx['q'].to(DEV).requires_grad_()
    qv = torch.zeros(B, Cv, H, W, device=DEV)
    torch.manual_seed(c['mask_seed'])
    feats, n = core.matching_features(q, qv)
    mem_out, S = feats[:, :Cv].reshape(B, N, Cv, H, W), feats[:, 2 * Cv:]
    check('mem_out', maxrel(mem_out, fx['mem_out']), 2e-4)
    check('S', maxrel(S, fx['S']), 2e-4)
    loss = (mem_out * fx['w_mem'].to(DEV)).sum() + (S * fx['w_s'].to(DEV)).sum()
    loss.backward()
    check('grad_q', maxrel(q.grad, fx['grad_q']), 2e-3)
    check('grad_nu', maxrel(torch.cat([nu.grad for nu in nus], dim=-1), fx['grad_mv']), 2e-3)
    # the dropped bases receive no gradient through the attention
    dropped = (fx['mask'].reshape(B, N, 1, 1, Lt) == 0).expand(B, N, 2, Cv, Lt)
    assert torch.cat([nu.grad for nu in nus], dim=-1).cpu()[dropped].abs().max() == 0
    core.eval()
    with torch.no_grad():
        f_eval, _ = core.matching_features(q.detach(), qv)
    want_S, want_m = O.readout(O.l2norm(fx['q'], dim=1), O.l2norm(fx['mk'], dim=-2), fx['mv'], c['tau'], c['topl'])
    check('eval_mem', maxrel(f_eval[:, :Cv].reshape(B, N, Cv, H, W), want_m), 2e-4)
