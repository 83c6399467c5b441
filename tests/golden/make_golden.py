"""Generate the golden fixtures in this directory by running the UNMODIFIED reference on CPU.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes ``core_small.pt``, ``core_prod.pt``, ``steps_small.pt``, ``model_tiny.pt`` next to this
file (``--wide``: only ``core_wide.pt``, the reference's class default of 256 bases per side; ``--mkm``: only ``mkm_small.pt``,
the kernelised-memory readout branch; ``--drop``: only ``drop_small.pt``, the memory-dropout branch with the reference's gradients).  Each fixture holds the seeded inputs and the reference's outputs; tests replay the inputs
through ``oracle/swem_oracle.py`` (CPU) and through the CUDA library (GPU).
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_shim                      # noqa: E402
from swem_b200.evaluator import evaluate_davis_seq   # noqa: E402  (loop only; model is the reference's)
from swem_b200.synthetic import em_inputs, fill_deterministic_, rectangle_masks, smooth_video  # noqa: E402


def core_case(ref, name, B, n_seq, Ck, Cv, L, H, W, n_iters, tau, topl, seed, empty_obj=None):
    """n_seq: list of object counts per memorize call (may grow).  Records bases after each
    memorize and (S, mem_out, kappa/nu read) after each."""
    core = ref.SWEMCore(n_bases=L, valdim=Cv, n_iters=n_iters, tau=tau, topl=topl)
    core.eval()
    calls = []
    for i, N in enumerate(n_seq):
        x, v, masks = em_inputs(B, N, Ck, Cv, H, W, seed=seed + i)
        # soft-ish masks after the first call, like predicted masks
        if i > 0:
            g = torch.Generator().manual_seed(seed + 100 + i)
            soft = torch.rand(B, N, H, W, generator=g)
            fg = masks[:, :, 1]
            masks = torch.stack([(1 - fg) * (1 - soft), fg * soft], dim=2)
        if empty_obj is not None and N > empty_obj[1]:
            masks[empty_obj[0], empty_obj[1], 1] = 0      # an object with an all-zero fg mask
        rng_seed = 1000 + seed + i
        torch.manual_seed(rng_seed)
        with torch.no_grad():
            core.memorize(x, v, masks)
        q, _, _ = em_inputs(B, 1, Ck, Cv, H, W, seed=seed + 50 + i)
        with torch.no_grad():
            mk, mv = core.get_mem()
            S, mem_out = core.get_affinity(ref.l2norm(q, dim=1), ref.l2norm(mk, dim=-2), mv)
        upd = core.memories['update'].bases if core.memories['update'].bases is not None else core.memories['first'].bases
        calls.append(dict(x=x, v=v, masks=masks, rng_seed=rng_seed, q=q,
                          kappa=upd['kappa'].clone(), nu=upd['nu'].clone(), zita=upd['zita'].clone(),
                          first_kappa=core.memories['first'].bases['kappa'].clone(),
                          S=S.clone(), mem_out=mem_out.clone()))
    fx = dict(cfg=dict(B=B, Ck=Ck, Cv=Cv, L=L, H=H, W=W, n_iters=n_iters, tau=tau, topl=topl), calls=calls)
    torch.save(fx, os.path.join(HERE, name + '.pt'))
    print(name, os.path.getsize(os.path.join(HERE, name + '.pt')) // 1024, 'KiB')


def steps_case(ref, name, seed=3):
    """Per-step intermediates: drive the reference's swe/swm/sww step functions one by one."""
    B, N, Ck, Cv, L, H, W, I, tau = 1, 2, 16, 24, 8, 5, 7, 3, 0.05
    core = ref.SWEMCore(n_bases=L, valdim=Cv, n_iters=I, tau=tau, topl=4)
    x, v, masks = em_inputs(B, N, Ck, Cv, H, W, seed=seed)
    torch.manual_seed(77)
    kappa_, nu_, zita_ = core.random_init((B, N, 2, Ck, L), dtype=x.type(), device=x.device)
    zita_ = zita_ + torch.rand(B, N, 2, 1, L) * 3.0          # a "used" prior
    xf = x.flatten(start_dim=-2).unsqueeze(1).unsqueeze(2)
    x_t = xf.transpose(-2, -1)
    m = masks.flatten(start_dim=-2).unsqueeze(-1)
    weights, kappa = m.clone(), kappa_.clone()
    zs, ws, ks, zitas = [], [weights.clone()], [], []
    with torch.no_grad():
        for i in range(I):
            z = core.swe_step(x_t, kappa, weights)
            kappa, zita = core.swm_step(z, xf, kappa_, zita_)
            zs.append(z.clone()); ks.append(kappa.clone()); zitas.append(zita.clone())
            if i < I - 1:
                weights = core.sww_step(kappa, x_t, m)
                ws.append(weights.clone())
        bases = core.swem(x, v, masks, dict(kappa=kappa_, nu=nu_, zita=zita_))
    fx = dict(cfg=dict(B=B, N=N, Ck=Ck, Cv=Cv, L=L, H=H, W=W, n_iters=I, tau=tau),
              x=x, v=v, masks=masks, kappa_prior=kappa_, nu_prior=nu_, zita_prior=zita_,
              z=zs, weights=ws, kappa_iter=ks, zita_iter=zitas,
              kappa=bases['kappa'], nu=bases['nu'], zita=bases['zita'])
    torch.save(fx, os.path.join(HERE, name + '.pt'))
    print(name, os.path.getsize(os.path.join(HERE, name + '.pt')) // 1024, 'KiB')


def model_case(SWEM, name, seed=5):
    """Whole reference model (deterministic weights) free-running on a tiny 2-object sequence."""
    cfg = ref_shim.model_config(keydim=64, n_bases=16, n_iters=4, topl=8, single_obj=False)
    model = SWEM(cfg).eval()
    with torch.no_grad():
        fill_deterministic_(model, seed=seed)
    T, h, w = 4, 64, 96
    frames = smooth_video(T, h, w, seed=seed)
    init = rectangle_masks(2, h, w, seed=seed)
    torch.manual_seed(4321)
    preds, scores = evaluate_davis_seq(model, frames, [init] + [None] * (T - 1), (h, w))
    with torch.no_grad():
        qk16, qv16, s16, s8, s4 = model('encode_key', frames[:, 0])
        mv16 = model('encode_value', frames[:, 0], init, s16)
    fx = dict(cfg=vars(cfg), weight_seed=seed, rng_seed=4321, frames=frames, init_mask=init,
              preds=torch.stack(preds), last_scores=scores[-1], qk16=qk16, qv16=qv16, mv16=mv16,
              s8_sum=s8.double().sum(), s4_sum=s4.double().sum())
    torch.save(fx, os.path.join(HERE, name + '.pt'))
    print(name, os.path.getsize(os.path.join(HERE, name + '.pt')) // 1024, 'KiB')


def mkm_case(ref, name, B=1, N=2, Ck=64, Cv=64, Lt=32, H=12, W=20, tau=0.05, topl=16, n_kernel=7, sigma=7, seed=11):
    """The reference's kernelised-memory readout (get_affinity with n_kernel > 0, modules.py:232-276 + gen_kernels :210-230; off
    by default, inference only) on a seeded clustered memory: raw inputs and the reference's (S, mem_out)."""
    core = ref.SWEMCore(n_bases=Lt // 2, valdim=Cv, n_iters=1, tau=tau, topl=topl)
    core.eval()
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, Ck, H, W, generator=g)
    # memory keys near some of the query pixels, so that the best-matching pixels of a basis are not arbitrary
    pix = torch.randint(0, H * W, (B, N, 2, Lt), generator=g)
    qf = q.flatten(start_dim=-2)                                             # B,Ck,HW
    mk = torch.stack([torch.stack([qf[b][:, pix[b, n].reshape(-1)].reshape(Ck, 2, Lt).permute(1, 0, 2) for n in range(N)]) for b in range(B)])
    mk = mk + 0.3 * torch.randn(B, N, 2, Ck, Lt, generator=g)
    mv = torch.randn(B, N, 2, Cv, Lt, generator=g)
    with torch.no_grad():
        S, mem_out = core.get_affinity(ref.l2norm(q, dim=1), ref.l2norm(mk, dim=-2), mv, n_kernel=n_kernel, sigma=sigma)
        S0, mem_out0 = core.get_affinity(ref.l2norm(q, dim=1), ref.l2norm(mk, dim=-2), mv)
    assert torch.equal(S, S0) and not torch.allclose(mem_out, mem_out0)      # the kernels reweight the attention only
    fx = dict(cfg=dict(B=B, N=N, Ck=Ck, Cv=Cv, Lt=Lt, H=H, W=W, tau=tau, topl=core.topl, n_kernel=n_kernel, sigma=sigma),
              q=q, mk=mk, mv=mv, S=S.clone(), mem_out=mem_out.clone())
    torch.save(fx, os.path.join(HERE, name + '.pt'))
    print(name, os.path.getsize(os.path.join(HERE, name + '.pt')) // 1024, 'KiB')


def drop_case(ref, name, B=2, N=2, Ck=64, Cv=64, Lt=32, H=10, W=12, tau=0.05, topl=16, p_drop=0.3, seed=21):
    """The reference's memory dropout (get_affinity in training mode with p_drop > 0, modules.py:258-263; p_drop is hard-wired to
    0.0 by the reference's constructor): raw inputs, the seed of the mask, the reference's (S, mem_out) and its autograd's gradients
    of a fixed linear functional with respect to the raw query key and the memory values."""
    core = ref.SWEMCore(n_bases=Lt // 2, valdim=Cv, n_iters=1, tau=tau, topl=topl)
    core.train()
    core.p_drop = p_drop
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, Ck, H, W, generator=g, requires_grad=True)
    mk = torch.randn(B, N, 2, Ck, Lt, generator=g)
    mv = torch.randn(B, N, 2, Cv, Lt, generator=g, requires_grad=True)
    w_mem = torch.randn(B, N, Cv, H, W, generator=g)
    w_s = torch.randn(B * N, 2 * core.topl, H, W, generator=g)
    mask_seed = 777
    torch.manual_seed(mask_seed)
    S, mem_out = core.get_affinity(ref.l2norm(q, dim=1), ref.l2norm(mk, dim=-2), mv)
    loss = (mem_out * w_mem).sum() + (S * w_s).sum()
    gq, gmv = torch.autograd.grad(loss, [q, mv])
    torch.manual_seed(mask_seed)
    mask = (torch.rand(B, N, 1, Lt, 1) > p_drop).float()
    assert 0 < mask.mean() < 1
    fx = dict(cfg=dict(B=B, N=N, Ck=Ck, Cv=Cv, Lt=Lt, H=H, W=W, tau=tau, topl=core.topl, p_drop=p_drop, mask_seed=mask_seed),
              q=q.detach(), mk=mk, mv=mv.detach(), w_mem=w_mem, w_s=w_s, mask=mask, S=S.detach().clone(), mem_out=mem_out.detach().clone(),
              grad_q=gq.clone(), grad_mv=gmv.clone())
    torch.save(fx, os.path.join(HERE, name + '.pt'))
    print(name, os.path.getsize(os.path.join(HERE, name + '.pt')) // 1024, 'KiB')


def main():
    assert ref_shim.available(), 'reference checkout not found'
    torch.set_num_threads(1)                       # fixed reduction order inside ATen
    ref = ref_shim.load_modules()
    if '--mkm' in sys.argv:                        # added later: only this fixture is (re)generated
        mkm_case(ref, 'mkm_small')
        return
    if '--drop' in sys.argv:
        drop_case(ref, 'drop_small')
        return
    if '--wide' in sys.argv:                       # added later: only this fixture is (re)generated
        core_case(ref, 'core_wide', B=1, n_seq=[1, 1], Ck=64, Cv=512, L=256, H=6, W=10, n_iters=3, tau=0.05,
                  topl=64, seed=31)
        return
    core_case(ref, 'core_small', B=2, n_seq=[2, 2, 3], Ck=16, Cv=24, L=8, H=5, W=7, n_iters=3, tau=0.05,
              topl=4, seed=11, empty_obj=(1, 1))
    core_case(ref, 'core_prod', B=1, n_seq=[1, 1], Ck=64, Cv=512, L=128, H=6, W=10, n_iters=4, tau=0.05,
              topl=64, seed=21)
    steps_case(ref, 'steps_small')
    SWEM, _ = ref_shim.load_swem()
    model_case(SWEM, 'model_tiny')


if __name__ == '__main__':
    main()
