"""GPU-side diagnostics for the fused EM kernel: per-tensor / per-side errors against the fp32 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import swem_oracle as O
from swem_b200 import SWEMCore, _lib
from swem_b200.synthetic import em_inputs

def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()

def run(B, N, H, W, I, seed=0, zp=0.0):
    Ck, Cv, L = 64, 512, 128
    x, v, masks = em_inputs(B, N, Ck, Cv, H, W, seed=seed)
    g = torch.Generator().manual_seed(7)
    prior = dict(zip(('kappa', 'nu', 'zita'), O.random_init(B, N, Ck, L, Cv, generator=g)))
    if zp:
        prior['zita'] = prior['zita'] + zp * torch.rand(prior['zita'].shape, generator=g)
        prior['nu'] = torch.randn(prior['nu'].shape, generator=g)
    trace = {}
    want = O.em_memorize(x, v, masks, prior, L, I, 0.05, trace=trace)
    outs = {}
    for fam, path in (('generic', _lib.PATH_GENERIC), ('fused', _lib.PATH_FUSED)):
        core = SWEMCore(n_bases=L, valdim=Cv, n_iters=I, tau=0.05, topl=64).cuda().eval()
        core.em_path = path
        with torch.no_grad():
            got = core.swem(x.cuda(), v.cuda(), masks.cuda(), {k: t.cuda() for k, t in prior.items()}, return_z=True)
        torch.cuda.synchronize()
        outs[fam] = {k: t.cpu() for k, t in got.items()}
    print(f'--- B={B} N={N} HW={H*W} I={I} zp={zp}')
    zl = trace['z'][-1]
    for fam, got in outs.items():
        z = got['z'].view_as(zl)
        line = f'{fam:8s} z {rel(z, zl):.2e} (bg {rel(z[:, :, 0], zl[:, :, 0]):.2e} fg {rel(z[:, :, 1], zl[:, :, 1]):.2e})'
        live = want['zita'] > 1e-3
        for k in ('zita', 'kappa', 'nu'):
            m = live.expand_as(want[k])
            line += f' | {k} {rel(got[k], want[k]):.2e} live {rel(got[k][m], want[k][m]):.2e}'
        print(line)
    f = outs['fused']
    z = f['z'].view_as(zl)
    d = (z - zl).abs()
    idx = d.flatten().argmax().item()
    print('  worst z idx', tuple(int(i) for i in torch.unravel_index(torch.tensor(idx), d.shape)), 'got', z.flatten()[idx].item(), 'want', zl.flatten()[idx].item())
    colsum = z.sum(dim=3)
    print('  fused: zita - zita_prior vs colsum(z_fused):', rel(f['zita'].view_as(prior['zita']) - prior['zita'], colsum.view_as(prior['zita'])))
    print('  sample zita fused', f['zita'].flatten()[:6].tolist(), 'want', want['zita'].flatten()[:6].tolist())
    print('  sample kappa fused', f['kappa'][0, 0, 0, :3, :3].flatten().tolist(), 'want', want['kappa'][0, 0, 0, :3, :3].flatten().tolist())
    print('  sample nu fused', f['nu'][0, 0, 0, :2, :4].flatten().tolist(), 'want', want['nu'][0, 0, 0, :2, :4].flatten().tolist())
    return outs

if __name__ == '__main__' and len(sys.argv) == 1:
    run(1, 1, 8, 16, 1)
    run(1, 1, 8, 16, 1, zp=3.0)
    run(1, 1, 6, 10, 1)
    run(1, 1, 16, 16, 1)
    run(1, 1, 8, 16, 2)
    run(1, 2, 30, 54, 1)
    run(1, 2, 30, 54, 4)


def free_running():
    from swem_b200 import SWEM, make_config
    from swem_b200.evaluator import evaluate_davis_seq
    from swem_b200.synthetic import davis_sequence
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    cfg = make_config(keydim=64, n_bases=128, n_iters=4, topl=64)
    nets_cpu = SWEM(cfg).eval()
    T, N, h, w = 6, 3, 240, 432
    frames, init = davis_sequence(T, N, seed=1, size=(h, w))
    prior = dict(zip(('kappa', 'nu', 'zita'), O.random_init(1, N, 64, 128, 512, generator=torch.Generator().manual_seed(4))))
    oracle = O.OracleSWEM(nets_cpu, 128, 4, 0.05, 64)
    real_init = O.random_init
    O.random_init = lambda *a, **k: (prior['kappa'], prior['nu'], prior['zita'])
    want = torch.stack(O.run_davis_sequence(oracle, frames, init, (h, w)))
    O.random_init = real_init
    for fam, path in (('generic', _lib.PATH_GENERIC), ('fused', _lib.PATH_AUTO)):
        model = SWEM(cfg).eval()
        model.load_state_dict(nets_cpu.state_dict())
        model = model.cuda()
        model.swem_core.em_path = path
        model.swem_core.random_init = lambda size, norm_dim=-2, dtype=None, device=None: tuple(t.to(device) for t in (prior['kappa'], prior['nu'], prior['zita']))
        got, _ = evaluate_davis_seq(model, frames.cuda(), [init.cuda()] + [None] * (T - 1), (h, w))
        got = torch.stack(got).cpu()
        print(fam, 'free-running agreement per frame', [round(v, 5) for v in (got == want).flatten(1).float().mean(dim=1).tolist()])


if __name__ == '__main__' and len(sys.argv) == 1:
    free_running()


def free_running_480p(T=7, N=5):
    """BASELINE configs[1]-shaped: 480x864, 5 objects; GPU (AUTO = fused) vs CPU oracle, same weights."""
    from swem_b200 import SWEM, make_config
    from swem_b200.evaluator import evaluate_davis_seq
    from swem_b200.synthetic import davis_sequence
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    cfg = make_config(keydim=64, n_bases=128, n_iters=4, topl=64)
    nets_cpu = SWEM(cfg).eval()
    h, w = 480, 864
    frames, init = davis_sequence(T, N, seed=1, size=(h, w))
    prior = dict(zip(('kappa', 'nu', 'zita'), O.random_init(1, N, 64, 128, 512, generator=torch.Generator().manual_seed(4))))
    oracle = O.OracleSWEM(nets_cpu, 128, 4, 0.05, 64)
    real_init = O.random_init
    O.random_init = lambda *a, **k: (prior['kappa'], prior['nu'], prior['zita'])
    want = torch.stack(O.run_davis_sequence(oracle, frames, init, (h, w)))
    O.random_init = real_init
    G, A = _lib.PATH_GENERIC, _lib.PATH_AUTO
    for fam, em_path, ro_path in (('generic/generic', G, G), ('fusedEM/genericRO', A, G), ('genericEM/fusedRO', G, A), ('fused/fused', A, A)):
        model = SWEM(cfg).eval()
        model.load_state_dict(nets_cpu.state_dict())
        model = model.cuda()
        model.swem_core.em_path, model.swem_core.readout_path = em_path, ro_path
        model.swem_core.random_init = lambda size, norm_dim=-2, dtype=None, device=None: tuple(t.to(device) for t in (prior['kappa'], prior['nu'], prior['zita']))
        got, _ = evaluate_davis_seq(model, frames.cuda(), [init.cuda()] + [None] * (T - 1), (h, w))
        got = torch.stack(got).cpu()
        print(fam, '480p free-running agreement per frame', [round(v, 5) for v in (got == want).flatten(1).float().mean(dim=1).tolist()])


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == '480p':
    free_running_480p()
