"""Generate the golden fixtures in tests/golden/ by running the UNTOUCHED reference.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The reference is imported from /root/reference (read-only) behind two import shims:
`ml_collections` (absent here) -> soft_truncation_b200.config_dict.ConfigDict, and `op` ->
a module exposing the reference's own CPU implementation `upfirdn2d_native`, extracted from
op/upfirdn2d.py by AST so that importing it does not trigger the 2-minute CUDA JIT build.
Nothing from the reference is written into the repo except the numeric vectors it produces.

Model weights are NOT stored (61.8 M parameters): they are regenerated from a seed by
oracle.ref_model.make_state_dict and loaded into the reference model with strict=True, which
also pins the parameter names and shapes.
"""
import ast
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('ST_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)

from soft_truncation_b200.config_dict import ConfigDict  # noqa: E402
from oracle import ref_model  # noqa: E402


def import_reference():
  shim = types.ModuleType('ml_collections')
  shim.ConfigDict = ConfigDict
  sys.modules['ml_collections'] = shim

  ns = {'torch': torch, 'F': F}
  for fname, wanted in (('op/upfirdn2d.py', 'upfirdn2d_native'),):
    tree = ast.parse(open(os.path.join(REF, fname)).read())
    for node in tree.body:
      if isinstance(node, ast.FunctionDef) and node.name == wanted:
        exec(compile(ast.Module([node], []), fname, 'exec'), ns)
  native = ns['upfirdn2d_native']

  def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    return native(input, kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])

  def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    # CPU branch of op/fused_act.py:87-94 (slope hard-coded to 0.2 there)
    rest = [1] * (input.ndim - bias.ndim - 1)
    return F.leaky_relu(input + bias.view(1, bias.shape[0], *rest), negative_slope=0.2) * scale

  op = types.ModuleType('op')
  op.upfirdn2d = upfirdn2d
  op.upfirdn2d_native = native
  op.fused_leaky_relu = fused_leaky_relu
  op.FusedLeakyReLU = None
  sys.modules['op'] = op
  sys.path.insert(0, REF)
  import sde_lib, losses, sampling  # noqa: E401
  from models import ncsnpp, utils as mutils, ema
  return types.SimpleNamespace(sde_lib=sde_lib, losses=losses, sampling=sampling, ncsnpp=ncsnpp,
                               mutils=mutils, ema=ema, op=op)


def ref_config(path):
  import importlib
  mod = importlib.import_module('configs.' + path.replace('/', '.'))
  cfg = mod.get_config()
  cfg.device = torch.device('cpu')
  return cfg


def reduced(cfg, **model_over):
  for k, v in model_over.items():
    setattr(cfg.model, k, v)
  return cfg


def build_ref_model(R, cfg, seed, rezero=True):
  sde = R.sde_lib.get_sde(cfg, None)
  model = R.ncsnpp.NCSNpp(cfg, sde)
  sd = ref_model.make_state_dict(cfg, seed=seed, rezero=rezero)
  model.load_state_dict(sd, strict=True)
  return model, sde, sd


def sample_idx(numel, k=6):
  return np.unique(np.linspace(0, numel - 1, k).astype(np.int64))


# ----------------------------------------------------------------------------- fixtures
def golden_configs(R):
  out = {}
  for path in ('vp/CIFAR10/ddpmpp_nll_st', 'vp/IMAGENET32/ddpmpp_nll', 've/CELEBA/uncsnpp_st',
               've/celebahq/uncsnpp_st', 'vp/CIFAR10/ddpmpp_fid_st_deepest'):
    d = ref_config(path).to_dict()
    d.pop('device', None)
    d.get('data', {}).pop('tfrecords_path', None)
    out[path] = d
  with open(os.path.join(HERE, 'configs_golden.json'), 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True, default=list)


def golden_unet_cifar(R):
  cfg = ref_config('vp/CIFAR10/ddpmpp_nll_st')
  model, sde, sd = build_ref_model(R, cfg, seed=1)
  model.eval()
  g = torch.Generator().manual_seed(7)
  x = torch.randn(2, 3, 32, 32, generator=g)
  labels = torch.tensor([0.3, 0.9]) * 999
  wout = torch.randn(2, 3, 32, 32, generator=g)
  acts = {}
  hooks = []
  for i, mod in enumerate(model.all_modules):
    hooks.append(mod.register_forward_hook(lambda m, a, o, i=i: acts.__setitem__(i, o.detach())))
  out = model(x, labels)
  (out * wout).sum().backward()
  for h in hooks:
    h.remove()
  n_mod = len(model.all_modules)
  act_stats = np.zeros((n_mod, 2), np.float64)
  act_samples = np.zeros((n_mod, 6), np.float32)
  for i in range(n_mod):
    a = acts[i].double().reshape(-1)
    act_stats[i] = (a.mean().item(), a.std().item())
    idx = sample_idx(a.numel())
    act_samples[i, :len(idx)] = a[idx].float().numpy()
  names = [k for k, p in model.named_parameters()]
  gnorm = np.array([p.grad.double().norm().item() for _, p in model.named_parameters()])
  gsamp = np.zeros((len(names), 6), np.float32)
  for j, (_, p) in enumerate(model.named_parameters()):
    idx = sample_idx(p.numel())
    gsamp[j, :len(idx)] = p.grad.reshape(-1)[idx].numpy()
  np.savez_compressed(os.path.join(HERE, 'unet_cifar_golden.npz'),
                      x=x.numpy(), labels=labels.numpy(), wout=wout.numpy(), out=out.detach().numpy(),
                      act_stats=act_stats, act_samples=act_samples, param_names=np.array(names),
                      grad_norms=gnorm, grad_samples=gsamp, seed=1,
                      n_params=sum(p.numel() for p in model.parameters()))
  # the reference's own initialisation statistics (pins make_state_dict(rezero=False) in distribution)
  torch.manual_seed(3)
  fresh = R.ncsnpp.NCSNpp(cfg, sde)
  init_std = np.array([p.double().std().item() if p.numel() > 1 else 0. for _, p in fresh.named_parameters()])
  init_shapes = [list(p.shape) for _, p in fresh.named_parameters()]
  with open(os.path.join(HERE, 'init_golden.json'), 'w') as f:
    json.dump({'names': names, 'std': init_std.tolist(), 'shapes': init_shapes}, f)


def golden_train(R):
  base = ref_config('vp/CIFAR10/ddpmpp_nll_st')
  out = {}
  B, steps = 4, 10
  for tag, warmup in (('w5000', 5000), ('w0', 0)):
    cfg = ref_config('vp/CIFAR10/ddpmpp_nll_st')
    cfg.model.dropout = 0.
    cfg.optim.warmup = warmup
    model, sde, sd = build_ref_model(R, cfg, seed=2)
    wrapped = torch.nn.DataParallel(model)      # same wrapper create_model applies (CPU: pass-through)
    optimizer = R.losses.get_optimizer(cfg, wrapped.parameters())
    ema = R.ema.ExponentialMovingAverage(wrapped.parameters(), decay=cfg.model.ema_rate)
    state = dict(optimizer=optimizer, model=wrapped, ema=ema, step=0)
    step_fn = R.losses.get_step_fn(cfg, sde, train=True, optimize_fn=R.losses.optimization_manager(cfg))
    g = torch.Generator().manual_seed(1234)
    batch = torch.rand(B, 3, 32, 32, generator=g) * 2. - 1.
    losses, Us, us, zs = [], [], [], []
    for s in range(steps):
      np.random.seed(100 + s)
      torch.manual_seed(200 + s)
      # replay of the draws the reference is about to make, in its order
      Us.append(np.random.rand())
      us.append(torch.rand(B).numpy())
      zs.append(torch.randn(B, 3, 32, 32).numpy())
      np.random.seed(100 + s)
      torch.manual_seed(200 + s)
      losses.append(step_fn(state, batch).numpy())
    names = [k for k, _ in model.named_parameters()]
    probe = [names.index(n) for n in ('all_modules.2.weight', 'all_modules.3.Conv_0.weight',
                                      'all_modules.9.NIN_0.W', 'all_modules.27.GroupNorm_0.weight',
                                      'all_modules.54.weight', 'all_modules.1.bias')]
    params = list(model.parameters())
    pnorm = np.array([params[i].double().norm().item() for i in probe])
    psamp = np.stack([params[i].detach().reshape(-1)[:6].numpy() for i in probe])
    enorm = np.array([ema.shadow_params[i].double().norm().item() for i in probe])
    esamp = np.stack([ema.shadow_params[i].reshape(-1)[:6].numpy() for i in probe])
    out.update({f'{tag}_losses': np.stack(losses), f'{tag}_pnorm': pnorm, f'{tag}_psamp': psamp,
                f'{tag}_enorm': enorm, f'{tag}_esamp': esamp})
    out.update(batch=batch.numpy(), U=np.array(Us), u=np.stack(us), z=np.stack(zs).astype(np.float32),
               probe_names=np.array([names[i] for i in probe]))
  np.savez_compressed(os.path.join(HERE, 'train_golden.npz'), seed=2, **out)


def _trace_sampler(R, cfg, sde, model, shape, eps, seed):
  """Run the reference pc_sampler, recording the state after every predictor update."""
  trace = []
  orig = R.sampling.shared_predictor_update_fn

  def spy(*a, **k):
    x, x_mean = orig(*a, **k)
    trace.append(x.clone())
    return x, x_mean

  R.sampling.shared_predictor_update_fn = spy
  try:
    fn = R.sampling.get_sampling_fn(cfg, sde, shape, lambda v: v, eps)
    torch.manual_seed(seed)
    x, nfe = fn(model)
  finally:
    R.sampling.shared_predictor_update_fn = orig
  return x, nfe, trace


def golden_sampler(R):
  out = {}
  # VP, Euler-Maruyama, no corrector: C1's 8-step gate (model keeps num_scales=1000)
  cfg = ref_config('vp/CIFAR10/ddpmpp_nll_st')
  cfg.sampling.method = 'pc'
  model, _, _ = build_ref_model(R, cfg, seed=1)
  model.eval()
  sde8 = R.sde_lib.VPSDE(truncation_time=cfg.training.truncation_time, beta_min=cfg.model.beta_min,
                         beta_max=cfg.model.beta_max, N=8)
  shape = (2, 3, 32, 32)
  torch.manual_seed(42)
  x_T = torch.randn(*shape)
  zs = [torch.randn(*shape) for _ in range(8)]
  x, nfe, trace = _trace_sampler(R, cfg, sde8, model, shape, cfg.sampling.truncation_time, 42)
  out.update(vp_xT=x_T.numpy(), vp_z=np.stack([z.numpy() for z in zs]), vp_x=x.numpy(), vp_nfe=nfe,
             vp_trace=np.stack([t.numpy() for t in trace]), vp_eps=cfg.sampling.truncation_time)

  # VE, reverse diffusion + Langevin on a reduced C5-shaped network, 4 steps
  cfg = reduced_c5(ref_config('ve/celebahq/uncsnpp_st'))
  model, _, _ = build_ref_model(R, cfg, seed=5)
  model.eval()
  sde4 = R.sde_lib.VESDE(sigma_min=cfg.model.sigma_min, sigma_max=cfg.model.sigma_max, N=4)
  shape = (2, 3, 32, 32)
  torch.manual_seed(43)
  x_T = torch.randn(*shape) * cfg.model.sigma_max
  zs = [torch.randn(*shape) for _ in range(8)]     # per step: corrector noise, predictor noise
  x, nfe, trace = _trace_sampler(R, cfg, sde4, model, shape, cfg.sampling.truncation_time, 43)
  out.update(ve_xT=x_T.numpy(), ve_z=np.stack([z.numpy() for z in zs]), ve_x=x.numpy(), ve_nfe=nfe,
             ve_trace=np.stack([t.numpy() for t in trace]), ve_eps=cfg.sampling.truncation_time)
  np.savez_compressed(os.path.join(HERE, 'sampler_golden.npz'), **out)


def reduced_c5(cfg):
  cfg.data.image_size = 32
  return reduced(cfg, nf=32, ch_mult=(1, 1, 2, 2), num_res_blocks=1, num_scales=2000)


def reduced_c3(cfg):
  cfg.data.image_size = 32
  return reduced(cfg, nf=32, ch_mult=(1, 2, 2), num_res_blocks=1)


def golden_variants(R):
  """Scores + one training loss on reduced-width copies of C3 (RVE, FIR, residual input pyramid)
  and C5 (VE, FIR, input_skip/output_skip) so the oracle is pinned on those code paths too."""
  out = {}
  for tag, path, shrink, seed in (('c3', 've/CELEBA/uncsnpp_st', reduced_c3, 3),
                                  ('c5', 've/celebahq/uncsnpp_st', reduced_c5, 5)):
    cfg = shrink(ref_config(path))
    cfg.model.dropout = 0.
    model, sde, _ = build_ref_model(R, cfg, seed=seed)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2, 3, 32, 32, generator=g)
    sig = torch.tensor([0.5, 20.])
    model.eval()
    score = model(x, sig).detach()
    # one loss evaluation with replayed draws
    loss_fn = R.losses.get_sde_loss_fn(cfg, sde, train=True)
    np.random.seed(5)
    t_min = sde.get_t_min(cfg)
    torch.manual_seed(9)
    u = torch.rand(2)
    z = torch.randn(2, 3, 32, 32)
    torch.manual_seed(9)
    losses = loss_fn(model, x, importance_sampling=cfg.training.importance_sampling, t_min=t_min)
    torch.mean(losses).backward()
    names = [k for k, _ in model.named_parameters()]
    gnorm = np.array([0. if p.grad is None else p.grad.double().norm().item() for p in model.parameters()])
    out.update({f'{tag}_x': x.numpy(), f'{tag}_sig': sig.numpy(), f'{tag}_score': score.numpy(),
                f'{tag}_tmin': t_min, f'{tag}_u': u.numpy(), f'{tag}_z': z.numpy(),
                f'{tag}_losses': losses.detach().numpy(), f'{tag}_gnorm': gnorm,
                f'{tag}_names': np.array(names), f'{tag}_seed': seed})
  np.savez_compressed(os.path.join(HERE, 'variants_golden.npz'), **out)


def reduced_deepest(cfg):
  """configs/vp/CIFAR10/ddpmpp_fid_st_deepest.py at reduced width: keeps ch_mult (1,1,1), FIR resampling, attention at
  16x16, the `lsgm` embedding and training.mixed / ddpm_weight 100; dropout off and warm-up 0 for the trajectory."""
  cfg.model.dropout = 0.
  cfg.optim.warmup = 0
  return reduced(cfg, nf=32, num_res_blocks=2, embedding_dim=16)


def golden_deepest(R):
  """the README's headline FID config (nf=512, ch_mult (1,1,1), 8 res-blocks, FIR, lsgm embedding, mixed
  IS + uniform-time loss).  Score + 3 optimizer steps of step_fn_mixed with replayed draws."""
  cfg = reduced_deepest(ref_config('vp/CIFAR10/ddpmpp_fid_st_deepest'))
  seed, B, steps = 7, 4, 3
  model, sde, _ = build_ref_model(R, cfg, seed=seed)
  g = torch.Generator().manual_seed(21)
  x = torch.rand(2, 3, 32, 32, generator=g) * 2. - 1.
  labels = torch.tensor([3.7, 911.2])
  model.eval()
  out = model(x, labels).detach()
  wrapped = torch.nn.DataParallel(model)
  optimizer = R.losses.get_optimizer(cfg, wrapped.parameters())
  ema = R.ema.ExponentialMovingAverage(wrapped.parameters(), decay=cfg.model.ema_rate)
  state = dict(optimizer=optimizer, model=wrapped, ema=ema, step=0)
  step_fn = R.losses.get_step_fn(cfg, sde, train=True, optimize_fn=R.losses.optimization_manager(cfg))
  assert step_fn.__name__ == 'step_fn_mixed'
  batch = torch.rand(B, 3, 32, 32, generator=g) * 2. - 1.
  h = B // 2
  losses, Us, us, zs = [], [], [], []
  for s in range(steps):
    np.random.seed(300 + s)
    torch.manual_seed(400 + s)
    # replay of the draws step_fn_mixed is about to make: t_min (NumPy), then per half u (rand) and z (randn)
    Us.append(np.random.rand())
    u_is, z_is = torch.rand(h), torch.randn(h, 3, 32, 32)
    u_dd, z_dd = torch.rand(h), torch.randn(h, 3, 32, 32)
    us.append(torch.cat([u_is, u_dd]).numpy())
    zs.append(torch.cat([z_is, z_dd]).numpy())
    np.random.seed(300 + s)
    torch.manual_seed(400 + s)
    losses.append(step_fn(state, batch).numpy())
  names = [k for k, _ in model.named_parameters()]
  params = list(model.parameters())
  pnorm = np.array([p.double().norm().item() for p in params])
  enorm = np.array([e.double().norm().item() for e in ema.shadow_params])
  np.savez_compressed(os.path.join(HERE, 'deepest_golden.npz'), seed=seed, x=x.numpy(), labels=labels.numpy(),
                      out=out.numpy(), batch=batch.numpy(), U=np.array(Us), u=np.stack(us),
                      z=np.stack(zs).astype(np.float32), losses=np.stack(losses), pnorm=pnorm, enorm=enorm,
                      names=np.array(names))


def reduced_cifar(cfg):
  return reduced(cfg, nf=32, ch_mult=(1, 2), num_res_blocks=1)


def golden_likelihood(R):
  """the reference's likelihood.py on a reduced-width CIFAR DDPM++ (VP): Hutchinson divergence of the
  probability-flow drift, one bits/dim evaluation (RK45, 'correct' mode) and one NELBO sample, draws recorded."""
  import likelihood as L
  cfg = reduced_cifar(ref_config('vp/CIFAR10/ddpmpp_nll_st'))
  seed, B = 13, 2
  model, sde, _ = build_ref_model(R, cfg, seed=seed)
  model.eval()
  inv = lambda v: (v + 1.) / 2.          # datasets.get_data_inverse_scaler for centered data (datasets.py:65-71)
  g = torch.Generator().manual_seed(31)
  data = (torch.randint(0, 256, (B, 3, 32, 32), generator=g).float() / 255.) * 2. - 1.
  out = dict(seed=seed, data=data.numpy())
  # --- drift and divergence at fixed (x, t, eps)
  t = torch.tensor([0.3, 0.7])
  x = torch.randn(B, 3, 32, 32, generator=g)
  eps_h = torch.randint(0, 2, (B, 3, 32, 32), generator=g).float() * 2 - 1.
  score_fn = R.mutils.get_score_fn(cfg, sde, model, train=False, continuous=True)
  rsde = sde.reverse(score_fn, probability_flow=cfg.eval.probability_flow, lambda_=cfg.eval.lambda_)
  drift_fn = lambda xx, tt: rsde.sde(xx, tt)[0]
  with torch.no_grad():
    drift = drift_fn(x, t)
    div = L.get_div_fn(drift_fn)(x.clone(), t, eps_h)
  out.update(t=t.numpy(), x=x.numpy(), eps_h=eps_h.numpy(), drift=drift.numpy(), div=div.numpy())
  # --- bits/dim
  rtol = atol = 1e-3
  lik = L.get_likelihood_fn(cfg, sde, inv, rtol=rtol, atol=atol)
  torch.manual_seed(77)
  epsilon = torch.randint_like(data, low=0, high=2).float() * 2 - 1.
  z = torch.randn_like(data)
  z_res = torch.randn_like(data)
  torch.manual_seed(77)
  bpd, latent, nfe = lik(model, data, eps=1e-3)
  out.update(lik_rtol=rtol, lik_eps=1e-3, lik_epsilon=epsilon.numpy(), lik_z=z.numpy(), lik_z_res=z_res.numpy(),
             lik_bpd=bpd.numpy(), lik_latent=latent.numpy(), lik_nfe=nfe)
  # --- NELBO sample
  elbo = L.get_elbo_fn(cfg, sde, inv)
  torch.manual_seed(78)
  u = torch.rand(B)
  z = torch.randn_like(data)
  epsilon = torch.randint_like(data, low=0, high=2).float() * 2 - 1.
  lp_z = torch.randn_like(data)
  z_res = torch.randn_like(data)
  torch.manual_seed(78)
  nelbo, resid = elbo(model, data, eps=1e-3)
  out.update(elbo_eps=1e-3, elbo_u=u.numpy(), elbo_z=z.numpy(), elbo_epsilon=epsilon.numpy(), elbo_lp_z=lp_z.numpy(),
             elbo_z_res=z_res.numpy(), elbo_nelbo=nelbo.detach().numpy(), elbo_resid=resid.detach().numpy())
  np.savez_compressed(os.path.join(HERE, 'likelihood_golden.npz'), **out)


def golden_lossbranches(R):
  """The likelihood-weighted loss branch (reference losses.py:125-129: (score + z/std)^2 * g^2, uniform times), which no
  shipped config reaches because importance sampling is tested first (:122).  VP (reduced CIFAR net) and VE (reduced C5)."""
  out = {}
  for tag, path, shrink, seed in (('vp', 'vp/CIFAR10/ddpmpp_nll_st', reduced_cifar, 17),
                                  ('ve', 've/celebahq/uncsnpp_st', reduced_c5, 18)):
    cfg = shrink(ref_config(path))
    cfg.model.dropout = 0.
    cfg.training.likelihood_weighting, cfg.training.importance_sampling = True, False
    model, sde, _ = build_ref_model(R, cfg, seed=seed)
    g = torch.Generator().manual_seed(41)
    x = torch.rand(2, 3, 32, 32, generator=g)
    if cfg.data.centered:
      x = x * 2. - 1.
    loss_fn = R.losses.get_sde_loss_fn(cfg, sde, train=True)
    t_min = 1e-3
    torch.manual_seed(9)
    u, z = torch.rand(2), torch.randn(2, 3, 32, 32)
    torch.manual_seed(9)
    losses = loss_fn(model, x, importance_sampling=False, t_min=t_min)
    torch.mean(losses).backward()
    gnorm = np.array([0. if p.grad is None else p.grad.double().norm().item() for p in model.parameters()])
    out.update({f'{tag}_x': x.numpy(), f'{tag}_u': u.numpy(), f'{tag}_z': z.numpy(), f'{tag}_tmin': t_min,
                f'{tag}_losses': losses.detach().numpy(), f'{tag}_gnorm': gnorm, f'{tag}_seed': seed})
  np.savez_compressed(os.path.join(HERE, 'lossbranch_golden.npz'), **out)


def golden_recon(R):
  """The reconstruction (decoder) term of the loss (reference losses.py:134-164), which no shipped config switches on:
  reduced CIFAR VP net, importance-sampled DSM loss + decoder term, for 'uniform' data (Gaussian cross-entropy minus the
  perturbation entropy; both decoder variances) and 'lossless' data (discretised Gaussian likelihood).  The three torch
  draws of loss_fn (u, z, second z) are replayed from the seed."""
  out = {}
  for tag, deq, variance, reduce_mean in (('uni_sf', 'uniform', 'scoreflow', False), ('uni_ddpm', 'uniform', 'ddpm', True),
                                          ('lossless', 'lossless', 'scoreflow', False)):
    cfg = reduced_cifar(ref_config('vp/CIFAR10/ddpmpp_nll_st'))
    cfg.model.dropout = 0.
    cfg.training.reconstruction_loss = True
    cfg.training.reduce_mean = reduce_mean
    cfg.data.dequantization = deq
    model, sde, _ = build_ref_model(R, cfg, seed=23)
    g = torch.Generator().manual_seed(43)
    x = torch.rand(2, 3, 32, 32, generator=g)
    if deq == 'lossless':
      x = torch.round(x * 255.) / 255.
      x[0, 0, 0, :4] = torch.tensor([0., 1., 0., 1.])        # exercise the open-ended first / last bins
    x = x * 2. - 1.
    loss_fn = R.losses.get_sde_loss_fn(cfg, sde, train=True, variance=variance)
    t_min = 2e-3
    torch.manual_seed(10)
    u, z, z2 = torch.rand(2), torch.randn(2, 3, 32, 32), torch.randn(2, 3, 32, 32)
    torch.manual_seed(10)
    losses = loss_fn(model, x, importance_sampling=True, t_min=t_min)
    torch.mean(losses).backward()
    gnorm = np.array([0. if p.grad is None else p.grad.double().norm().item() for p in model.parameters()])
    out.update({f'{tag}_x': x.numpy(), f'{tag}_u': u.numpy(), f'{tag}_z': z.numpy(), f'{tag}_z2': z2.numpy(),
                f'{tag}_tmin': t_min, f'{tag}_losses': losses.detach().numpy(), f'{tag}_gnorm': gnorm, f'{tag}_seed': 23})
  np.savez_compressed(os.path.join(HERE, 'recon_golden.npz'), **out)


def golden_sde_reverse(R):
  """a6-a9 on every SDE class incl. subVPSDE: sde / marginal_prob / prior_logp / discretize and the reverse-time SDE
  (lambda 1), a lambda 0.5 interpolation and the probability-flow ODE (lambda 0) built by SDE.reverse with a fixed
  analytic score (reference sde_lib.py:55-119)."""
  out = {}
  g = torch.Generator().manual_seed(8)
  x = torch.randn(3, 3, 4, 4, generator=g)
  t = torch.tensor([0.05, 0.5, 0.95])
  score = lambda xx, tt: -0.3 * xx + 0.1 * tt[:, None, None, None]
  sdes = dict(vp=R.sde_lib.VPSDE(truncation_time=1e-5, beta_min=0.1, beta_max=20., N=1000),
              subvp=R.sde_lib.subVPSDE(beta_min=0.1, beta_max=20., N=1000),
              ve=R.sde_lib.VESDE(sigma_min=0.01, sigma_max=50., N=1000))
  for tag, sde in sdes.items():
    f, gg = sde.sde(x, t)
    mean, std = sde.marginal_prob(x, t)
    fd, Gd = sde.discretize(x, t)
    out.update({f'{tag}_f': f.numpy(), f'{tag}_g': gg.numpy(), f'{tag}_mean': mean.numpy(), f'{tag}_std': std.numpy(),
                f'{tag}_logp': sde.prior_logp(x).numpy(), f'{tag}_fd': fd.numpy(), f'{tag}_Gd': Gd.numpy(),
                f'{tag}_T': float(sde.T)})
    for name, pf, lam in (('rsde', False, 1.), ('mix', False, 0.5), ('ode', True, 0.)):
      r = sde.reverse(score, probability_flow=pf, lambda_=lam)
      rf, rg = r.sde(x, t)
      rfd, rGd = r.discretize(x, t)
      out.update({f'{tag}_{name}_f': rf.numpy(), f'{tag}_{name}_g': rg.numpy(), f'{tag}_{name}_fd': rfd.numpy(),
                  f'{tag}_{name}_Gd': rGd.numpy()})
  out.update(x=x.numpy(), t=t.numpy())
  np.savez_compressed(os.path.join(HERE, 'sde_reverse_golden.npz'), **out)


class _LabelEcho(torch.nn.Module):
  """Stand-in network: returns its conditioning labels broadcast over x, so that a score function's label and scale
  conventions can be read off its output."""

  def forward(self, x, labels):
    return torch.zeros_like(x) + labels.float()[:, None, None, None]


def golden_score_fn(R):
  """models/utils.py:128-190 (a10): labels and output scaling of get_score_fn for VP continuous (plain and
  `unbounded_parametrization`), VP discrete, VP without ddpm_score, VE continuous and VE discrete."""
  out = {}
  net = _LabelEcho()
  x = torch.zeros(3, 3, 4, 4)
  t = torch.tensor([0.05, 0.5, 0.95])
  cfg = ref_config('vp/CIFAR10/ddpmpp_nll_st')
  vp = R.sde_lib.get_sde(cfg, None)
  out['vp_cont'] = R.mutils.get_score_fn(cfg, vp, net, train=False, continuous=True)(x, t).numpy()
  out['vp_disc'] = R.mutils.get_score_fn(cfg, vp, net, train=False, continuous=False)(x, t).numpy()
  cfg.training.unbounded_parametrization = True
  out['vp_unbounded'] = R.mutils.get_score_fn(cfg, vp, net, train=False, continuous=True)(x, t).numpy()
  cfg.training.unbounded_parametrization = False
  cfg.training.ddpm_score = False
  out['vp_raw'] = R.mutils.get_score_fn(cfg, vp, net, train=False, continuous=True)(x, t).numpy()
  cfg5 = ref_config('ve/celebahq/uncsnpp_st')
  ve = R.sde_lib.get_sde(cfg5, None)
  out['ve_cont'] = R.mutils.get_score_fn(cfg5, ve, net, train=False, continuous=True)(x, t).numpy()
  out['ve_disc'] = R.mutils.get_score_fn(cfg5, ve, net, train=False, continuous=False)(x, t.clone()).numpy()
  out['t'] = t.numpy()
  np.savez_compressed(os.path.join(HERE, 'score_fn_golden.npz'), **out)


def golden_predictors(R):
  """Every registered predictor / corrector of the reference (sampling.py:185-340) on an analytic score function:
  outputs (x, x_mean) with the noise draws replayed, VP and VE, 50-step schedules, two inner corrector steps."""
  out = {}
  cfg = ref_config('vp/CIFAR10/ddpmpp_nll_st')
  score = lambda xx, tt: -0.3 * xx + 0.1 * tt[:, None, None, None]
  gen = torch.Generator().manual_seed(3)
  x = torch.randn(3, 3, 8, 8, generator=gen)
  t = torch.tensor([0.9, 0.5, 0.11])
  out.update(x=x.numpy(), t=t.numpy())
  for kind, sde in (('vp', R.sde_lib.VPSDE(truncation_time=1e-5, beta_min=0.1, beta_max=20., N=50)),
                    ('ve', R.sde_lib.VESDE(sigma_min=0.01, sigma_max=50., N=50))):
    torch.manual_seed(77)
    z, z2 = torch.randn_like(x), torch.randn_like(x)
    out.update({f'{kind}_z': z.numpy(), f'{kind}_z2': z2.numpy()})
    for name in ('euler_maruyama', 'reverse_diffusion', 'ancestral_sampling'):
      p = R.sampling.get_predictor(name)(cfg, sde, score, probability_flow=False)
      torch.manual_seed(77)
      a, b = p.update_fn(x.clone(), t.clone())
      out.update({f'{kind}_{name}_x': a.numpy(), f'{kind}_{name}_mean': b.numpy()})
    for name in ('langevin', 'ald'):
      c = R.sampling.get_corrector(name)(sde, score, 0.16, 2)
      torch.manual_seed(77)
      a, b = c.update_fn(x.clone(), t.clone())
      out.update({f'{kind}_{name}_x': a.numpy(), f'{kind}_{name}_mean': b.numpy()})
  np.savez_compressed(os.path.join(HERE, 'predictors_golden.npz'), **out)


def golden_sde(R):
  out = {}
  u = torch.linspace(0.01, 0.99, 7)
  x = torch.zeros(7, 1, 1, 1)
  cfg = ref_config('vp/CIFAR10/ddpmpp_nll_st')
  vp = R.sde_lib.get_sde(cfg, None)
  np.random.seed(0)
  tmins = np.array([vp.get_t_min(cfg) for _ in range(5)])
  np.random.seed(0)
  out.update(vp_tmin_U=np.array([np.random.rand() for _ in range(5)]), vp_tmin=tmins)
  cfg.training.k = 1.2
  np.random.seed(0)
  out['vp_tmin_k12'] = np.array([vp.get_t_min(cfg) for _ in range(5)])
  t_min = 3e-4
  torch.manual_seed(0)
  t_is, Z = vp.get_diffusion_time(cfg, 7, 'cpu', t_min, importance_sampling=True)
  torch.manual_seed(0)
  out.update(vp_u=torch.rand(7).numpy(), vp_t_is=t_is.numpy(), vp_Z=Z.numpy(), vp_tmin_used=t_min)
  mean, std = vp.marginal_prob(torch.ones(7, 1, 1, 1), t_is)
  drift, g = vp.sde(torch.ones(7, 1, 1, 1), t_is)
  out.update(vp_mean=mean.reshape(-1).numpy(), vp_std=std.numpy(), vp_g=g.numpy())

  cfg5 = ref_config('ve/celebahq/uncsnpp_st')
  ve = R.sde_lib.get_sde(cfg5, None)
  t = torch.linspace(1e-5, 1., 7)
  out.update(ve_t=t.numpy(), ve_std=ve.marginal_prob(x, t)[1].numpy(), ve_g=ve.sde(x, t)[1].numpy(),
             ve_G=ve.discretize(x, t)[1].numpy(), ve_tmin=ve.get_t_min(cfg5), ve_eps=ve.eps)
  cfg3 = ref_config('ve/CELEBA/uncsnpp_st')
  rve = R.sde_lib.get_sde(cfg3, None)
  torch.manual_seed(1)
  t_r, _ = rve.get_diffusion_time(cfg3, 7, 'cpu', rve.get_t_min(cfg3))
  torch.manual_seed(1)
  out.update(rve_u=torch.rand(7).numpy(), rve_t=t_r.numpy(), rve_std=rve.marginal_prob(x, t_r)[1].numpy(),
             rve_g=rve.sde(x, t_r)[1].numpy(), rve_tmin=rve.get_t_min(cfg3))
  np.savez_compressed(os.path.join(HERE, 'sde_golden.npz'), **out)


def golden_ops(R):
  out = {}
  g = torch.Generator().manual_seed(21)
  k1 = np.asarray([1., 3., 3., 1.], dtype=np.float32)
  k = np.outer(k1, k1)
  k /= k.sum()
  cases = {'up2': (k * 4, 2, 1, (2, 1)), 'down2': (k, 1, 2, (1, 1)), 'pre': (k, 1, 1, (2, 2)),
           'crop': (k, 1, 1, (-1, 0)), 'up3_3x3': (np.outer([1., 2., 1.], [1., 2., 1.]).astype(np.float32) / 16, 3, 2, (1, 2))}
  for name, (kk, up, down, pad) in cases.items():
    x = torch.randn(2, 3, 8, 8, generator=g, requires_grad=True)
    y = R.op.upfirdn2d(x, torch.tensor(kk), up=up, down=down, pad=pad)
    gy = torch.randn(y.shape, generator=g)
    (y * gy).sum().backward()
    out.update({f'{name}_x': x.detach().numpy(), f'{name}_k': kk, f'{name}_y': y.detach().numpy(),
                f'{name}_gy': gy.numpy(), f'{name}_gx': x.grad.numpy(),
                f'{name}_args': np.array([up, down, pad[0], pad[1]])})
  x = torch.randn(2, 5, 4, 4, generator=g)
  b = torch.randn(5, generator=g)
  out.update(lrelu_x=x.numpy(), lrelu_b=b.numpy(), lrelu_y=R.op.fused_leaky_relu(x, b).numpy())
  np.savez_compressed(os.path.join(HERE, 'ops_golden.npz'), **out)


def reduced_ckpt(cfg):
  cfg.model.dropout = 0.
  cfg.optim.warmup = 0
  return reduced(cfg, nf=16, ch_mult=(1, 2), num_res_blocks=1)


def golden_checkpoint(R):
  """a checkpoint WRITTEN BY THE REFERENCE - its NCSNpp (DataParallel-wrapped, `module.` keys), its
  torch.optim.Adam and its ExponentialMovingAverage after two CPU optimizer steps, saved by the reference's own
  utils.save_checkpoint (utils.py:29-36, imported behind the tensorflow.io.gfile stub) - plus the losses, parameter and
  EMA norms of the reference's THIRD step, which the GPU test must reproduce after restoring the file."""
  sys.path.insert(0, ROOT)
  from baseline import ref_env
  ref_env.install_shims(REF, drivers=True)
  import utils as ref_utils
  cfg = reduced_ckpt(ref_config('vp/CIFAR10/ddpmpp_nll_st'))
  seed, B = 23, 4
  model, sde, _ = build_ref_model(R, cfg, seed=seed)
  wrapped = torch.nn.DataParallel(model)
  optimizer = R.losses.get_optimizer(cfg, wrapped.parameters())
  ema = R.ema.ExponentialMovingAverage(wrapped.parameters(), decay=cfg.model.ema_rate)
  state = dict(optimizer=optimizer, model=wrapped, ema=ema, step=0)
  step_fn = R.losses.get_step_fn(cfg, sde, train=True, optimize_fn=R.losses.optimization_manager(cfg))
  g = torch.Generator().manual_seed(55)
  batch = torch.rand(B, 3, 32, 32, generator=g) * 2. - 1.
  losses = []
  for s in range(3):
    if s == 2:
      ref_utils.save_checkpoint(cfg, os.path.join(HERE, 'ref_checkpoint.pth'), state)
      np.random.seed(500 + s)
      torch.manual_seed(600 + s)
      U, u, z = np.random.rand(), torch.rand(B).numpy(), torch.randn(B, 3, 32, 32).numpy()
    np.random.seed(500 + s)
    torch.manual_seed(600 + s)
    losses.append(step_fn(state, batch).numpy())
  names = [k for k, _ in model.named_parameters()]
  pnorm = np.array([p.double().norm().item() for p in model.parameters()])
  enorm = np.array([e.double().norm().item() for e in ema.shadow_params])
  np.savez_compressed(os.path.join(HERE, 'checkpoint_golden.npz'), seed=seed, batch=batch.numpy(), U=U, u=u,
                      z=z.astype(np.float32), losses=np.stack(losses), pnorm=pnorm, enorm=enorm, names=np.array(names),
                      step_after=state['step'])


def golden_fullwidth(R):
  """FULL-width C3 (RVE, 64x64, FIR, residual input pyramid) and C5 (VE, 256x256, 7 levels, input_skip / output_skip)
  on ONE image: score and training loss.  Inputs and draws are regenerated from seeds by the tests (torch CPU
  generator), weights from oracle.ref_model.make_state_dict; the fixture keeps a strided sample of the score."""
  out = {}
  for tag, path, seed in (('c3', 've/CELEBA/uncsnpp_st', 31), ('c5', 've/celebahq/uncsnpp_st', 32)):
    cfg = ref_config(path)
    cfg.model.dropout = 0.
    Rz = cfg.data.image_size
    model, sde, _ = build_ref_model(R, cfg, seed=seed)
    model.eval()
    g = torch.Generator().manual_seed(seed + 100)
    x = torch.rand(1, 3, Rz, Rz, generator=g)
    sig = torch.tensor([1.7])
    with torch.no_grad():
      score = model(x, sig)
    loss_fn = R.losses.get_sde_loss_fn(cfg, sde, train=True)
    t_min = sde.get_t_min(cfg)
    torch.manual_seed(seed + 200)
    u, z = torch.rand(1), torch.randn(1, 3, Rz, Rz)
    torch.manual_seed(seed + 200)
    with torch.no_grad():
      losses = loss_fn(model, x, importance_sampling=cfg.training.importance_sampling, t_min=t_min)
    flat = score.reshape(-1)
    idx = np.arange(0, flat.numel(), 97)
    out.update({f'{tag}_seed': seed, f'{tag}_sig': sig.numpy(), f'{tag}_idx': idx, f'{tag}_score_samples': flat[idx].numpy(),
                f'{tag}_score_norm': flat.double().norm().item(), f'{tag}_tmin': t_min, f'{tag}_u': u.numpy(),
                f'{tag}_z_checksum': z.double().sum().item(), f'{tag}_x_checksum': x.double().sum().item(),
                f'{tag}_losses': losses.numpy(), f'{tag}_n_params': sum(p.numel() for p in model.parameters())})
  np.savez_compressed(os.path.join(HERE, 'fullwidth_golden.npz'), **out)


def golden_traj100(R):
  """north_star: "score-matching loss trajectory matching the reference within tolerance" - 100 optimizer steps of the
  reference (fp32, CPU) on the full-size CIFAR-10 DDPM++ at batch 16, dropout 0, warm-up 0, lr 2e-4, with replayable
  draws: step s seeds NumPy with 100+s (t_min) and torch with 200+s (u = rand(B), z = randn(B,3,32,32))."""
  cfg = ref_config('vp/CIFAR10/ddpmpp_nll_st')
  cfg.model.dropout = 0.
  cfg.optim.warmup = 0
  seed, B, steps = 2, 16, 100
  model, sde, _ = build_ref_model(R, cfg, seed=seed)
  wrapped = torch.nn.DataParallel(model)
  optimizer = R.losses.get_optimizer(cfg, wrapped.parameters())
  ema = R.ema.ExponentialMovingAverage(wrapped.parameters(), decay=cfg.model.ema_rate)
  state = dict(optimizer=optimizer, model=wrapped, ema=ema, step=0)
  step_fn = R.losses.get_step_fn(cfg, sde, train=True, optimize_fn=R.losses.optimization_manager(cfg))
  g = torch.Generator().manual_seed(1234)
  batch = torch.rand(B, 3, 32, 32, generator=g) * 2. - 1.
  losses, Us = [], []
  for s in range(steps):
    np.random.seed(100 + s)
    Us.append(np.random.rand())
    np.random.seed(100 + s)
    torch.manual_seed(200 + s)
    losses.append(step_fn(state, batch).numpy())
    if s % 10 == 0:
      print('  traj100 step', s, float(losses[-1].mean()), flush=True)
  np.savez_compressed(os.path.join(HERE, 'traj100_golden.npz'), seed=seed, B=B, U=np.array(Us), losses=np.stack(losses),
                      batch_checksum=batch.double().sum().item())


def main(which):
  torch.set_num_threads(8)
  R = import_reference()
  jobs = dict(configs=golden_configs, ops=golden_ops, sde=golden_sde, unet=golden_unet_cifar,
              variants=golden_variants, sampler=golden_sampler, train=golden_train, deepest=golden_deepest,
              likelihood=golden_likelihood, lossbranches=golden_lossbranches, recon=golden_recon,
              sde_reverse=golden_sde_reverse, score_fn=golden_score_fn,
              predictors=golden_predictors, checkpoint=golden_checkpoint, fullwidth=golden_fullwidth,
              traj100=golden_traj100)
  for name in (which or jobs):
    print('golden:', name, flush=True)
    jobs[name](R)


if __name__ == '__main__':
  main(sys.argv[1:])
