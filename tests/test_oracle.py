"""Pins the oracle (oracle/) against fixtures produced by the untouched reference
(tests/golden/make_golden.py).  CPU only.

Tolerances: the reference itself moves by 2e-7 rel-L2 between thread counts;
the oracle evaluates the same fp32 math in a different op order (functional walk, einsum
forms), so forward quantities are held to 2e-5 rel-L2 and gradients to 1e-4.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import ref_model, ref_ops, ref_train
from soft_truncation_b200 import configs
from soft_truncation_b200.models import utils as mutils


def _reduced_c3():
  cfg = configs.celeba_uncsnpp_st()
  cfg.data.image_size = 32
  cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks = 32, (1, 2, 2), 1
  cfg.model.dropout = 0.
  return cfg


def _reduced_c5():
  cfg = configs.celebahq_uncsnpp_st()
  cfg.data.image_size = 32
  cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks = 32, (1, 1, 2, 2), 1
  return cfg


def _reduced_deepest():
  cfg = configs.cifar10_ddpmpp_fid_st_deepest()
  cfg.model.nf, cfg.model.num_res_blocks, cfg.model.embedding_dim = 32, 2, 16
  cfg.model.dropout = 0.
  cfg.optim.warmup = 0
  return cfg


def test_configs_match_reference():
  want = json.load(open(os.path.join(GOLDEN, 'configs_golden.json')))
  for path, ref in want.items():
    got = configs.get_config(path).to_dict()
    got.pop('device')
    got = json.loads(json.dumps(got, default=list))
    assert got == ref, path


def test_upfirdn2d_oracle_vs_reference(golden):
  g = golden('ops_golden.npz')
  for name in ('up2', 'down2', 'pre', 'crop', 'up3_3x3'):
    up, down, p0, p1 = [int(v) for v in g[f'{name}_args']]
    x = g[f'{name}_x']
    y = ref_ops.upfirdn2d_ref(x.reshape(-1, *x.shape[2:]), g[f'{name}_k'], (up, up), (down, down),
                              (p0, p1, p0, p1))
    want = g[f'{name}_y']
    assert y.shape == (want.shape[0] * want.shape[1],) + want.shape[2:]
    np.testing.assert_allclose(y.reshape(want.shape), want, rtol=1e-5, atol=1e-6)
    # backward = same op on grad_out with flipped kernel, up<->down swapped (op/upfirdn2d.py:101-116)
    kh, kw = g[f'{name}_k'].shape
    gup, gdown, gpad = ref_ops.upfirdn2d_backward_args(x.shape[2], x.shape[3], want.shape[2], want.shape[3],
                                                       kh, kw, (up, up), (down, down), (p0, p1, p0, p1))
    gy = g[f'{name}_gy']
    if kh % gup[1] or kw % gup[0]:
      # the reference kernel walks kernel_h / up_y taps (upfirdn2d_kernel.cu:182-203), i.e. it needs
      # the FIR size to be a multiple of the up factor; 3 taps with backward-up 2 is outside its domain
      continue
    gx = ref_ops.upfirdn2d_ref(gy.reshape(-1, *gy.shape[2:]), g[f'{name}_k'][::-1, ::-1].copy(), gup, gdown, gpad)
    np.testing.assert_allclose(gx.reshape(x.shape), g[f'{name}_gx'], rtol=1e-5, atol=1e-6)
    # the torch form used inside the oracle network agrees as well
    y2 = ref_model.upfirdn(torch.tensor(x), torch.tensor(g[f'{name}_k']), up=up, down=down, pad=(p0, p1))
    np.testing.assert_allclose(y2.numpy(), want, rtol=1e-5, atol=1e-6)


def test_fused_bias_act_oracle_vs_reference(golden):
  g = golden('ops_golden.npz')
  y = ref_ops.fused_bias_act_ref(g['lrelu_x'], g['lrelu_b'], None, 3, 0, 0.2, 2 ** 0.5)
  np.testing.assert_allclose(y, g['lrelu_y'], rtol=1e-6, atol=1e-7)
  # grad=1 gates on the sign of ref; grad=2 is identically zero (fused_bias_act_kernel.cu:36-45)
  x = g['lrelu_x']
  gi = ref_ops.fused_bias_act_ref(x, None, y, 3, 1, 0.2, 2 ** 0.5)
  np.testing.assert_allclose(gi, np.where(y > 0, x, 0.2 * x) * 2 ** 0.5, rtol=1e-6)
  assert not ref_ops.fused_bias_act_ref(x, None, y, 3, 2, 0.2, 1.).any()


def test_sde_oracle_vs_reference(golden):
  g = golden('sde_golden.npz')
  cfg = configs.cifar10_ddpmpp_nll_st()
  vp = ref_train.make_sde(cfg)
  got = np.array([vp.t_min_from_uniform(cfg, U) for U in g['vp_tmin_U']])
  np.testing.assert_allclose(got, g['vp_tmin'], rtol=1e-12)
  cfg.training.k = 1.2
  got = np.array([vp.t_min_from_uniform(cfg, U) for U in g['vp_tmin_U']])
  np.testing.assert_allclose(got, g['vp_tmin_k12'], rtol=1e-12)
  t, Z = vp.time_from_uniform(torch.tensor(g['vp_u']), float(g['vp_tmin_used']), True)
  np.testing.assert_array_equal(t.numpy(), g['vp_t_is'])
  np.testing.assert_array_equal(Z.numpy(), g['vp_Z'])
  np.testing.assert_array_equal(vp.std(t).numpy(), g['vp_std'])
  np.testing.assert_array_equal(vp.mean_coeff(t).numpy(), g['vp_mean'])
  np.testing.assert_array_equal(torch.sqrt(vp.beta(t)).numpy(), g['vp_g'])
  ve = ref_train.make_sde(configs.celebahq_uncsnpp_st())
  np.testing.assert_array_equal(ve.std(torch.tensor(g['ve_t'])).numpy(), g['ve_std'])
  assert ve.eps == float(g['ve_eps']) == float(g['ve_tmin'])
  cfg3 = configs.celeba_uncsnpp_st()
  rve = ref_train.make_sde(cfg3)
  t, _ = rve.time_from_uniform(torch.tensor(g['rve_u']), rve.t_min_from_uniform(cfg3, 0.3), False)
  np.testing.assert_array_equal(t.numpy(), g['rve_t'])
  np.testing.assert_array_equal(rve.std(t).numpy(), g['rve_std'])
  assert float(g['rve_tmin']) == rve.eps


def test_product_sde_lib_vs_reference(golden):
  """The host-side sde_lib of the product restates the same scalars bit-for-bit."""
  from soft_truncation_b200 import sde_lib
  g = golden('sde_golden.npz')
  cfg = configs.cifar10_ddpmpp_nll_st()
  vp = sde_lib.get_sde(cfg)
  np.random.seed(0)
  np.testing.assert_allclose([vp.get_t_min(cfg) for _ in range(5)], g['vp_tmin'], rtol=0)
  torch.manual_seed(0)
  t, Z = vp.get_diffusion_time(cfg, 7, 'cpu', float(g['vp_tmin_used']), importance_sampling=True)
  np.testing.assert_array_equal(t.numpy(), g['vp_t_is'])
  np.testing.assert_array_equal(Z.numpy(), g['vp_Z'])
  mean, std = vp.marginal_prob(torch.ones(7, 1, 1, 1), t)
  np.testing.assert_array_equal(std.numpy(), g['vp_std'])
  np.testing.assert_array_equal(mean.reshape(-1).numpy(), g['vp_mean'])
  np.testing.assert_array_equal(vp.sde(torch.ones(7, 1, 1, 1), t)[1].numpy(), g['vp_g'])
  cfg5 = configs.celebahq_uncsnpp_st()
  ve = sde_lib.get_sde(cfg5)
  x = torch.zeros(7, 1, 1, 1)
  tt = torch.tensor(g['ve_t'])
  np.testing.assert_array_equal(ve.marginal_prob(x, tt)[1].numpy(), g['ve_std'])
  np.testing.assert_array_equal(ve.sde(x, tt)[1].numpy(), g['ve_g'])
  np.testing.assert_array_equal(ve.discretize(x, tt)[1].numpy(), g['ve_G'])
  assert ve.get_t_min(cfg5) == float(g['ve_tmin'])        # st ignored for VE (F6)
  cfg3 = configs.celeba_uncsnpp_st()
  rve = sde_lib.get_sde(cfg3)
  torch.manual_seed(1)
  t_r, _ = rve.get_diffusion_time(cfg3, 7, 'cpu', rve.get_t_min(cfg3))
  np.testing.assert_array_equal(t_r.numpy(), g['rve_t'])
  np.testing.assert_array_equal(rve.marginal_prob(x, t_r)[1].numpy(), g['rve_std'])
  np.testing.assert_allclose(rve.sde(x, t_r)[1].numpy(), g['rve_g'], rtol=1e-6)
  with pytest.raises(AttributeError):
    rve.discretize(x, t_r)                                  # F7: broken as shipped


def test_unet_oracle_vs_reference_cifar(golden):
  g = golden('unet_cifar_golden.npz')
  cfg = configs.cifar10_ddpmpp_nll_st()
  sd = ref_model.make_state_dict(cfg, seed=int(g['seed']))
  n_params = sum(v.numel() for k, v in sd.items() if k != 'sigmas')
  assert n_params == int(g['n_params']) == 61804419
  names = [k for k in sd if k != 'sigmas']
  assert names == list(g['param_names'])
  for k in names:
    sd[k].requires_grad_(True)
  acts = {}
  x = torch.tensor(g['x'])
  out = ref_model.unet_forward(sd, cfg, x, torch.tensor(g['labels']), taps_out=acts)
  assert rel_l2(out, g['out']) < 2e-5
  for i, a in acts.items():
    a = a.double().reshape(-1)
    mean, std = g['act_stats'][i]
    assert abs(a.mean().item() - mean) < 1e-4 * (abs(mean) + std), i
    assert abs(a.std().item() - std) < 1e-4 * std, i
  (out * torch.tensor(g['wout'])).sum().backward()
  gn = np.array([sd[k].grad.double().norm().item() for k in names])
  # biases that feed a GroupNorm have mathematically zero gradient (pure rounding noise ~1e-8)
  np.testing.assert_allclose(gn, g['grad_norms'], rtol=1e-4, atol=1e-6)
  for j in (0, 5, 100, 300, len(names) - 1):
    p = sd[names[j]]
    idx = np.unique(np.linspace(0, p.numel() - 1, 6).astype(np.int64))
    np.testing.assert_allclose(p.grad.reshape(-1)[idx].numpy(), g['grad_samples'][j][:len(idx)],
                               rtol=2e-3, atol=1e-5 * g['grad_norms'][j])


def test_init_statistics_vs_reference():
  want = json.load(open(os.path.join(GOLDEN, 'init_golden.json')))
  cfg = configs.cifar10_ddpmpp_nll_st()
  sd = ref_model.make_state_dict(cfg, seed=0, rezero=False)
  names = [k for k in sd if k != 'sigmas']
  assert names == want['names']
  for k, std, shape in zip(names, want['std'], want['shapes']):
    assert list(sd[k].shape) == shape
    if sd[k].numel() > 512:
      assert abs(sd[k].double().std().item() - std) <= 0.1 * std + 1e-12, k


@pytest.mark.parametrize('tag', ['c3', 'c5'])
def test_unet_oracle_vs_reference_variants(golden, tag):
  g = golden('variants_golden.npz')
  cfg = _reduced_c3() if tag == 'c3' else _reduced_c5()
  sd = ref_model.make_state_dict(cfg, seed=int(g[f'{tag}_seed']))
  names = [k for k in sd if k != 'sigmas']
  assert names == list(g[f'{tag}_names'])
  x = torch.tensor(g[f'{tag}_x'])
  score = ref_model.unet_forward(sd, cfg, x, torch.tensor(g[f'{tag}_sig']))
  assert rel_l2(score, g[f'{tag}_score']) < 2e-5
  sde = ref_train.make_sde(cfg)
  state = ref_train.TrainState(sd)
  for k in state.trainable:
    state.sd[k].requires_grad_(True)
  losses = ref_train.dsm_losses(state.sd, cfg, sde, x, torch.tensor(g[f'{tag}_u']), torch.tensor(g[f'{tag}_z']),
                                float(g[f'{tag}_tmin']))
  np.testing.assert_allclose(losses.detach().numpy(), g[f'{tag}_losses'], rtol=2e-4)
  torch.mean(losses).backward()
  gn = np.array([0. if state.sd[k].grad is None else state.sd[k].grad.double().norm().item() for k in names])
  np.testing.assert_allclose(gn, g[f'{tag}_gnorm'], rtol=2e-3, atol=1e-6 * g[f'{tag}_gnorm'].max())


def test_deepest_lsgm_mixed_step_oracle_vs_reference(golden):
  """reduced-width "deepest" DDPM++ (lsgm embedding, FIR, ch_mult (1,1,1)) and three optimizer steps of
  step_fn_mixed (reference losses.py:295-320) with replayed draws."""
  g = golden('deepest_golden.npz')
  cfg = _reduced_deepest()
  sd = ref_model.make_state_dict(cfg, seed=int(g['seed']))
  names = [k for k in sd if k != 'sigmas']
  assert names == list(g['names'])
  out = ref_model.unet_forward(sd, cfg, torch.tensor(g['x']), torch.tensor(g['labels']))
  assert rel_l2(out, g['out']) < 2e-5
  sde = ref_train.make_sde(cfg)
  state = ref_train.TrainState(sd)
  batch = torch.tensor(g['batch'])
  for s in range(g['losses'].shape[0]):
    losses, _ = ref_train.train_step(state, cfg, sde, batch, torch.tensor(g['u'][s]), torch.tensor(g['z'][s]), float(g['U'][s]))
    np.testing.assert_allclose(losses.numpy(), g['losses'][s], rtol=5e-4)
  pn = np.array([state.sd[k].double().norm().item() for k in names])
  en = np.array([state.ema[k].double().norm().item() for k in names])
  np.testing.assert_allclose(pn, g['pnorm'], rtol=1e-4, atol=2e-6)     # Adam turns noise-level gradients into lr-sized steps
  np.testing.assert_allclose(en, g['enorm'], rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize('tag', ['vp', 've'])
def test_likelihood_weighted_loss_branch_oracle_vs_reference(golden, tag):
  """losses.py:125-129 ((score + z/std)^2 g^2 with uniform times): reached only with importance_sampling off."""
  g = golden('lossbranch_golden.npz')
  if tag == 'vp':
    cfg = configs.cifar10_ddpmpp_nll_st()
    cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks = 32, (1, 2), 1
  else:
    cfg = _reduced_c5()
  cfg.model.dropout = 0.
  cfg.training.likelihood_weighting, cfg.training.importance_sampling = True, False
  sde = ref_train.make_sde(cfg)
  state = ref_train.TrainState(ref_model.make_state_dict(cfg, seed=int(g[f'{tag}_seed'])))
  for k in state.trainable:
    state.sd[k].requires_grad_(True)
  losses = ref_train.dsm_losses(state.sd, cfg, sde, torch.tensor(g[f'{tag}_x']), torch.tensor(g[f'{tag}_u']),
                                torch.tensor(g[f'{tag}_z']), float(g[f'{tag}_tmin']))
  np.testing.assert_allclose(losses.detach().numpy(), g[f'{tag}_losses'], rtol=2e-4)
  torch.mean(losses).backward()
  names = [k for k in state.sd if k != 'sigmas']
  gn = np.array([0. if state.sd[k].grad is None else state.sd[k].grad.double().norm().item() for k in names])
  np.testing.assert_allclose(gn, g[f'{tag}_gnorm'], rtol=2e-3, atol=1e-6 * g[f'{tag}_gnorm'].max())


_RECON = {'uni_sf': ('uniform', 'scoreflow', False), 'uni_ddpm': ('uniform', 'ddpm', True), 'lossless': ('lossless', 'scoreflow', False)}


@pytest.mark.parametrize('tag', sorted(_RECON))
def test_reconstruction_term_oracle_vs_reference(golden, tag):
  """losses.py:134-164 (training.reconstruction_loss, off in every shipped config): DSM loss + decoder term, both decoder
  variances, Gaussian cross-entropy and discretised-likelihood forms - losses and every gradient norm."""
  g = golden('recon_golden.npz')
  deq, variance, reduce_mean = _RECON[tag]
  cfg = configs.cifar10_ddpmpp_nll_st()
  cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks = 32, (1, 2), 1
  cfg.model.dropout = 0.
  cfg.training.reconstruction_loss, cfg.training.reduce_mean, cfg.data.dequantization = True, reduce_mean, deq
  sde = ref_train.make_sde(cfg)
  state = ref_train.TrainState(ref_model.make_state_dict(cfg, seed=int(g[f'{tag}_seed'])))
  for k in state.trainable:
    state.sd[k].requires_grad_(True)
  losses = ref_train.dsm_losses(state.sd, cfg, sde, torch.tensor(g[f'{tag}_x']), torch.tensor(g[f'{tag}_u']),
                                torch.tensor(g[f'{tag}_z']), float(g[f'{tag}_tmin']), importance_sampling=True,
                                z2=torch.tensor(g[f'{tag}_z2']), variance=variance)
  np.testing.assert_allclose(losses.detach().numpy(), g[f'{tag}_losses'], rtol=2e-4)
  torch.mean(losses).backward()
  names = [k for k in state.sd if k != 'sigmas']
  gn = np.array([0. if state.sd[k].grad is None else state.sd[k].grad.double().norm().item() for k in names])
  np.testing.assert_allclose(gn, g[f'{tag}_gnorm'], rtol=2e-3, atol=1e-6 * g[f'{tag}_gnorm'].max())


class _OracleNet(torch.nn.Module):
  """The oracle U-Net behind the model call convention, so that the host-side estimators can run on the CPU."""

  def __init__(self, sd, cfg):
    super().__init__()
    self.sd, self.cfg = sd, cfg

  def forward(self, x, labels):
    return ref_model.unet_forward(self.sd, self.cfg, x, labels)


def test_likelihood_host_logic_vs_reference_fixture(golden):
  """likelihood.py + ode.py (host logic; the network is the CPU oracle here) against the reference's
  likelihood.py: Hutchinson divergence, bits/dim through the device-resident RK45 (same nfev as scipy), NELBO sample."""
  from soft_truncation_b200 import likelihood, sde_lib
  g = golden('likelihood_golden.npz')
  cfg = configs.cifar10_ddpmpp_nll_st()
  cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks = 32, (1, 2), 1
  sde = sde_lib.get_sde(cfg)
  net = _OracleNet(ref_model.make_state_dict(cfg, seed=int(g['seed'])), cfg)
  inv = lambda v: (v + 1.) / 2.
  score_fn = mutils.get_score_fn(cfg, sde, net, train=False, continuous=True)
  rsde = sde.reverse(score_fn, probability_flow=cfg.eval.probability_flow, lambda_=cfg.eval.lambda_)
  drift_fn = lambda xx, tt: rsde.sde(xx, tt)[0]
  x, t = torch.tensor(g['x']), torch.tensor(g['t'])
  with torch.no_grad():
    assert rel_l2(drift_fn(x, t), g['drift']) < 1e-5
  div = likelihood.get_div_fn(drift_fn)(x, t, torch.tensor(g['eps_h']))
  np.testing.assert_allclose(div.numpy(), g['div'], rtol=2e-4)
  data = torch.tensor(g['data'])
  lik = likelihood.get_likelihood_fn(cfg, sde, inv, rtol=float(g['lik_rtol']), atol=float(g['lik_rtol']))
  bpd, latent, nfe = lik(net, data, eps=float(g['lik_eps']),
                         injected=dict(epsilon=torch.tensor(g['lik_epsilon']), z=torch.tensor(g['lik_z']),
                                       z_res=torch.tensor(g['lik_z_res'])))
  assert nfe == int(g['lik_nfe'])
  np.testing.assert_allclose(bpd.numpy(), g['lik_bpd'], rtol=1e-4)
  assert rel_l2(latent, g['lik_latent']) < 1e-4
  elbo = likelihood.get_elbo_fn(cfg, sde, inv)
  nelbo, resid = elbo(net, data, eps=float(g['elbo_eps']),
                      injected=dict(u=torch.tensor(g['elbo_u']), z=torch.tensor(g['elbo_z']),
                                    epsilon=torch.tensor(g['elbo_epsilon']), lp_z=torch.tensor(g['elbo_lp_z']),
                                    z_res=torch.tensor(g['elbo_z_res'])))
  np.testing.assert_allclose(nelbo.numpy(), g['elbo_nelbo'], rtol=1e-4)
  np.testing.assert_allclose(resid.numpy(), g['elbo_resid'], rtol=1e-4)


@pytest.mark.parametrize('tag', ['w5000', 'w0'])
def test_train_trajectory_oracle_vs_reference(golden, tag):
  """10 optimizer steps, B=4 (BASELINE configs[0]); warm-up 5000 as shipped and 0."""
  g = golden('train_golden.npz')
  cfg = configs.cifar10_ddpmpp_nll_st()
  cfg.model.dropout = 0.
  cfg.optim.warmup = 5000 if tag == 'w5000' else 0
  sde = ref_train.make_sde(cfg)
  state = ref_train.TrainState(ref_model.make_state_dict(cfg, seed=int(g['seed'])))
  batch = torch.tensor(g['batch'])
  n_steps = 10 if tag == 'w0' else 4      # the warm-up variant barely moves the weights; 4 steps suffice
  for s in range(n_steps):
    losses, _ = ref_train.train_step(state, cfg, sde, batch, torch.tensor(g['u'][s]), torch.tensor(g['z'][s]),
                                     float(g['U'][s]))
    np.testing.assert_allclose(losses.numpy(), g[f'{tag}_losses'][s], rtol=5e-4), s
  if tag == 'w0':
    for j, name in enumerate(g['probe_names']):
      p = state.sd[str(name)]
      np.testing.assert_allclose(p.double().norm().item(), g['w0_pnorm'][j], rtol=1e-4)
      np.testing.assert_allclose(state.ema[str(name)].double().norm().item(), g['w0_enorm'][j], rtol=1e-4)


def test_pc_sampler_oracle_vs_reference(golden):
  g = golden('sampler_golden.npz')
  cfg = configs.cifar10_ddpmpp_nll_st()
  sd = ref_model.make_state_dict(cfg, seed=1)
  sde8 = ref_train.make_sde(cfg, N=8)
  trace = []
  x = ref_train.pc_sample(sd, cfg, sde8, torch.tensor(g['vp_xT']), list(torch.tensor(g['vp_z'])),
                          float(g['vp_eps']), predictor='euler_maruyama', corrector='none', trace=trace)
  assert int(g['vp_nfe']) == 8 * 2
  for i, st in enumerate(trace):
    assert rel_l2(st, g['vp_trace'][i]) < 1e-4, i
  assert rel_l2(x, g['vp_x']) < 1e-4

  cfg = _reduced_c5()
  sd = ref_model.make_state_dict(cfg, seed=5)
  sde4 = ref_train.make_sde(cfg, N=4)
  z = torch.tensor(g['ve_z'])
  noises = [(z[2 * i], z[2 * i + 1]) for i in range(4)]
  trace = []
  x = ref_train.pc_sample(sd, cfg, sde4, torch.tensor(g['ve_xT']), noises, float(g['ve_eps']),
                          predictor='reverse_diffusion', corrector='langevin', snr=cfg.sampling.snr, trace=trace)
  for i, st in enumerate(trace):
    assert rel_l2(st, g['ve_trace'][i]) < 1e-4, i
  assert rel_l2(x, g['ve_x']) < 1e-4
