"""CPU-side tests: the C-ABI library loads and exports every symbol include/st_b200.h declares, the host-side
mirror of the reference interface (registries, parameter names/shapes/initialisation, checkpoint formats) behaves
like the reference, and the data-parallel plumbing works under a world_size-2 gloo group.  No kernel is launched."""
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol():
  import ctypes
  from soft_truncation_b200 import _lib
  header = open(os.path.join(ROOT, 'include', 'st_b200.h')).read()
  declared = set(re.findall(r'^(?:int|const char\*)\s+(st_\w+)\s*\(', header, flags=re.M))
  assert len(declared) >= 30
  raw = ctypes.CDLL(_lib.LIB_PATH)
  for name in declared:
    assert hasattr(raw, name), f'{name} declared in st_b200.h but not exported'
  assert declared - {'st_last_error', 'st_gemm_simt_fallback_reason'} == set(_lib.SIGNATURES), 'ctypes signatures out of sync with the header'
  assert _lib.lib.st_version() >= 100
  assert isinstance(_lib.lib.st_last_error(), bytes)


def test_gemm_args_struct_matches_header_layout():
  from soft_truncation_b200._lib import GemmArgs
  header = open(os.path.join(ROOT, 'include', 'st_b200.h')).read()
  body = header[header.index('typedef struct st_gemm_args {'):header.index('} st_gemm_args;')]
  fields = []
  for decl in re.findall(r'^\s*(?:const\s+)?(?:int32_t|int64_t|uint8_t|float|void\*|float\*|void)\W[^;]*;', body, flags=re.M):
    decl = decl.split('/*')[0].strip().rstrip(';')
    names = decl.replace('const', '').replace('*', ' ').split(None, 1)[1]
    fields += [n.strip() for n in names.split(',')]
  assert fields == [f[0] for f in GemmArgs._fields_]


def _cifar():
  from soft_truncation_b200 import configs
  return configs.cifar10_ddpmpp_nll_st()


@pytest.fixture(scope='module')
def cifar_model():
  from soft_truncation_b200.models import ncsnpp
  return ncsnpp.NCSNpp(_cifar(), None, compute_dtype=torch.float32, seed=0)


def test_model_parameters_match_reference_names_shapes_and_init(cifar_model):
  want = json.load(open(os.path.join(GOLDEN, 'init_golden.json')))
  sd = cifar_model.state_dict()
  names = [k for k in sd if k != 'sigmas']
  assert names == want['names']
  assert [list(sd[k].shape) for k in names] == want['shapes']
  assert sum(p.numel() for p in cifar_model.parameters()) == 61804419
  assert len(list(cifar_model.parameters())) == 564
  for k, std in zip(names, want['std']):
    if sd[k].numel() > 512:
      assert abs(sd[k].double().std().item() - std) <= 0.1 * std + 1e-12, k
  assert sd['sigmas'].shape == (1000,)


def test_state_dict_roundtrip_and_physical_layout(cifar_model):
  from oracle import ref_model
  m = cifar_model
  osd = ref_model.make_state_dict(_cifar(), seed=3)
  res = m.load_state_dict(osd, strict=True)
  assert not res.missing_keys and not res.unexpected_keys
  sd = m.state_dict()
  for k, v in osd.items():
    assert torch.equal(sd[k], v), k
  # kernel-order views: conv weights [Cout][kh*kw][Cin(padded)], NIN [out][in], packed q/k/v
  w = osd['all_modules.3.Conv_0.weight']
  assert torch.equal(m.P.f('all_modules.3.Conv_0.weight'), w.permute(0, 2, 3, 1).reshape(128, -1))
  first = m.P.f('all_modules.2.weight').view(128, 9, 64)
  assert torch.equal(first[:, :, :3], osd['all_modules.2.weight'].permute(0, 2, 3, 1).reshape(128, 9, 3))
  assert not first[:, :, 3:].any()                      # zero padding of the 3-channel image axis
  head = m.P.f('all_modules.54.weight')
  assert head.shape == (64, 9 * 128) and not head[3:].any()
  qkv = m.P.f_group([f'all_modules.9.NIN_{j}.W' for j in range(3)])
  assert torch.equal(qkv, torch.cat([osd[f'all_modules.9.NIN_{j}.W'].t() for j in range(3)]))
  dense = m.P.f_region('dense_w').view(-1, 512)
  assert dense.shape[0] == sum(b.cout for b in m._all_resblocks())
  blk = m._all_resblocks()[5]
  assert torch.equal(dense[blk.dense_off:blk.dense_off + blk.cout], osd[f'all_modules.{blk.idx}.Dense_0.weight'])
  # gradients are views of one flat buffer with the parameters' shapes
  for k, p in m.named_parameters():
    assert p.grad is not None and p.grad.shape == p.shape, k
  m._grad.fill_(1.)
  assert all(bool((p.grad == 1).all()) for p in m.parameters())
  m.zero_grad()
  assert not m._grad.any()


def test_model_refuses_cpu_inputs_and_unbuilt_variants(cifar_model):
  with pytest.raises(RuntimeError):
    cifar_model(torch.zeros(1, 3, 32, 32), torch.zeros(1))
  from soft_truncation_b200.models import ncsnpp
  cfg = _cifar()
  cfg.model.resblock_type = 'ddpm'
  with pytest.raises(NotImplementedError):
    ncsnpp.NCSNpp(cfg, None)


def test_registries_behave_like_the_reference():
  from soft_truncation_b200 import sampling
  from soft_truncation_b200.models import utils as mutils
  assert mutils.get_model('ncsnpp').__name__ == 'NCSNpp'
  with pytest.raises(ValueError):
    mutils.register_model(name='ncsnpp')(type('X', (), {}))
  with pytest.raises(KeyError):
    mutils.get_model('nope')
  for name in ('euler_maruyama', 'reverse_diffusion', 'ancestral_sampling', 'none'):
    assert sampling.get_predictor(name)
  for name in ('langevin', 'ald', 'none'):
    assert sampling.get_corrector(name)
  with pytest.raises(ValueError):
    sampling.register_predictor(name='euler_maruyama')(type('Y', (), {}))
  cfg = _cifar()
  cfg.sampling.method = 'bogus'
  with pytest.raises(ValueError):
    sampling.get_sampling_fn(cfg, None, (1, 3, 32, 32), lambda v: v, 1e-5)
  np.testing.assert_allclose(mutils.get_sigmas(cfg)[[0, -1]], [cfg.model.sigma_max, cfg.model.sigma_min])


def test_fused_adam_and_ema_keep_the_reference_checkpoint_format(cifar_model):
  from soft_truncation_b200 import losses
  from soft_truncation_b200.models.ema import ExponentialMovingAverage
  m = cifar_model
  opt = losses.FusedAdam(m.parameters(), m, lr=2e-4)
  ref = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in m.parameters()], lr=2e-4)
  for p in ref.param_groups[0]['params']:
    p.grad = torch.zeros_like(p)
  ref.step()
  a, b = opt.state_dict(), ref.state_dict()
  assert a['param_groups'][0]['params'] == b['param_groups'][0]['params']
  assert set(a['state'][0]) == set(b['state'][0]) == {'step', 'exp_avg', 'exp_avg_sq'}
  assert all(a['state'][i]['exp_avg'].shape == b['state'][i]['exp_avg'].shape for i in b['state'])
  opt.load_state_dict(b)
  assert opt.t == 1
  with pytest.raises(RuntimeError):
    opt.step()                                           # CPU parameters: no fallback
  ema = ExponentialMovingAverage(m.parameters(), decay=0.9999)
  assert ema.owner is m and len(ema.shadow_params) == 564
  sd = ema.state_dict()
  assert set(sd) == {'decay', 'num_updates', 'shadow_params'}
  assert [ema.next_decay() for _ in range(3)] == [min(0.9999, (1 + n) / (10 + n)) for n in (1, 2, 3)]
  ema.load_state_dict(sd)
  # foreign parameters fall back to the reference's per-tensor behaviour
  lin = torch.nn.Linear(3, 2)
  e2 = ExponentialMovingAverage(lin.parameters(), decay=0.5, use_num_updates=False)
  with torch.no_grad():
    lin.weight.add_(1.)
  before = e2.shadow_params[0].clone()
  e2.update(lin.parameters())
  assert torch.allclose(e2.shadow_params[0], before + 0.5)
  cfg = _cifar()
  assert isinstance(losses.get_optimizer(cfg, lin.parameters()), torch.optim.Adam)


def _dp_worker(rank, world, port, out):
  import torch.distributed as dist
  from soft_truncation_b200 import configs, losses, sde_lib
  from soft_truncation_b200.models import ncsnpp
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  cfg = configs.cifar10_ddpmpp_nll_st()
  cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks = 64, (1,), 1
  cfg.model.attn_resolutions = ()
  m = ncsnpp.NCSNpp(cfg, None, compute_dtype=torch.float32, seed=7)
  m._grad.fill_(float(rank + 1))
  losses.sync_gradients(m)
  ok = bool((m._grad == sum(range(1, world + 1))).all())
  ok = ok and all(bool((p.grad == 3.).all()) for p in m.parameters() if p.requires_grad)
  # t_min: ranks start with different NumPy generator states; after the one-time sync every rank draws rank 0's
  # sequence locally (no per-step collective)
  np.random.seed(100 + rank)
  losses.sync_numpy_rng()
  sde = sde_lib.get_sde(cfg)
  drawn = [losses.shared_t_min(sde, cfg) for _ in range(3)]
  np.random.seed(100)
  ok = ok and drawn == [sde.get_t_min(cfg) for _ in range(3)]
  # replicas that were initialised apart are made equal to rank 0's (parameters, step counter)
  ref = m._flat.clone()                          # seed=7 on every rank: identical so far
  with torch.no_grad():
    m._flat.add_(float(rank))
  state = dict(model=m, optimizer=None, ema=None, step=5 * rank + 2)
  losses.sync_replicas(state)
  ok = ok and bool((m._flat == ref).all()) and state['step'] == 2
  out[rank] = ok
  dist.destroy_process_group()


def test_data_parallel_plumbing_gloo_world2():
  import torch.multiprocessing as mp
  mgr = mp.Manager()
  out = mgr.dict()
  port = 29500 + os.getpid() % 1000
  mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
  assert dict(out) == {0: True, 1: True}


def test_device_rk45_takes_the_same_steps_as_scipy():
  """ode.solve_ivp_rk45 (state as one torch tensor, here on the CPU) against scipy.integrate.solve_ivp(RK45), the
  solver the reference calls (sampling.py:492, likelihood.py:111): same accepted times, nfev and final state,
  forward and backward in time."""
  from scipy import integrate
  from soft_truncation_b200.ode import solve_ivp_rk45
  rng = np.random.default_rng(0)
  A = rng.normal(size=(40, 40)) * 0.5 - np.eye(40)
  At = torch.tensor(A)
  y0 = rng.normal(size=40)
  for span, rtol, atol in (((0., 2.), 1e-5, 1e-5), ((1., 1e-3), 1e-3, 1e-6), ((0., 0.5), 1e-8, 1e-10)):
    want = integrate.solve_ivp(lambda t, y: A @ y * np.cos(3 * t) + np.sin(5 * t + y), span, y0, rtol=rtol, atol=atol,
                               method='RK45')
    got = solve_ivp_rk45(lambda t, y: At @ y * np.cos(3 * t) + torch.sin(5 * t + y), span, torch.tensor(y0),
                         rtol=rtol, atol=atol)
    assert got.success and got.nfev == want.nfev and len(got.ts) == len(want.t)
    np.testing.assert_allclose(np.array(got.ts), want.t, rtol=1e-6, atol=1e-9)    # step factors: err_norm ** -0.2
    np.testing.assert_allclose(got.y.numpy(), want.y[:, -1], rtol=1e-7, atol=1e-9)


def test_u8_loader_epochs_and_scalers():
  """datasets.U8Loader / get_batch (reference datasets.py:106-113: restart the epoch when exhausted) and the data
  scalers (datasets.py:56-71)."""
  from soft_truncation_b200 import configs, datasets
  imgs = (np.arange(10 * 4 * 4 * 3) % 251).astype(np.uint8).reshape(10, 4, 4, 3)
  ds = datasets.U8Loader(imgs, 4, shuffle=True, seed=3)
  assert len(ds) == 2
  it = iter(ds)
  seen = []
  for _ in range(5):                       # 2 batches per epoch -> crosses two epoch boundaries
    batch, it = datasets.get_batch(None, it, ds)
    assert batch.dtype == torch.uint8 and batch.shape == (4, 4, 4, 3)
    seen.append(batch.numpy().copy())
  flat = imgs.reshape(10, -1)
  for b in seen:                           # every row is one of the source images
    for row in b.reshape(4, -1):
      assert (flat == row).all(1).any()
  first_epoch = np.concatenate(seen[:2]).reshape(8, -1)
  assert len({r.tobytes() for r in first_epoch}) == 8          # no repeats inside an epoch
  with pytest.raises(ValueError):
    datasets.U8Loader(imgs.astype(np.float32), 4)
  cfg = configs.cifar10_ddpmpp_nll_st()
  x = torch.linspace(0, 1, 7)
  assert torch.allclose(datasets.get_data_inverse_scaler(cfg)(datasets.get_data_scaler(cfg)(x)), x)
  assert datasets.get_data_scaler(cfg)(torch.tensor(0.)).item() == -1.
  cfg.data.centered = False
  assert datasets.get_data_scaler(cfg)(x) is x and datasets.get_data_inverse_scaler(cfg)(x) is x
  with pytest.raises(RuntimeError):        # batch preparation is a CUDA kernel: no CPU fallback
    cfg.device = torch.device('cpu')
    datasets.prepare_batch(cfg, torch.zeros(2, 32, 32, 3, dtype=torch.uint8))


def test_groupnorm_launch_policies_for_the_bench_shapes():
  """The cluster-form policies are host functions of the shape (no GPU needed; 148 SMs assumed without a device):
  pins what the B=512 training step launches (measured policy) and the fall-backs."""
  from soft_truncation_b200._lib import lib
  B = 512
  # (hw, C) -> (resident forward cluster, resident backward 2-stream, 3-stream, fused-pair backward 2-stream, 3-stream)
  want = {(1024, 128): (4, 8, 0, 2, 0), (1024, 256): (8, 0, 0, 2, 0), (1024, 384): (13, 0, 0, 2, 0),
          (256, 256): (2, 4, 0, 2, 2), (256, 384): (4, 7, 0, 2, 2), (256, 512): (4, 8, 0, 2, 2),
          (64, 256): (1, 1, 2, 1, 1), (64, 512): (1, 2, 0, 1, 1), (16, 256): (1, 1, 1, 1, 1), (16, 512): (1, 1, 1, 1, 1)}
  for (hw, C), w in want.items():
    got = (lib.st_gn_fwd_fused_chunks(B, hw, C), lib.st_gn_bwd_resident_chunks(B, hw, C, 2),
           lib.st_gn_bwd_resident_chunks(B, hw, C, 3), lib.st_gn_bwd_fused_chunks(B, hw, C, 1, 2),
           lib.st_gn_bwd_fused_chunks(B, hw, C, 1, 3))
    assert got == w, ((hw, C), got, w)
  # an image that does not fit 16 CTAs (CelebA-HQ 256x256), a batch too small to fill the GPU, a 4-stream call
  assert lib.st_gn_fwd_fused_chunks(16, 65536, 128) == 0 and lib.st_gn_bwd_resident_chunks(4, 1024, 128, 2) == 0
  assert lib.st_gn_bwd_resident_chunks(B, 64, 256, 4) == 0 and lib.st_gn_bwd_resident_chunks(B, 64, 1024, 2) == 0


def test_sde_classes_and_reverse_sde_vs_reference_fixture():
  """SURVEY 8(a) a6-a9 for VPSDE, subVPSDE and VESDE: sde / marginal_prob / prior_logp / discretize and the
  reverse-time SDE, its lambda = 0.5 interpolation and the probability-flow ODE (reference sde_lib.py:55-119)."""
  from soft_truncation_b200 import sde_lib
  g = np.load(os.path.join(GOLDEN, 'sde_reverse_golden.npz'))
  x, t = torch.tensor(g['x']), torch.tensor(g['t'])
  score = lambda xx, tt: -0.3 * xx + 0.1 * tt[:, None, None, None]
  sdes = dict(vp=sde_lib.VPSDE(truncation_time=1e-5, beta_min=0.1, beta_max=20., N=1000),
              subvp=sde_lib.subVPSDE(beta_min=0.1, beta_max=20., N=1000),
              ve=sde_lib.VESDE(sigma_min=0.01, sigma_max=50., N=1000))
  close = lambda a, b: np.testing.assert_allclose(a.numpy(), b, rtol=1e-6, atol=1e-7)
  for tag, sde in sdes.items():
    f, gg = sde.sde(x, t)
    mean, std = sde.marginal_prob(x, t)
    fd, Gd = sde.discretize(x, t)
    for got, key in ((f, 'f'), (gg, 'g'), (mean, 'mean'), (std, 'std'), (sde.prior_logp(x), 'logp'), (fd, 'fd'), (Gd, 'Gd')):
      close(got, g[f'{tag}_{key}'])
    assert float(sde.T) == float(g[f'{tag}_T'])
    for name, pf, lam in (('rsde', False, 1.), ('mix', False, 0.5), ('ode', True, 0.)):
      r = sde.reverse(score, probability_flow=pf, lambda_=lam)
      rf, rg = r.sde(x, t)
      rfd, rGd = r.discretize(x, t)
      for got, key in ((rf, 'f'), (rg, 'g'), (rfd, 'fd'), (rGd, 'Gd')):
        close(got, g[f'{tag}_{name}_{key}'])
    with pytest.raises(AssertionError):          # probability_flow and lambda_ must agree (sde_lib.py:82)
      sde.reverse(score, probability_flow=True, lambda_=1.)


class _LabelEcho(torch.nn.Module):
  def forward(self, x, labels):
    return torch.zeros_like(x) + labels.float()[:, None, None, None]


def test_score_fn_label_and_scale_conventions_vs_reference_fixture():
  """SURVEY 8(a) a10 (models/utils.py:128-190): what get_score_fn feeds the network and how it scales the output, for
  VP continuous (plain / unbounded parametrisation / without ddpm_score), VP discrete, VE continuous and discrete.
  The network is a stub that echoes its labels, so the fixture pins labels and 1/std factors."""
  from soft_truncation_b200 import configs, sde_lib
  from soft_truncation_b200.models import utils as mutils
  g = np.load(os.path.join(GOLDEN, 'score_fn_golden.npz'))
  net = _LabelEcho()
  x, t = torch.zeros(3, 3, 4, 4), torch.tensor(g['t'])
  cfg = configs.cifar10_ddpmpp_nll_st()
  vp = sde_lib.get_sde(cfg)
  close = lambda a, key: np.testing.assert_allclose(a.numpy(), g[key], rtol=2e-6)
  close(mutils.get_score_fn(cfg, vp, net, train=False, continuous=True)(x, t), 'vp_cont')
  close(mutils.get_score_fn(cfg, vp, net, train=False, continuous=False)(x, t), 'vp_disc')
  cfg.training.unbounded_parametrization = True
  close(mutils.get_score_fn(cfg, vp, net, train=False, continuous=True)(x, t), 'vp_unbounded')
  cfg.training.unbounded_parametrization = False
  cfg.training.ddpm_score = False
  close(mutils.get_score_fn(cfg, vp, net, train=False, continuous=True)(x, t), 'vp_raw')
  cfg5 = configs.celebahq_uncsnpp_st()
  ve = sde_lib.get_sde(cfg5)
  close(mutils.get_score_fn(cfg5, ve, net, train=False, continuous=True)(x, t), 've_cont')
  close(mutils.get_score_fn(cfg5, ve, net, train=False, continuous=False)(x, t.clone()), 've_disc')


def test_ema_store_copy_to_restore_like_run_lib(cifar_model):
  """The evaluation dance of run_lib.py:94-97,105-109 / models/ema.py:53-89: store the raw weights, copy the shadow in
  (the flat parameter buffer itself changes: views stay valid), restore."""
  from soft_truncation_b200.models.ema import ExponentialMovingAverage
  m = cifar_model
  ema = ExponentialMovingAverage(m.parameters(), decay=0.999)
  raw = m._flat.clone()
  with torch.no_grad():
    ema.shadow_flat.mul_(0.5)                    # a shadow that differs from the weights
  ema.store(m.parameters())
  ema.copy_to(m.parameters())
  first = next(p for p in m.parameters() if p.requires_grad)
  assert torch.equal(first.data, ema.shadow_params[0]) and not torch.equal(m._flat, raw)
  assert first.data.data_ptr() >= m._flat.data_ptr()          # still a view of the flat buffer
  ema.restore(m.parameters())
  assert torch.equal(m._flat, raw)


def test_bench_reference_arm_prints_exactly_one_json_line():
  """bench.py contract: stdout carries ONE JSON line (library banners go to stderr); the reference arm runs the staged
  reference's own training step (else the oracle port) on the host cores and reports it as its own cpu_baseline / e2e."""
  import subprocess
  import sys
  r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                     capture_output=True, text=True, timeout=600, env=dict(os.environ, ST_BENCH_REF_BATCH='4'))
  assert r.returncode == 0, r.stderr[-2000:]
  lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
  assert len(lines) == 1, r.stdout
  d = json.loads(lines[0])
  assert d['impl'] == 'reference' and d['metric'] == 'DDPM++ CIFAR-10 train images/sec' and d['unit'] == 'images/s'
  assert d['higher_is_better'] is True and d['value'] > 0 and d['n_gpus'] == 1 and d['steps'] == 1
  from baseline import ref_env
  assert d['cpu_baseline']['kind'] == ('reference' if ref_env.locate() else 'port') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
  assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
  assert 'workload' in d['config'] and d['vs_baseline'] is None


def test_reference_utils_drives_this_package_on_the_host():
  """CPU half of tests/test_gpu_round2.py::test_reference_drivers_run_on_this_package: the reference's UNCHANGED utils.py,
  imported with its module names redirected to this package behind the TensorFlow stubs,
  builds the model / optimizer / EMA through `load_model` and the step / sampling / likelihood functions through
  `get_loss_fns` (reference utils.py:49-82).  Runs where the reference is available (build container or staged copy)."""
  import json
  import subprocess
  import sys
  from baseline import ref_env
  from test_gpu_round2 import _DRIVER_SCRIPT
  if ref_env.locate(allow_source=True) is None:
    pytest.skip('the reference is not available here')
  r = subprocess.run([sys.executable, '-c', _DRIVER_SCRIPT.format(root=ROOT, dev='cpu')], capture_output=True, text=True, timeout=600)
  assert r.returncode == 0, r.stderr[-3000:]
  res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')][-1][7:])
  assert res == {'model': 'soft_truncation_b200.models.ncsnpp', 'opt': 'Adam'}


def test_reference_is_staged_and_importable():
  """baseline/_ref (the untouched reference, staged by baseline/ref_env.stage() from /root/reference) is what
  `bench.py --impl reference` and the `gpu_reference` leg run: it must import behind the shims and expose the path's API."""
  from baseline import ref_env
  if ref_env.locate() is None:
    pytest.skip('baseline/_ref is not staged')
  import subprocess
  import sys
  code = ("import sys; sys.path.insert(0, %r); from baseline import ref_env; R = ref_env.import_reference(); "
          "c = ref_env.ref_config('vp/CIFAR10/ddpmpp_nll_st'); "
          "assert callable(R.losses.get_step_fn) and callable(R.sampling.get_sampling_fn) and c.training.sde == 'vpsde'; "
          "print('ok', R.root)" % ROOT)
  r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
  assert r.returncode == 0 and 'ok' in r.stdout, r.stderr[-2000:]


def test_committed_bench_lines_keep_the_contract():
  """The bench lines committed under profiles/ are what DESIGN.md quotes: every one must carry the contract's keys and be
  self-consistent (value = batch * GPUs / time per step; e2e not faster than value beyond run-to-run noise; roofline fraction = achieved / peak)."""
  import glob
  files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r02_bench_*.json')))
  assert len(files) >= 10
  for f in files:
    lines = [l for l in open(f).read().splitlines() if l.strip().startswith('{')]
    assert len(lines) == 1, f
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'dtype', 'data', 'config'):
      assert k in d, (f, k)
    assert 'workload' in d['config'] and 'model' not in d['config'], f
    assert d['value'] > 0 and d['ms_per_step'] > 0
    if d.get('impl') == 'reference':
      assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['e2e']['h2d_bytes_per_step'] == 0
      continue
    if 'sampler' not in os.path.basename(f) or '8gpu' not in f:      # (the 8-GPU sampler lines predate the count)
      assert d['gpu_launches'] > 0, f
    assert d['clocks']['sm_mhz'] > 0, f
    assert not set(d['clocks'].get('reasons', [])) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}, f
    r = d.get('roofline')
    if r and r.get('frac') is not None:
      assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-6 and 0. < r['frac'] < 1., f
    if d['unit'] == 'images/s':
      B = int(re.search(r'batch (\d+)/GPU', d['config']['workload']).group(1))
      assert abs(d['value'] - B * d['n_gpus'] / (d['ms_per_step'] * 1e-3)) < 1e-3 * d['value'], f
      assert d['e2e']['value'] <= d['value'] * 1.03 and d['e2e']['h2d_bytes_per_step'] > 0, f
