"""Hyper-parameter trees for the BASELINE.json workloads.

The reference keeps these as `ml_collections` files that cannot travel to the GPU
box, so the values of the four configs the hot path is measured on are restated
here as data (tests/test_configs.py diffs them against the reference files when
/root/reference is present):

  cifar10_ddpmpp_nll_st   <- configs/vp/CIFAR10/ddpmpp_nll_st.py      (C1 / C2)
  imagenet32_ddpmpp_nll   <- configs/vp/IMAGENET32/ddpmpp_nll.py       (C4)
  celeba_uncsnpp_st       <- configs/ve/CELEBA/uncsnpp_st.py           (C3)
  celebahq_uncsnpp_st     <- configs/ve/celebahq/uncsnpp_st.py         (C5)

Defaults follow configs/default_{cifar10,celeba,lsun}_configs.py.
"""
import torch

from .config_dict import ConfigDict


def _device():
  return torch.device('cuda:0') if torch.cuda.is_available() else torch.device('cpu')


_TRAINING = dict(
    batch_size=128, n_iters=13000001, snapshot_freq=100000, log_freq=100, eval_freq=100,
    snapshot_freq_for_preemption=10000, snapshot_sampling=False, likelihood_weighting=True,
    continuous=True, reduce_mean=False, importance_sampling=True,
    unbounded_parametrization=False, ddpm_score=True, st=False, truncation_time=1e-5,
    num_train_data=50000, reconstruction_loss=False, stabilizing_constant=1e-3,
    whatever_sampling=False, mixed=False, ddpm_weight=0.01, balanced=False)

_SAMPLING = dict(n_steps_each=1, noise_removal=True, probability_flow=False, snr=0.16,
                 batch_size=1024, truncation_time=1e-5, sample_more=True)

_EVAL = dict(begin_ckpt=9, end_ckpt=26, batch_size=200, enable_sampling=False,
             num_samples=50000, enable_loss=True, enable_bpd=False, bpd_dataset='test',
             num_test_data=10000, residual=True, lambda_=0.0, probability_flow=True,
             nelbo_iter=0, nll_iter=0)

_DATA = dict(dataset='CIFAR10', image_size=32, random_flip=True, centered=False,
             dequantization='none', num_channels=3)

_MODEL = dict(sigma_min=0.01, sigma_max=50, num_scales=1000, beta_min=0.1, beta_max=20.,
              dropout=0.1, embedding_type='fourier', auxiliary_resblock=True, attention=True,
              fourier_feature=False, lsgm=False)

_OPTIM = dict(weight_decay=0.0, optimizer='Adam', lr=2e-4, beta1=0.9, eps=1e-8, warmup=5000,
              grad_clip=1., num_micro_batch=1, amsgrad=False)

# the NCSN++ body shared by all four (configs/*/…: model.* block)
_NCSNPP = dict(name='ncsnpp', normalization='GroupNorm', nonlinearity='swish', nf=128,
               ch_mult=(1, 2, 2, 2), num_res_blocks=4, attn_resolutions=(16,),
               resamp_with_conv=True, conditional=True, fir=False, fir_kernel=[1, 3, 3, 1],
               skip_rescale=True, resblock_type='biggan', progressive='none',
               progressive_input='none', progressive_combine='sum', attention_type='ddpm',
               init_scale=0., fourier_scale=16, conv_size=3)


def _base(training=None, sampling=None, evaluate=None, data=None, model=None, optim=None):
  cfg = ConfigDict()
  cfg.training = ConfigDict({**_TRAINING, **(training or {})})
  cfg.sampling = ConfigDict({**_SAMPLING, **(sampling or {})})
  cfg.eval = ConfigDict({**_EVAL, **(evaluate or {})})
  cfg.data = ConfigDict({**_DATA, **(data or {})})
  cfg.model = ConfigDict({**_MODEL, **_NCSNPP, **(model or {})})
  cfg.optim = ConfigDict({**_OPTIM, **(optim or {})})
  cfg.seed = 42
  cfg.device = _device()
  return cfg


def cifar10_ddpmpp_nll_st():
  """DDPM++ (VP) on CIFAR-10 with soft truncation (C1 parity gate, C2 bench)."""
  return _base(
      training=dict(sde='vpsde', continuous=True, reduce_mean=True, st=True, k=1.0,
                    likelihood_weighting=False, truncation_time=1e-5),
      sampling=dict(method='ode', predictor='euler_maruyama', corrector='none'),
      data=dict(centered=True),
      model=dict(scale_by_sigma=False, ema_rate=0.9999, embedding_type='positional'))


def cifar10_ddpmpp_fid_st_deepest():
  """The README's headline FID model: DDPM++ "deepest" - nf=512, three 32x32/16x16/8x8 levels of 8
  res-blocks, FIR resampling, `lsgm` time embedding, mixed importance-sampled + uniform-time loss (step_fn_mixed)."""
  return _base(
      training=dict(sde='vpsde', continuous=True, reduce_mean=True, likelihood_weighting=False,
                    importance_sampling=False, st=True, k=0.9, mixed=True, ddpm_weight=100.0,
                    truncation_time=1e-5),
      sampling=dict(method='pc', predictor='euler_maruyama', corrector='none'),
      data=dict(centered=True),
      model=dict(scale_by_sigma=False, ema_rate=0.9999, nf=512, ch_mult=(1, 1, 1), num_res_blocks=8, fir=True,
                 embedding_type='positional', embedding_dim=128, dropout=0.2, lsgm=True))


def imagenet32_ddpmpp_nll():
  """DDPM++ (VP) on ImageNet32 with likelihood weighting (C4)."""
  return _base(
      # likelihood_weighting and importance_sampling are both on by default -> eps = 5e-5
      training=dict(sde='vpsde', continuous=True, reduce_mean=True, num_train_data=1281149,
                    truncation_time=5e-5),
      sampling=dict(method='ode', predictor='euler_maruyama', corrector='none'),
      evaluate=dict(num_test_data=49999),
      data=dict(centered=True, dataset='IMAGENET32'),
      model=dict(scale_by_sigma=False, ema_rate=0.9999, embedding_type='positional',
                 dropout=0.))


def celeba_uncsnpp_st():
  """UNCSN++ (RVE) on CelebA 64x64 (C3)."""
  return _base(
      training=dict(n_iters=1300001, snapshot_freq=50000, log_freq=50, snapshot_sampling=True,
                    likelihood_weighting=False, num_train_data=162770,
                    sde='reciprocal_vesde', continuous=True, importance_sampling=False, st=True,
                    truncation_time=1e-5, model_mode='reciprocal', eta=1e-3),
      sampling=dict(snr=0.17, batch_size=512, method='pc', predictor='reverse_diffusion',
                    corrector='langevin'),
      evaluate=dict(begin_ckpt=1, batch_size=1024, num_test_data=19962),
      data=dict(dataset='CELEBA', image_size=64),
      model=dict(sigma_max=90., scale_by_sigma=True, sigma_begin=90, ema_rate=0.999, fir=True,
                 progressive_input='residual', init_scale=0.0, fourier_feature=False,
                 sigma_min=1e-3),
      optim=dict(weight_decay=0))


def celebahq_uncsnpp_st():
  """NCSN++ (VE; the reference labels it UNCSN++) on CelebA-HQ 256x256 (C5)."""
  cfg = _base(
      training=dict(batch_size=64, n_iters=24000001, snapshot_freq=200000, log_freq=1000,
                    eval_freq=500, snapshot_freq_for_preemption=5000,
                    likelihood_weighting=False, importance_sampling=False,
                    num_train_data=162770, sde='vesde', continuous=True, st=True, k=2.0,
                    truncation_time=1e-5),
      sampling=dict(snr=0.075, batch_size=16, truncation_time=1e-3, method='pc',
                    predictor='reverse_diffusion', corrector='langevin',
                    probability_flow=False),
      evaluate=dict(begin_ckpt=50, end_ckpt=96, batch_size=512, enable_sampling=True),
      data=dict(dataset='CelebAHQ', image_size=256),
      model=dict(sigma_max=348, num_scales=2000, dropout=0., scale_by_sigma=True,
                 ema_rate=0.999, ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2, fir=True,
                 progressive='output_skip', progressive_input='input_skip', init_scale=0.,
                 fourier_feature=False),
      optim=dict(weight_decay=0))
  # default_lsun_configs.py has no sample_more / several eval keys
  for k in ('sample_more',):
    cfg.sampling._fields.pop(k, None)
  for k in ('num_test_data', 'residual', 'lambda_', 'probability_flow', 'nelbo_iter', 'nll_iter'):
    cfg.eval._fields.pop(k, None)
  return cfg


_REGISTRY = {
    'vp/CIFAR10/ddpmpp_nll_st': cifar10_ddpmpp_nll_st,
    'vp/CIFAR10/ddpmpp_fid_st_deepest': cifar10_ddpmpp_fid_st_deepest,
    'vp/IMAGENET32/ddpmpp_nll': imagenet32_ddpmpp_nll,
    've/CELEBA/uncsnpp_st': celeba_uncsnpp_st,
    've/celebahq/uncsnpp_st': celebahq_uncsnpp_st,
}


def get_config(name):
  """Look a config up by its path under the reference's `configs/` (no `.py`)."""
  name = name.replace('configs/', '').replace('.py', '')
  if name not in _REGISTRY:
    raise KeyError(f'unknown config {name!r}; have {sorted(_REGISTRY)}')
  return _REGISTRY[name]()
