"""Round-2 GPU parity tests (the gaps the round-1 review named).

  * the captured-graph training step against the eager step (same draws, same dropout seeds)
  * bf16 mode: EVERY parameter-gradient tensor against the fp32 CPU oracle (cosine and rel-L2, not only norms)
  * bf16 mode: a 100-step loss trajectory against the reference's fp32 trajectory (fixture traj100_golden.npz)
  * the tcgen05 weight / data gradient forms at the BENCH shapes (M*K = 524288 rows, the split-K the dispatcher picks)
    against the fp32-FMA kernel on identical bf16 inputs
  * a checkpoint WRITTEN BY THE REFERENCE (its NCSNpp + torch Adam + its EMA) restored and continued
  * FULL-width C3 (64x64) and C5 (256x256) on one image against the reference fixture
  * the non-fused optimizer path (AdamW) really updates the parameters
  * no bf16 GEMM of the full-size networks falls back to the SIMT kernel
  * the reference's own utils.load_model / get_loss_fns drive this package unchanged (sys.modules redirect)

Tolerances are written next to each comparison.
"""
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, rel_l2
from oracle import ref_model, ref_train

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'


def _note(msg):
  """Measured margins of the tolerance checks, kept beside the run (gpurun_out/test_stats.txt) for DESIGN.md."""
  print(msg)
  d = os.path.join(ROOT, 'gpurun_out')
  if os.path.isdir(d):
    with open(os.path.join(d, 'test_stats.txt'), 'a') as f:
      f.write(msg + '\n')


def _cfg(dropout=None):
  from soft_truncation_b200 import configs
  cfg = configs.cifar10_ddpmpp_nll_st()
  cfg.device = torch.device(DEV)
  if dropout is not None:
    cfg.model.dropout = dropout
  return cfg


def _small(cfg):
  cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks = 128, (1, 2), 1      # the block types of the full net
  return cfg


def _model(cfg, seed, dtype):
  from soft_truncation_b200 import sde_lib
  from soft_truncation_b200.models import utils as mutils
  cfg.model.compute_dtype = 'fp32' if dtype == torch.float32 else 'bf16'
  sde = sde_lib.get_sde(cfg)
  model = mutils.create_model(cfg, sde)
  sd = ref_model.make_state_dict(cfg, seed=seed)
  mutils.unwrap(model).load_state_dict(sd, strict=True)
  return model, sde, sd


def _state(cfg, model):
  from soft_truncation_b200 import losses
  from soft_truncation_b200.models.ema import ExponentialMovingAverage
  return dict(optimizer=losses.get_optimizer(cfg, model.parameters()), model=model,
              ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)


# ------------------------------------------------------------------------------------------------ captured-graph step
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('dropout', [0., 0.1])
def test_step_graph_matches_eager(dtype, dropout):
  """Six optimizer steps eagerly and six with the step captured in CUDA graphs (capture happens on the third call):
  same injected t_min / u / z, same torch seed (-> same in-kernel dropout seeds, which the graph receives through
  device memory).  Losses, parameters and EMA agree to the summation-order noise of the split-K reductions."""
  from soft_truncation_b200 import losses
  from soft_truncation_b200.models import utils as mutils
  B, steps = 8, 6
  gen = torch.Generator().manual_seed(3)
  batch = (torch.rand(B, 3, 32, 32, generator=gen) * 2 - 1).to(DEV)
  draws = [dict(t_min=10 ** float(-5 + 4 * torch.rand(1, generator=gen)), u=torch.rand(B, generator=gen),
                z=torch.randn(B, 3, 32, 32, generator=gen)) for _ in range(steps)]
  runs = {}
  for mode in ('eager', 'graph'):
    torch.manual_seed(11)
    cfg = _small(_cfg(dropout=dropout))
    cfg.optim.warmup = 4                       # the warm-up learning rate changes every step: exercises `dyn`
    cfg.optim.cuda_graph = (mode == 'graph')
    model, sde, _ = _model(cfg, 5, dtype)
    state = _state(cfg, model)
    step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
    ls = [step_fn(state, batch, injected=d) for d in draws]
    net = mutils.unwrap(model)
    runs[mode] = (torch.stack(ls), net._flat.clone(), state['ema'].shadow_flat.clone(), state['step'], state['optimizer'].t,
                  state['ema'].num_updates)
    if mode == 'graph':
      assert step_fn.__closure__ is not None
  (le, pe, ee, se, te, ne), (lg, pg, eg, sg, tg, ng) = runs['eager'], runs['graph']
  assert (se, te, ne) == (sg, tg, ng) == (steps, steps, steps)
  tol = 2e-5 if dtype == torch.float32 else 2e-3
  _note(f'graph vs eager ({dtype}, dropout {dropout}): losses {rel_l2(lg, le):.2e}, params {rel_l2(pg, pe):.2e}, ema {rel_l2(eg, ee):.2e}')
  assert rel_l2(lg, le) < tol
  assert rel_l2(pg, pe) < tol and rel_l2(eg, ee) < tol
  if dropout > 0:
    # masks really change from step to step inside the graph: with identical draws and lr = 0 the losses of
    # consecutive replays differ only through the dropout mask
    cfg = _small(_cfg(dropout=dropout))
    cfg.optim.lr = 0.
    model, sde, _ = _model(cfg, 5, dtype)
    state = _state(cfg, model)
    step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
    ls = torch.stack([step_fn(state, batch, injected=draws[0]) for _ in range(5)])
    assert not torch.allclose(ls[3], ls[4], rtol=1e-4)


@pytest.mark.parametrize('st', [True, False])
def test_step_graph_is_used_and_keeps_rng_stream(st):
  """Without injected draws the graph step consumes torch's CUDA generator exactly like the eager step (u = rand(B), then
  z = randn) - the per-step losses of both modes agree - and the captured graphs are really what runs, with and without
  soft truncation (st=False: the ImageNet32 / C4 recipe, fixed t_min)."""
  from soft_truncation_b200 import losses
  import warnings
  B = 8
  batch = (torch.rand(B, 3, 32, 32, generator=torch.Generator().manual_seed(1)) * 2 - 1).to(DEV)
  out = {}
  for mode in ('eager', 'graph'):
    cfg = _small(_cfg(dropout=0.))
    cfg.training.st = st
    cfg.optim.warmup = 0
    cfg.optim.cuda_graph = (mode == 'graph')
    model, sde, _ = _model(cfg, 6, torch.float32)
    state = _state(cfg, model)
    step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
    np.random.seed(7)
    torch.manual_seed(7)
    torch.cuda.manual_seed(7)
    from soft_truncation_b200 import _lib
    l0 = _lib.launches
    with warnings.catch_warnings():
      warnings.filterwarnings('error', message='.*CUDA-graph capture.*')      # a failed capture only warns
      ls = [step_fn(state, batch) for _ in range(5)]
    out[mode] = (torch.stack(ls), _lib.launches - l0)
  assert rel_l2(out['graph'][0], out['eager'][0]) < 2e-5
  # two eager steps + one capture pass instead of five eager steps worth of C-ABI calls
  assert out['graph'][1] < 0.75 * out['eager'][1]


# ------------------------------------------------------------------------------------------------ bf16 gradients
def test_bf16_gradients_every_tensor_vs_fp32_oracle():
  """bf16 mode (the tcgen05 path the benchmark times): every one of the 564 parameter-gradient tensors of the FULL
  CIFAR-10 DDPM++ against the fp32 CPU oracle on the same inputs (batch 2).  Per tensor: cosine >= 0.99 and
  rel-L2 <= 0.15.  Measured (gpurun_out/test_stats.txt): worst cosine 0.9949, worst rel-L2 0.106, median 0.059 - the
  deviation is direction-preserving noise: every activation is stored with 8 mantissa bits (2^-9 relative rounding)
  and a gradient crosses ~200 such roundings on its way through 55 blocks forward and back, 2^-9 * sqrt(200) ~ 3-6 %
  (the forward output of the same run is within 1.3 %).  The orchestration itself is pinned much tighter by the
  fp32-mode test on the same code (every tensor <= 2e-4) and the tcgen05 kernels by the SIMT comparison on identical
  bf16 inputs (<= 1.1e-4).  Tensors whose true gradient is identically zero up to rounding (the key bias of every
  attention block: softmax is invariant to it) carry no signal and are exempt."""
  from soft_truncation_b200.models import utils as mutils
  cfg = _cfg(dropout=0.)
  model, sde, sd = _model(cfg, 4, torch.bfloat16)
  net = mutils.unwrap(model)
  gen = torch.Generator().manual_seed(5)
  B = 2
  x = torch.randn(B, 3, 32, 32, generator=gen)
  labels = torch.tensor([10., 700.])
  wout = torch.randn(B, 3, 32, 32, generator=gen)
  for k in sd:
    if k != 'sigmas':
      sd[k].requires_grad_(True)
  out_o = ref_model.unet_forward(sd, cfg, x, labels, train=True, drop_masks=None)
  (out_o * wout).sum().backward()
  model.train()
  net.zero_grad()
  out = model(x.to(DEV), labels.to(DEV))
  assert rel_l2(out, out_o.detach()) < 3e-2
  (out * wout.to(DEV)).sum().backward()
  gmax = max(float(sd[k].grad.norm()) for k, _ in net.named_parameters() if sd[k].grad is not None)
  checked, exempt, bad, rows = 0, [], [], []
  for k, p in net.named_parameters():
    ref = sd[k].grad
    if ref is None:
      continue
    if float(ref.norm()) < 1e-6 * gmax:
      exempt.append(k)
      continue
    a, b = p.grad.detach().cpu().double().reshape(-1), ref.double().reshape(-1)
    cos = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300))
    rel = float((a - b).norm() / b.norm())
    rows.append((rel, cos, k))
    if not (cos >= 0.99 and rel <= 0.15):
      bad.append((k, round(cos, 5), round(rel, 4)))
    checked += 1
  rows.sort(reverse=True)
  _note(f'bf16 gradients: {checked} tensors checked, worst rel-L2 {rows[0][0]:.3e} ({rows[0][2]}), worst cosine '
        f'{min(r[1] for r in rows):.5f}, median rel-L2 {rows[len(rows) // 2][0]:.3e}, exempt {exempt}; five worst: '
        + '; '.join(f'{k} rel {r:.3e} cos {c:.5f}' for r, c, k in rows[:5]))
  assert not bad, bad
  assert checked >= 550 and all('NIN_1.b' in k for k in exempt)


# ------------------------------------------------------------------------------------------------ GroupNorm by-product
@pytest.mark.parametrize('case', [(64, 32, 128, 128, False), (64, 16, 256, 256, True), (256, 8, 256, 256, True),
                                  (1024, 4, 256, 256, True), (96, 16, 128, 256, True)])
def test_gemm_epilogue_groupnorm_partials(case):
  """st_gemm's GroupNorm by-product (gn_part: per (min(hw,128) rows, 4 channels) sum / sum of squares of the STORED bf16
  output, emitted by the TMA epilogue of the 256-column tiles, single-CTA and CTA-pair kernels) against the same sums
  taken from the output tensor; then GroupNorm + SiLU driven by them against the st_gn_stats path (mean / rstd to 1e-5,
  y bit-for-bit up to bf16 rounding flips)."""
  from soft_truncation_b200 import ops
  if not ops.tc_available():
    pytest.skip('tcgen05 backend unavailable on this device')
  B, H, Ci, Co, with_res = case
  gen = torch.Generator().manual_seed(B + H)
  x = torch.randn(B, H, H, Ci, generator=gen).to(DEV).to(torch.bfloat16)
  w = (torch.randn(Co, 9 * Ci, generator=gen) / math.sqrt(9 * Ci)).to(DEV).to(torch.bfloat16)
  bias = torch.randn(Co, generator=gen).to(DEV)
  rb = torch.randn(B, Co, generator=gen).to(DEV)
  res = torch.randn(B, H, H, Co, generator=gen).to(DEV).to(torch.bfloat16) if with_res else None
  out, q = ops.conv_fwd(x, w, Co, 3, 3, bias=bias, rowbias=None if with_res else rb, rowbias_ld=Co, residual=res,
                        alpha=0.7 if with_res else 1.0, want_quads=True)
  assert q is not None, 'the epilogue did not emit GroupNorm partial sums for a 256-column tile shape'
  R = min(H * H, 128)
  assert q.rows == R and q.t.shape == (B * H * H // R, Co // 4, 2)
  o = out.float().view(B * H * H // R, R, Co // 4, 4)
  want = torch.stack([o.sum(dim=(1, 3)), (o * o).sum(dim=(1, 3))], dim=-1)
  err = float((q.t - want).abs().max() / want.abs().max())
  _note(f'gn partials {case}: max abs err / max {err:.2e}')
  assert err < 2e-6
  G = 32
  gamma, beta = torch.randn(Co, generator=gen).to(DEV), torch.randn(Co, generator=gen).to(DEV)
  y1, st1 = ops.gn_norm_act(out, None, G, gamma, beta, act=1, quads=(q, None))
  y0, st0 = ops.gn_norm_act(out, None, G, gamma, beta, act=1, fused_chunks=0)
  assert torch.allclose(st1[0], st0[0], rtol=1e-5, atol=1e-6) and torch.allclose(st1[1], st0[1], rtol=1e-5)
  assert rel_l2(y1.float(), y0.float()) < 1e-3
  # two concatenated sources, group size 12 (384 channels): groups straddle the seam; statistics from both buffers
  if Co == 256 and H == 16:
    x2 = torch.randn(B, H, H, 128, generator=gen).to(DEV).to(torch.bfloat16)
    w2 = (torch.randn(128, 9 * 128, generator=gen) / 34.).to(DEV).to(torch.bfloat16)
    out2, q2 = ops.conv_fwd(x2, w2, 128, 3, 3, want_quads=True)
    if q2 is not None:
      g3, b3 = torch.randn(384, generator=gen).to(DEV), torch.randn(384, generator=gen).to(DEV)
      ya, sa = ops.gn_norm_act(out, out2, G, g3, b3, act=1, quads=(q, q2))
      yb, sb = ops.gn_norm_act(out, out2, G, g3, b3, act=1, fused_chunks=0)
      assert torch.allclose(sa[0], sb[0], rtol=1e-5, atol=1e-6) and torch.allclose(sa[1], sb[1], rtol=1e-5)
      assert rel_l2(ya.float(), yb.float()) < 1e-3


def test_network_with_and_without_groupnorm_by_product():
  """The full CIFAR-10 DDPM++ forward (batch 128, bf16, eval) with the GroupNorm statistics taken from the GEMM
  epilogues against the same network computing them from the tensors: same output up to bf16 rounding flips."""
  from soft_truncation_b200 import ops
  from soft_truncation_b200.models import utils as mutils
  if not ops.tc_available():
    pytest.skip('tcgen05 backend unavailable on this device')
  cfg = _cfg(dropout=0.)
  model, sde, _ = _model(cfg, 3, torch.bfloat16)
  model.eval()
  gen = torch.Generator().manual_seed(9)
  x = torch.randn(128, 3, 32, 32, generator=gen).to(DEV)
  labels = (torch.rand(128, generator=gen) * 999).to(DEV)
  outs = {}
  calls = {'n': 0}
  orig = ops.gn_apply

  def spy(*a, **k):
    calls['n'] += int(k.get('quads') is not None)
    return orig(*a, **k)
  ops.gn_apply = spy
  try:
    for flag in (True, False):
      ops.GN_QUADS = flag
      with torch.no_grad():
        outs[flag] = model(x, labels).clone()
  finally:
    ops.GN_QUADS = True
    ops.gn_apply = orig
  _note(f'network with / without GroupNorm by-product: rel-L2 {rel_l2(outs[True], outs[False]):.2e}, {calls["n"]} GroupNorms fed by GEMM epilogues')
  assert rel_l2(outs[True], outs[False]) < 2.5e-2      # measured 1.1e-2: rounding flips seeded by 1e-7 differences in rstd,
                                                          # amplified through 55 blocks exactly like any other bf16 perturbation
  assert calls['n'] >= 40


# ------------------------------------------------------------------------------------------------ fused attention
@pytest.mark.parametrize('n_img,save_p', [(1, True), (3, False), (149, True)])
def test_fused_attention_forward_vs_torch(n_img, save_p):
  """st_attn_fwd (S = QK^T in tensor memory, softmax in registers, P in shared memory, O = PV) against fp32 torch on
  the same bf16 inputs, L = 256, C = 256.  o: rel-L2 <= 6e-3 (P and o are rounded to bf16: 3 ulps rms, the tolerance of
  every bf16 kernel here); p (training only): rel-L2 <= 4e-3 against the fp32 softmax.  149 images = more tiles than
  CTAs: exercises the persistent loop and the buffer hand-over between tiles."""
  from soft_truncation_b200 import ops
  L = C = 256
  if not ops.attn_fused_ok(L, C, torch.bfloat16):
    pytest.skip('fused attention kernel unavailable on this device')
  gen = torch.Generator().manual_seed(n_img)
  qkv = (torch.randn(n_img * L, 3 * C, generator=gen) * 1.5).to(DEV).to(torch.bfloat16)
  scale = float(C) ** -0.5
  o, p = ops.attn_fwd(qkv, n_img, L, C, scale, save_p=save_p)
  q, k, v = (qkv[:, i * C:(i + 1) * C].float().reshape(n_img, L, C) for i in range(3))
  want_p = torch.softmax(torch.einsum('bic,bjc->bij', q, k) * scale, dim=-1)
  want_o = torch.einsum('bij,bjc->bic', want_p, v)
  _note(f'fused attention n_img={n_img}: o rel-L2 {rel_l2(o.float(), want_o):.2e}' + (f', p rel-L2 {rel_l2(p.float(), want_p):.2e}' if save_p else ''))
  assert rel_l2(o.float(), want_o) < 6e-3
  if save_p:
    assert p.shape == (n_img, L, L) and rel_l2(p.float(), want_p) < 4e-3
    assert torch.allclose(p.float().sum(-1), torch.ones(n_img, L, device=DEV), atol=2e-2)
  else:
    assert p is None


# ------------------------------------------------------------------------------------------------ bf16 trajectory
def test_bf16_loss_trajectory_100_steps_vs_reference(golden):
  """100 optimizer steps (batch 16, warm-up 0, dropout 0, lr 2e-4) of the bf16 path against the REFERENCE's fp32
  trajectory with the same replayed draws.  Both trajectories fall from ~1.0 as the (re-randomised) output layers
  are trained.  Tolerance: the batch-mean loss agrees within 3 % on each of the first 10 steps (measured 0.3 %), within
  5 % on average over the trajectory (measured 1.5 - 2.4 %) and within 60 % on every single step, and the bf16 run has
  learnt as much (mean of the last 10 steps within 10 %; measured 0.1 - 2 %).  bf16 rounding flips the sign of
  noise-level gradient entries, each of which Adam turns into a +-lr step, so the two PARAMETER trajectories separate
  slowly and chaotically while the losses keep tracking: the per-step maximum varies from run to run (16 - 35 %; the
  split-K reductions use fp32 atomics, so two runs of this very code already differ) and always falls on steps whose
  loss is 10 - 50x below the trajectory's mean (steps 37, 61: reference 0.24 / 0.17)."""
  from soft_truncation_b200 import losses
  g = golden('traj100_golden.npz')
  B, steps = int(g['B']), g['losses'].shape[0]
  cfg = _cfg(dropout=0.)
  cfg.optim.warmup = 0
  model, sde, _ = _model(cfg, int(g['seed']), torch.bfloat16)
  state = _state(cfg, model)
  step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
  batch = torch.rand(B, 3, 32, 32, generator=torch.Generator().manual_seed(1234)) * 2. - 1.
  assert abs(batch.double().sum().item() - float(g['batch_checksum'])) < 1e-6
  batch = batch.to(DEV)
  vp = ref_train.make_sde(cfg)
  got = []
  for s in range(steps):
    torch.manual_seed(200 + s)
    u, z = torch.rand(B), torch.randn(B, 3, 32, 32)
    inj = dict(u=u, z=z, t_min=vp.t_min_from_uniform(cfg, float(g['U'][s])))
    got.append(step_fn(state, batch, injected=inj).numpy())
  got, want = np.stack(got).mean(1), g['losses'].mean(1)
  rel = np.abs(got - want) / np.abs(want)
  worst = np.argsort(-rel)[:5]
  _note(f'bf16 100-step trajectory: max per-step rel diff {rel.max():.4f}, mean {rel.mean():.4f}, first 10 steps max '
        f'{rel[:10].max():.4f}; first/last reference loss {want[0]:.4f}/{want[-1]:.4f}, ours {got[0]:.4f}/{got[-1]:.4f}; '
        'worst steps: ' + ', '.join(f's{int(i)} ref {want[i]:.3f} ours {got[i]:.3f}' for i in worst))
  assert rel[:10].max() < 3e-2            # before the two Adam trajectories have had time to separate
  assert rel.max() < 0.6 and rel.mean() < 5e-2
  assert abs(got[-10:].mean() - want[-10:].mean()) < 0.1 * want[-10:].mean()


# ------------------------------------------------------------------------------------------------ bench-shape GEMMs
BENCH_SHAPES = [
    # B, H, Cin, Cout: the dominant layers of a B=512 step (profiles/r01_gemm_shapes.txt)
    (512, 32, 128, 128),      # wgrad M=128  N=1152 K=524288 (direct form, re-derived split-K), dgrad 128->128
    (512, 16, 256, 256),      # wgrad M=256  N=2304 K=131072 (transposed form, split_k 8)
    (512, 32, 256, 128),      # wgrad with a 256-channel input, 128 outputs
    (512, 8, 256, 256),
    (512, 4, 256, 256),
]


@pytest.mark.parametrize('case', BENCH_SHAPES)
def test_tcgen05_grads_at_bench_shapes_vs_simt(case):
  """The production tcgen05 weight-gradient (both forms, the split-K the dispatcher derives) and data-gradient
  (forward conv over the transposed weight copy) at the BENCH shapes, against gemm_simt_kernel on IDENTICAL bf16
  inputs: only the fp32 accumulation order differs -> rel-L2 <= 3e-3 (weights gradients sum 2^19 products)."""
  from soft_truncation_b200 import ops
  if not ops.tc_available():
    pytest.skip('tcgen05 backend unavailable on this device')
  B, H, Ci, Co = case
  gen = torch.Generator().manual_seed(H + Ci)
  x = (torch.randn(B, H, H, Ci, generator=gen)).to(DEV).to(torch.bfloat16)
  dy = (torch.randn(B, H, H, Co, generator=gen) * 0.05).to(DEV).to(torch.bfloat16)
  w = (torch.randn(Co, 9 * Ci, generator=gen) / math.sqrt(9 * Ci)).to(DEV).to(torch.bfloat16)
  bias, rb = torch.randn(Co, generator=gen).to(DEV), torch.randn(B, Co + 5, generator=gen).to(DEV)
  resid = torch.randn(B, H, H, Co, generator=gen).to(DEV).to(torch.bfloat16)
  res = {}
  try:
    for backend in ('simt', 'tcgen05'):
      ops.gemm_backend = backend
      dw = torch.zeros(Co, 9 * Ci, dtype=torch.float32, device=DEV)
      ops.conv_wgrad(dy, x, dw, 3, 3, alpha=0.7)
      dx = ops.conv_dgrad(dy, w, Ci, 3, 3)
      fwd = ops.conv_fwd(x, w, Co, 3, 3, bias=bias, rowbias=rb[:, 3:], rowbias_ld=rb.shape[1], residual=resid, alpha=0.7)
      res[backend] = (dw, dx.float(), fwd.float())
  finally:
    ops.gemm_backend = 'auto'
  _note(f'bench-shape tcgen05 vs simt {case}: ' + ', '.join(f'{n} {rel_l2(a, b):.2e}' for a, b, n in zip(res['tcgen05'], res['simt'], ('wgrad', 'dgrad', 'fwd'))))
  for got, want, name in zip(res['tcgen05'], res['simt'], ('wgrad', 'dgrad', 'fwd')):
    assert rel_l2(got, want) < (3e-3 if name == 'wgrad' else 6e-3), (name, rel_l2(got, want))


@pytest.mark.parametrize('case', [(256, 32, 128, 0, 128), (128, 32, 128, 128, 128), (256, 16, 256, 0, 256), (128, 16, 256, 256, 256),
                                  (32, 64, 128, 0, 128), (64, 32, 64, 0, 128), (256, 16, 256, 128, 256)])
def test_conv_halo_form_vs_plain_form(case):
  """The halo form of the 3x3 convolution (one (rows + 2)-row box feeds the three vertical taps through descriptor
  offsets, CTA pairs) against the plain one-box-per-tap form of the same kernel (ST_TC_HALO=0) and against the fp32-FMA
  kernel, on identical bf16 inputs, with bias + per-image bias + residual epilogues and channel-concatenated inputs:
  only the fp32 accumulation order differs -> rel-L2 <= 2e-5 between the two tensor-core forms (bf16 output rounding
  flips), <= 6e-3 against SIMT."""
  from soft_truncation_b200 import ops
  if not ops.tc_available():
    pytest.skip('tcgen05 backend unavailable on this device')
  B, H, C1, C2, Co = case
  gen = torch.Generator().manual_seed(B + H + C2)
  Ci = C1 + C2
  x1 = torch.randn(B, H, H, C1, generator=gen).to(DEV).to(torch.bfloat16)
  x2 = torch.randn(B, H, H, C2, generator=gen).to(DEV).to(torch.bfloat16) if C2 else None
  w = (torch.randn(Co, 9 * Ci, generator=gen) / math.sqrt(9 * Ci)).to(DEV).to(torch.bfloat16)
  bias, rb = torch.randn(Co, generator=gen).to(DEV), torch.randn(B, Co, generator=gen).to(DEV)
  resid = torch.randn(B, H, H, Co, generator=gen).to(DEV).to(torch.bfloat16)
  outs = {}
  try:
    for name, backend, halo in (('halo', 'tcgen05', '1'), ('plain', 'tcgen05', '0'), ('simt', 'simt', '1')):
      ops.gemm_backend = backend
      os.environ['ST_TC_HALO'] = halo
      a = ops.conv_fwd(x1, w, Co, 3, 3, x2=x2, bias=bias, rowbias=rb, rowbias_ld=Co)
      b = ops.conv_fwd(x1, w, Co, 3, 3, x2=x2, bias=bias, residual=resid, alpha=0.7)
      outs[name] = (a.float(), b.float())
  finally:
    ops.gemm_backend = 'auto'
    os.environ.pop('ST_TC_HALO', None)
  for i in range(2):
    _note(f'conv halo vs plain {case} [{i}]: {rel_l2(outs["halo"][i], outs["plain"][i]):.2e}; vs simt {rel_l2(outs["halo"][i], outs["simt"][i]):.2e}')
    assert rel_l2(outs['halo'][i], outs['plain'][i]) < 2e-3
    assert rel_l2(outs['halo'][i], outs['simt'][i]) < 6e-3


@pytest.mark.parametrize('mode', ['up2', 'down2', 'pre'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_upfirdn2d_fast_path_vs_generic_kernel(mode, dtype):
  """The register-blocked FIR kernel of the three network shapes (up 2 / pad (2,1), down 2 / pad (1,1), 1:1 / pad (2,2),
  4x4 taps, NHWC) against the generic kernel (which the reference fixture ops_golden.npz pins) on the same inputs, incl.
  odd image sizes: same taps, same fp32 FMA order per output up to association -> 1e-6 (fp32) / equal up to bf16
  rounding (bf16)."""
  from soft_truncation_b200 import ops
  kw = dict(up2=dict(up=2, pad=(2, 1)), down2=dict(down=2, pad=(1, 1)), pre=dict(pad=(2, 2)))[mode]
  k1 = np.asarray([1., 3., 3., 1.], dtype=np.float32)
  k = np.outer(k1, k1)
  k = torch.tensor(k / k.sum() * (4. if mode == 'up2' else 1.), device=DEV)
  k = k + 0.01 * torch.arange(16, device=DEV).float().view(4, 4)      # asymmetric: catches a flipped / transposed tap
  for shape in ((3, 16, 16, 64), (2, 10, 14, 32), (1, 64, 64, 128), (2, 7, 9, 8)):
    if mode == 'down2' and (shape[1] % 2 or shape[2] % 2):
      continue
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(sum(shape))).to(DEV).to(dtype)
    try:
      os.environ['ST_UPFIRDN_FAST'] = '1'
      fast = ops.upfirdn2d_nhwc(x, k, **kw)
      os.environ['ST_UPFIRDN_FAST'] = '0'
      ref = ops.upfirdn2d_nhwc(x, k, **kw)
    finally:
      os.environ.pop('ST_UPFIRDN_FAST', None)
    assert fast.shape == ref.shape
    assert rel_l2(fast.float(), ref.float()) < (1e-6 if dtype == torch.float32 else 3e-3), (mode, shape)


# ------------------------------------------------------------------------------------------------ reference checkpoint
def test_checkpoint_written_by_the_reference_restores_and_continues(golden, tmp_path):
  """tests/golden/ref_checkpoint.pth was written by the reference's utils.save_checkpoint from its own NCSNpp
  (DataParallel-wrapped), torch.optim.Adam and ExponentialMovingAverage after two optimizer steps.  This path restores
  it (utils.restore_checkpoint), takes the third step with the reference's draws and must reproduce the reference's
  third-step losses (rtol 5e-4), parameter and EMA norms (rtol 1e-4)."""
  from soft_truncation_b200 import losses, sde_lib, utils
  from soft_truncation_b200.models import utils as mutils
  g = golden('checkpoint_golden.npz')
  cfg = _cfg(dropout=0.)
  cfg.optim.warmup = 0
  cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks = 16, (1, 2), 1
  cfg.model.compute_dtype = 'fp32'
  sde = sde_lib.get_sde(cfg)
  torch.manual_seed(99)                                   # different initial weights: everything must come from the file
  model = mutils.create_model(cfg, sde)
  state = _state(cfg, model)
  state = utils.restore_checkpoint(cfg, os.path.join(GOLDEN, 'ref_checkpoint.pth'), state, torch.device(DEV))
  assert state['step'] == 2 and state['optimizer'].t == 2 and state['ema'].num_updates == 2
  net = mutils.unwrap(model)
  assert [k for k, _ in net.named_parameters()] == list(g['names'])
  step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
  vp = ref_train.make_sde(cfg)
  inj = dict(u=torch.tensor(g['u']), z=torch.tensor(g['z']), t_min=vp.t_min_from_uniform(cfg, float(g['U'])))
  got = step_fn(state, torch.tensor(g['batch'], device=DEV), injected=inj)
  np.testing.assert_allclose(got.numpy(), g['losses'][2], rtol=5e-4)
  assert state['step'] == int(g['step_after'])
  pn = np.array([p.double().norm().item() for p in net.parameters()])
  en = np.array([e.double().norm().item() for e in state['ema'].shadow_params])
  np.testing.assert_allclose(pn, g['pnorm'], rtol=1e-4, atol=1e-7)
  np.testing.assert_allclose(en, g['enorm'], rtol=1e-4, atol=1e-7)
  # and back: what this path writes, torch can load into plain containers of the reference's format
  path = str(tmp_path / 'out.pth')
  utils.save_checkpoint(cfg, path, state)
  raw = torch.load(path, map_location='cpu', weights_only=False)
  ref_raw = torch.load(os.path.join(GOLDEN, 'ref_checkpoint.pth'), map_location='cpu', weights_only=False)
  assert list(raw['model']) == list(ref_raw['model'])
  assert all(raw['model'][k].shape == ref_raw['model'][k].shape for k in raw['model'])
  assert set(raw['optimizer']['state'][0]) == set(ref_raw['optimizer']['state'][0])
  assert len(raw['ema']['shadow_params']) == len(ref_raw['ema']['shadow_params'])


# ------------------------------------------------------------------------------------------------ full-width C3 / C5
@pytest.mark.parametrize('tag,dtype', [('c3', torch.float32), ('c5', torch.float32), ('c3', torch.bfloat16), ('c5', torch.bfloat16)])
def test_full_width_c3_c5_single_image_vs_reference_fixture(golden, tag, dtype):
  """BASELINE configs[2] / [4] at FULL width and resolution (64x64 RVE with the FIR 'residual' input pyramid and the
  stride-2 im2col path; 256x256 VE with 7 levels, input_skip / output_skip pyramids, FIR at 128 / 256): score and
  training loss of one image against the reference.  fp32: score rel-L2 (on the fixture's strided sample) <= 5e-5, norm
  1e-4, loss rtol 5e-4; bf16: 3e-2 / 3e-2 / 3e-2."""
  from soft_truncation_b200 import configs, losses, _lib
  from soft_truncation_b200.models import utils as mutils
  g = golden('fullwidth_golden.npz')
  cfg = configs.celeba_uncsnpp_st() if tag == 'c3' else configs.celebahq_uncsnpp_st()
  cfg.device = torch.device(DEV)
  cfg.model.dropout = 0.
  seed = int(g[f'{tag}_seed'])
  model, sde, _ = _model(cfg, seed, dtype)
  net = mutils.unwrap(model)
  assert sum(p.numel() for p in net.parameters()) == int(g[f'{tag}_n_params'])
  R = cfg.data.image_size
  x = torch.rand(1, 3, R, R, generator=torch.Generator().manual_seed(seed + 100))
  assert abs(x.double().sum().item() - float(g[f'{tag}_x_checksum'])) < 1e-6
  _lib.lib.st_gemm_simt_fallbacks(1)
  model.eval()
  with torch.no_grad():
    score = model(x.to(DEV), torch.tensor(g[f'{tag}_sig'], device=DEV)).cpu()
  f32 = dtype == torch.float32
  idx = torch.tensor(g[f'{tag}_idx'])
  _note(f'full-width {tag} {dtype}: score sample rel-L2 {rel_l2(score.reshape(-1)[idx], g[f"{tag}_score_samples"]):.2e}')
  assert rel_l2(score.reshape(-1)[idx], g[f'{tag}_score_samples']) < (5e-5 if f32 else 3e-2)
  assert abs(score.double().norm().item() - float(g[f'{tag}_score_norm'])) < (1e-4 if f32 else 3e-2) * float(g[f'{tag}_score_norm'])
  torch.manual_seed(seed + 200)
  u, z = torch.rand(1), torch.randn(1, 3, R, R)
  assert abs(z.double().sum().item() - float(g[f'{tag}_z_checksum'])) < 1e-5
  loss_fn = losses.get_sde_loss_fn(cfg, sde, train=True)
  with torch.no_grad():
    ls = loss_fn(model, x.to(DEV), importance_sampling=cfg.training.importance_sampling, t_min=float(g[f'{tag}_tmin']),
                 injected=dict(u=u, z=z))
  np.testing.assert_allclose(ls.cpu().numpy(), g[f'{tag}_losses'], rtol=5e-4 if f32 else 3e-2)
  if not f32:
    n = int(_lib.lib.st_gemm_simt_fallbacks(0))
    assert n == 0, f'{n} bf16 GEMMs fell back to the SIMT kernel: {_lib.lib.st_gemm_simt_fallback_reason().decode()}'


def test_no_simt_fallback_in_a_full_size_training_step():
  """The counter behind bench.py's assertion: a full-size bf16 CIFAR-10 step (forward, backward, optimizer) issues no
  GEMM on the fp32-FMA kernel; an fp32-operand problem is not counted; a bf16 problem the tensor path cannot express is."""
  from soft_truncation_b200 import _lib, losses, ops
  if not ops.tc_available():
    pytest.skip('tcgen05 backend unavailable on this device')
  cfg = _cfg(dropout=0.1)
  model, sde, _ = _model(cfg, 3, torch.bfloat16)
  state = _state(cfg, model)
  step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
  batch = (torch.rand(16, 3, 32, 32, generator=torch.Generator().manual_seed(1)) * 2 - 1).to(DEV)
  _lib.lib.st_gemm_simt_fallbacks(1)
  assert torch.isfinite(step_fn(state, batch)).all()
  assert _lib.lib.st_gemm_simt_fallbacks(0) == 0, _lib.lib.st_gemm_simt_fallback_reason()
  a = torch.randn(8, 20, device=DEV).to(torch.bfloat16)       # K = 20: leading dimension not a multiple of 8
  ops.gemm_nt(a, a)
  assert _lib.lib.st_gemm_simt_fallbacks(1) == 1 and _lib.lib.st_gemm_simt_fallback_reason()
  ops.gemm_nt(a.float(), a.float())
  assert _lib.lib.st_gemm_simt_fallbacks(1) == 0


# ------------------------------------------------------------------------------------------------ optimizers
def test_adamw_step_through_step_fn_updates_parameters():
  """The non-fused optimizer path (optim.optimizer='AdamW' -> torch.optim.AdamW over the parameter views): the explicit
  backward writes the flat gradient buffer that the .grad views alias, zero_grad goes through the model (torch's
  set_to_none would hide the gradients from the optimizer), so one step must move the weights, and the gradient must
  not accumulate across steps."""
  from soft_truncation_b200 import losses
  from soft_truncation_b200.models import utils as mutils
  cfg = _small(_cfg(dropout=0.))
  cfg.optim.optimizer, cfg.optim.warmup = 'AdamW', 0
  model, sde, _ = _model(cfg, 5, torch.float32)
  state = _state(cfg, model)
  assert isinstance(state['optimizer'], torch.optim.AdamW)
  step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
  net = mutils.unwrap(model)
  before = net._flat.clone()
  batch = (torch.rand(4, 3, 32, 32, generator=torch.Generator().manual_seed(1)) * 2 - 1).to(DEV)
  inj = dict(t_min=1e-3, u=torch.rand(4, generator=torch.Generator().manual_seed(2)),
             z=torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(3)))
  step_fn(state, batch, injected=inj)
  moved = (net._flat - before).abs()
  assert float(moved.max()) > 1e-5 and all(p.grad is not None for p in net.parameters() if p.requires_grad)
  g1 = net._grad.clone()
  with torch.no_grad():
    net._flat.copy_(before)
  net.sync_compute_weights()
  step_fn(state, batch, injected=inj)
  assert rel_l2(net._grad, g1) < 1e-4          # same inputs, same weights: same gradient, not twice it


# ------------------------------------------------------------------------------------------------ reference drivers
_DRIVER_SCRIPT = r'''
import sys, os, json, tempfile
sys.path.insert(0, {root!r})
import numpy as np, torch
from baseline import ref_env
ref_root = ref_env.locate(allow_source=True)
ref_env.install_shims(ref_root, drivers=True)
# INTEGRATION.md section 1: redirect the reference's module names to this package, then import its drivers UNCHANGED
import soft_truncation_b200 as st
from soft_truncation_b200 import sde_lib, losses, sampling, likelihood, datasets, op, models
from soft_truncation_b200.models import utils as mutils, ncsnpp, ema
sys.modules.update({{'sde_lib': sde_lib, 'losses': losses, 'sampling': sampling, 'op': op, 'models': models,
                    'models.utils': mutils, 'models.ncsnpp': ncsnpp, 'models.ema': ema, 'likelihood': likelihood,
                    'datasets': datasets}})
sys.path.insert(0, ref_root)
import utils as ref_utils                      # the reference's utils.py (load_model, get_loss_fns, checkpoints)
assert os.path.samefile(os.path.dirname(ref_utils.__file__), ref_root)
from soft_truncation_b200 import configs
cfg = configs.cifar10_ddpmpp_nll_st()
cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks = 128, (1, 2), 1
cfg.optim.warmup = 0
cfg.sampling.method, cfg.sampling.batch_size = 'pc', 4
cfg.model.num_scales = 1000
dev = torch.device({dev!r})
cfg.device = dev
work = tempfile.mkdtemp()
sde = sde_lib.get_sde(cfg, None)
state, score_model, ema_, ckpt_dir, ckpt_meta = ref_utils.load_model(cfg, work, sde=sde)
inverse_scaler = datasets.get_data_inverse_scaler(cfg)
train_step_fn, nll_fn, nelbo_fn, sampling_fn = ref_utils.get_loss_fns(cfg, sde, inverse_scaler)
out = dict(model=type(mutils.unwrap(score_model)).__module__, opt=type(state['optimizer']).__name__)
if dev.type == 'cuda':
  batch = (torch.rand(4, 3, 32, 32) * 2 - 1).to(dev)
  np.random.seed(0)
  l = [train_step_fn(state, batch) for _ in range(2)]
  out.update(loss_finite=bool(torch.isfinite(torch.stack(l)).all()), loss_device=l[0].device.type, step=state['step'])
  ref_utils.save_checkpoint(cfg, ckpt_meta, state)
  state2, *_ = ref_utils.load_model(cfg, work, sde=sde)       # restore_checkpoint through the reference's own code
  out.update(restored_step=state2['step'],
             same_weights=bool(torch.equal(mutils.unwrap(state2['model'])._flat, mutils.unwrap(state['model'])._flat)))
print('RESULT ' + json.dumps(out))
'''


def test_reference_drivers_run_on_this_package():
  """north_star: "so main.py drives it unchanged".  The reference's utils.py (load_model, get_loss_fns, save / restore
  checkpoint - what run_lib.train calls, run_lib.py:54,65,82-89) is imported UNCHANGED from the staged reference with
  its module names redirected to this package and TensorFlow stubbed; it builds the model,
  takes two optimizer steps on the GPU, writes and restores a checkpoint."""
  from baseline import ref_env
  if ref_env.locate(allow_source=True) is None:
    pytest.skip('the reference is not staged under baseline/_ref')
  r = subprocess.run([sys.executable, '-c', _DRIVER_SCRIPT.format(root=ROOT, dev=DEV)], capture_output=True, text=True, timeout=600)
  assert r.returncode == 0, r.stderr[-3000:]
  import json
  res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')][-1][7:])
  assert res['model'] == 'soft_truncation_b200.models.ncsnpp' and res['opt'] == 'FusedAdam'
  assert res['loss_finite'] and res['loss_device'] == 'cpu' and res['step'] == 2
  assert res['restored_step'] == 2 and res['same_weights']
