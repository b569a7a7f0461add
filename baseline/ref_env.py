"""Locates, stages and imports the UNTOUCHED reference (Kim-Dongjun/Soft-Truncation) for the baseline arms.

The reference is ~10 kLoC of pure Python with no setup.py / pyproject.toml, so `pip install --target baseline/_ref`
does not apply; `stage()` instead copies its source tree verbatim from /root/reference into `baseline/_ref/`
(git-ignored, NOT gpurun-ignored: it travels to the GPU box with the snapshot).  Nothing under
`baseline/_ref/` is product source and the product package never imports it: it is used by

  * `bench.py --impl reference`      the reference's own `losses.get_step_fn` on the host cores,
  * `bench.py` (`gpu_reference`)     the same stock code path on cuda:0 (eager fp32 / bf16 autocast + channels_last),
  * `tests/golden/make_golden.py`    fixture generation in the build container.

Modules the reference imports at module scope but that are absent from this image are replaced by import shims:
`ml_collections` (-> soft_truncation_b200.config_dict.ConfigDict), `op` (-> the reference's own CPU function
`upfirdn2d_native`, extracted from op/upfirdn2d.py by AST so that importing it does not JIT-build the CUDA
extension for two minutes; the CIFAR-10 configs have fir=False and never call it), and for the drivers
(`utils.py`, `run_lib.py`): `tensorflow.io.gfile`, `tensorflow_gan`, `tensorflow_hub`, `tensorflow_datasets`, `natsort`.
"""
import ast
import glob as _glob
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGED = os.path.join(HERE, '_ref')
SOURCE = os.environ.get('ST_REFERENCE', '/root/reference')


def stage(force=False):
  """Copy the reference tree into baseline/_ref (only where /root/reference exists, i.e. the build container)."""
  if not os.path.isdir(SOURCE):
    return STAGED if os.path.isdir(STAGED) else None
  if os.path.isdir(STAGED) and not force:
    return STAGED
  if os.path.isdir(STAGED):
    shutil.rmtree(STAGED)
  shutil.copytree(SOURCE, STAGED, ignore=shutil.ignore_patterns('.git', '__pycache__', 'figure', '*.pyc', '.SUBMODULES.json'))
  return STAGED


def locate(allow_source=False):
  """Path of the reference tree: the staged copy, or (build container only) /root/reference itself."""
  if os.path.isfile(os.path.join(STAGED, 'losses.py')):
    return STAGED
  if allow_source and os.path.isfile(os.path.join(SOURCE, 'losses.py')):
    return SOURCE
  return None


def _gfile_stub():
  gfile = types.SimpleNamespace(
      exists=os.path.exists, makedirs=lambda p: os.makedirs(p, exist_ok=True), glob=_glob.glob,
      isdir=os.path.isdir, listdir=os.listdir, remove=os.remove, rmtree=shutil.rmtree,
      GFile=lambda name, mode='r': open(name, mode))
  return gfile


def _stub(name):
  """Empty module with a ModuleSpec (torch._dynamo probes importlib.util.find_spec('tensorflow') lazily and a
  spec-less entry in sys.modules makes that raise)."""
  import importlib.machinery
  mod = types.ModuleType(name)
  mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
  return mod


def install_shims(ref_root, drivers=False):
  """Import shims for modules that are absent from this image (see the module docstring)."""
  import torch
  import torch.nn.functional as F
  if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
  from soft_truncation_b200.config_dict import ConfigDict
  if 'ml_collections' not in sys.modules:
    shim = _stub('ml_collections')
    shim.ConfigDict = ConfigDict
    sys.modules['ml_collections'] = shim
  if 'op' not in sys.modules:
    ns = {'torch': torch, 'F': F}
    tree = ast.parse(open(os.path.join(ref_root, 'op', 'upfirdn2d.py')).read())
    for node in tree.body:
      if isinstance(node, ast.FunctionDef) and node.name == 'upfirdn2d_native':
        exec(compile(ast.Module([node], []), 'op/upfirdn2d.py', 'exec'), ns)
    native = ns['upfirdn2d_native']

    def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
      return native(input, kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])

    def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
      rest = [1] * (input.ndim - bias.ndim - 1)       # CPU branch of op/fused_act.py:87-94 (slope hard-coded there)
      return F.leaky_relu(input + bias.view(1, bias.shape[0], *rest), negative_slope=0.2) * scale

    op = _stub('op')
    op.upfirdn2d, op.upfirdn2d_native, op.fused_leaky_relu, op.FusedLeakyReLU = upfirdn2d, native, fused_leaky_relu, None
    sys.modules['op'] = op
  if drivers:
    if 'tensorflow' not in sys.modules:
      tf = _stub('tensorflow')
      tf.io = types.SimpleNamespace(gfile=_gfile_stub())
      sys.modules['tensorflow'] = tf
    for name in ('tensorflow_gan', 'tensorflow_hub', 'tensorflow_datasets', 'natsort'):
      if name not in sys.modules:
        sys.modules[name] = _stub(name)


def import_reference(ref_root=None, allow_source=False, drivers=False):
  """Import the reference's hot-path modules (and, with drivers=True, its `utils.py` glue) from `ref_root`."""
  ref_root = ref_root or locate(allow_source=allow_source)
  if ref_root is None:
    raise ImportError('the reference is not staged under baseline/_ref (run baseline/ref_env.stage() in the build container)')
  install_shims(ref_root, drivers=drivers)
  if ref_root not in sys.path:
    sys.path.insert(0, ref_root)
  import losses, sampling, sde_lib  # noqa: E401
  from models import ema, ncsnpp, utils as mutils
  ns = types.SimpleNamespace(root=ref_root, sde_lib=sde_lib, losses=losses, sampling=sampling, ncsnpp=ncsnpp, mutils=mutils,
                             ema=ema, op=sys.modules['op'])
  if drivers:
    import likelihood
    import utils as ref_utils
    ns.likelihood, ns.utils = likelihood, ref_utils
  return ns


def ref_config(path, device=None):
  """`configs/<path>.py:get_config()` of the reference, e.g. 'vp/CIFAR10/ddpmpp_nll_st'."""
  import importlib
  import torch
  cfg = importlib.import_module('configs.' + path.replace('/', '.')).get_config()
  cfg.device = torch.device(device) if device is not None else torch.device('cpu')
  return cfg


if __name__ == '__main__':
  print(stage(force='--force' in sys.argv))
