"""Checkpoint interchange with the reference format (reference utils.py:13-36).

A checkpoint is `torch.save({'optimizer', 'model', 'ema', 'step'})` where `model` is the `module.`-prefixed state_dict
with the reference's logical tensor shapes (OIHW convolutions, (in, out) NIN matrices), `optimizer` is a
`torch.optim.Adam` state_dict (per-parameter `step` / `exp_avg` / `exp_avg_sq` in `model.parameters()` order) and `ema`
is `{'decay', 'num_updates', 'shadow_params'}`.  The flat-buffer model, FusedAdam and the flat EMA shadow of this
package read and write exactly that, so checkpoints move between the reference and this path in both directions."""
import logging
import os

import torch


def restore_checkpoint(config, ckpt_dir, state, device):
  """Load `ckpt_dir` (a file path, as in the reference) into `state`; a missing file returns `state` unchanged."""
  if not os.path.exists(ckpt_dir):
    os.makedirs(os.path.dirname(ckpt_dir) or '.', exist_ok=True)
    logging.warning(f"No checkpoint found at {ckpt_dir}. Returned the same state as input")
    return state
  logging.info(ckpt_dir + ' loaded ...')
  loaded_state = torch.load(ckpt_dir, map_location=device, weights_only=False)
  state['optimizer'].load_state_dict(loaded_state['optimizer'])
  state['model'].load_state_dict(loaded_state['model'], strict=False)
  state['ema'].load_state_dict(loaded_state['ema'])
  state['step'] = loaded_state['step']
  return state


def save_checkpoint(config, ckpt_dir, state):
  saved_state = {
    'optimizer': state['optimizer'].state_dict(),
    'model': state['model'].state_dict(),
    'ema': state['ema'].state_dict(),
    'step': state['step']
  }
  torch.save(saved_state, ckpt_dir)
