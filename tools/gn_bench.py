"""GroupNorm kernels on the shapes of the bench workload: CUDA-event time per launch and achieved HBM GB/s against the
algorithmic bytes (forward stats: read x; apply: read x, write y; backward reduce: read x, dy; backward apply: read
x, dy, write dx)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soft_truncation_b200 import ops  # noqa: E402

DEV, BF = torch.device('cuda:0'), torch.bfloat16
B = int(os.environ.get('GB_BATCH', '512'))
ITERS = 20
SHAPES = [(32, 128, 0), (32, 128, 128), (16, 256, 0), (16, 256, 128), (16, 256, 256), (8, 256, 0), (8, 256, 256), (4, 256, 0), (4, 256, 256)]


def timed(fn):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
  tot = 0.
  for _ in range(ITERS):
    flush.sum()                         # evict the previous iteration's tensors from the 126 MB L2 with a READ-only
                                        # pass (a memset would leave 126 MB of dirty lines to write back under the kernel)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    tot += a.elapsed_time(b)
  return tot / ITERS * 1e3   # us


def main():
  print(f'B={B} bf16; us per launch (GB/s of algorithmic bytes)')
  print('H C1 C2 | stats | apply | apply+drop | bwd_reduce | bwd_apply | bwd_reduce+drop | bwd_apply+drop+csum')
  for H, C1, C2 in SHAPES:
    Ct = C1 + C2
    G = min(Ct // 4, 32)
    x = torch.randn(B, H, H, C1, device=DEV).to(BF)
    x2 = torch.randn(B, H, H, C2, device=DEV).to(BF) if C2 else None
    dy = torch.randn(B, H, H, Ct, device=DEV).to(BF)
    gamma, beta = torch.randn(Ct, device=DEV), torch.randn(Ct, device=DEV)
    dgamma, dbeta = torch.zeros(Ct, device=DEV), torch.zeros(Ct, device=DEV)
    bits = torch.empty(B * H * H * Ct // 8, dtype=torch.uint8, device=DEV)
    nbytes = B * H * H * Ct * 2
    stats = ops.gn_stats(x, x2, G)
    ops.gn_apply(x, x2, G, gamma, beta, stats, True, p_drop=0.1, seed=5, keepbits=bits)
    orig_red, orig_app = ops.lib.st_gn_bwd_reduce, ops.lib.st_gn_bwd_apply

    def only(which, **kw):
      # time one of the two backward passes by stubbing the other out
      def run():
        ops.gn_backward(x, x2, dy, G, gamma, beta, stats, True, dgamma, dbeta, fused_chunks=0, **kw)
      return run

    t_stats = timed(lambda: ops.gn_stats(x, x2, G))
    t_apply = timed(lambda: ops.gn_apply(x, x2, G, gamma, beta, stats, True))
    t_applyd = timed(lambda: ops.gn_apply(x, x2, G, gamma, beta, stats, True, p_drop=0.1, seed=5, keepbits=bits))
    res = []
    for kw in (dict(), dict(p_drop=0.1, seed=5, keepbits=bits, want_csum=True)):
      t_all = timed(only('all', **kw))
      ops.lib.st_gn_bwd_apply = lambda *a: 0
      t_red = timed(only('reduce', **kw))
      ops.lib.st_gn_bwd_apply = orig_app
      res.append((t_red, t_all - t_red))
    # single-launch cluster form (cluster size fc) against the two-kernel total, parameter gradients deferred
    q = ops.ColsumQueue()
    fused = []
    for kw in (dict(), dict(p_drop=0.1, seed=5, keepbits=bits, want_csum=True), dict(extra=dy, extra_scale=0.7, want_csum=True)):
      row = []
      for fc in (0, 1, 2, 4, 8, 16):
        if fc > max(1, H * H // (256 // (Ct // 8))):
          continue

        def run(fc=fc, kw=kw):
          ops.gn_backward(x, x2, dy, G, gamma, beta, stats, True, dgamma, dbeta, queue=q, fused_chunks=fc, **kw)
          q.jobs, q.keep = [], []
        row.append(f'{fc}:{timed(run):.1f}')
      fused.append(' '.join(row))
    # forward: st_gn_stats + st_gn_apply against the cluster-resident single launch
    fwd = []
    for kw in (dict(), dict(p_drop=0.1, seed=5, keepbits=bits)):
      t2 = timed(lambda: ops.gn_norm_act(x, x2, G, gamma, beta, True, fused_chunks=0, **kw))
      fcs = ops.lib.st_gn_fwd_fused_chunks(B, H * H, Ct)
      t1 = timed(lambda: ops.gn_norm_act(x, x2, G, gamma, beta, True, fused_chunks=fcs, **kw)) if fcs > 0 else float('nan')
      fwd.append(f'two kernels {t2:.1f}, resident x{fcs} {t1:.1f}')
    # backward with x / dy resident in shared memory (2-stream form): plain and dropout + column sums
    rc = ops.lib.st_gn_bwd_resident_chunks(B, H * H, Ct, 2)
    resid = []
    for kw in (dict(), dict(p_drop=0.1, seed=5, keepbits=bits, want_csum=True)):
      def run_r(kw=kw):
        ops.gn_backward(x, x2, dy, G, gamma, beta, stats, True, dgamma, dbeta, queue=q, resident=rc, **kw)
        q.jobs, q.keep = [], []
      resid.append(f'{timed(run_r):.1f}' if rc > 0 else 'n/a')
    f = lambda t, units: f'{t:7.1f} ({units * nbytes / t / 1e3:5.0f})'
    print(H, C1, C2, '|', f(t_stats, 1), '|', f(t_apply, 2), '|', f(t_applyd, 2), '|', f(res[0][0], 2), '|', f(res[0][1], 3), '|',
          f(res[1][0], 2), '|', f(res[1][1], 3), flush=True)
    print('      backward total us by cluster size (0 = two kernels): plain [', fused[0], '] drop+csum [', fused[1], '] extra+csum [', fused[2], ']',
          f'ideal 3-pass at 6.5 TB/s: {3 * nbytes / 6.5e6:.1f}', flush=True)
    rc3 = ops.lib.st_gn_bwd_resident_chunks(B, H * H, Ct, 3)

    def run_r3():
      ops.gn_backward(x, x2, dy, G, gamma, beta, stats, True, dgamma, dbeta, queue=q, resident=rc3, extra=dy, extra_scale=0.7,
                      want_csum=True)
      q.jobs, q.keep = [], []
    r3 = f'{timed(run_r3):.1f}' if rc3 > 0 else 'n/a'
    print(f'      backward resident x{rc}: plain {resid[0]}, drop+csum {resid[1]}; with extra (x{rc3}) + csum {r3}', flush=True)
    print('      forward total us: plain [', fwd[0], '] dropout [', fwd[1], f'] ideal 2-pass at 6.5 TB/s: {2 * nbytes / 6.5e6:.1f}', flush=True)


if __name__ == '__main__':
  main()
