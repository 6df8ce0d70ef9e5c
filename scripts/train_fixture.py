#!/usr/bin/env python
"""Train a qualitative-world checkpoint on the GPU with the repo's OWN training step (no reference code involved), then
sample the held-out pool and report the solved rate from the GPU checker.

    python scripts/train_fixture.py --steps 3000 --out gpurun_out/ckpt

Training pool: the committed RandomSplitQualitativeWorld fixture scenes (N = 3, 4, 6 and the SECOND half of the N = 8 pool);
evaluation: the FIRST 256 scenes of the N = 8 pool (disjoint).  Prints a JSON line with the loss curve and the solved rates.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_ccsp_b200 import scenes, synthetic  # noqa: E402
from diffusion_ccsp_b200.checker import SolvedChecker  # noqa: E402
from diffusion_ccsp_b200.trainer import create_trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=3000)
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--timesteps', type=int, default=1000)
    ap.add_argument('--eval-every', type=int, default=1000)
    ap.add_argument('--eval-scenes', type=int, default=256)
    ap.add_argument('--out', default='gpurun_out/ckpt')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--init', default=None, help='state dict (.pt) to start from')
    ap.add_argument('--lr', type=float, default=5e-4)
    ap.add_argument('--max-tiles', type=int, default=8, help='train only on scenes with at most this many tiles (the reference trains on 2-5)')
    ap.add_argument('--min-tiles', type=int, default=2, help='train only on scenes with at least this many tiles')
    ap.add_argument('--cosine', action='store_true', help='cosine decay of the learning rate to 2 % of --lr over --steps (the reference trains at a constant rate)')
    a = ap.parse_args()
    torch.manual_seed(a.seed)
    train_pool = scenes.qualitative_train_pool()                      # 24 000 scenes, 2..8 tiles (tests/golden/make_train_pool.py)
    if a.max_tiles < 8 or a.min_tiles > 2:
        off = train_pool.scene_node_ranges()
        tiles = np.diff(off) - 1
        keep = np.where((tiles <= a.max_tiles) & (tiles >= a.min_tiles))[0]
        train_pool = scenes._gather_scenes_fast(train_pool, keep)
        print(f'[fixture] training on {train_pool.num_graphs} scenes with {a.min_tiles}..{a.max_tiles} tiles', flush=True)
    eval_batch = scenes.qualitative_batch(a.eval_scenes, 8)           # evaluation fixtures: disjoint draws
    eval4 = scenes.qualitative_batch(64, 4)
    dims = synthetic.DIMS['qualitative']
    sd = synthetic.make_state_dict(dims, 'qualitative', seed=0)
    tr = create_trainer('qualitative', timesteps=a.timesteps, EBM='ULA', train_dataset=train_pool, train_num_steps=0,
                        train_batch_size=a.batch, train_lr=a.lr, save_and_sample_every=10 ** 9, results_folder=a.out, render_dir=a.out, device='cuda')
    if a.init:
        sd = {k: v.float() for k, v in torch.load(a.init, map_location='cpu').items()}
    tr.model.load_state_dict(sd, strict=False)
    checker = SolvedChecker(eval_batch, dims, 'qualitative', 'cuda')
    checker4 = SolvedChecker(eval4, dims, 'qualitative', 'cuda')
    log = dict(steps=[], loss=[], solved=[], train_s=[], config=vars(a))
    t_train = 0.0
    done = 0
    best = None
    while done < a.steps:
        n = min(a.eval_every, a.steps - done)
        tr.train_num_steps = done + n
        torch.cuda.synchronize(); t0 = time.time()
        if a.cosine:
            import math
            total = a.steps
            tr.lr_schedule = lambda step: a.lr * (0.02 + 0.98 * 0.5 * (1 + math.cos(math.pi * min(step, total) / total)))
        tr.train(log_every=max(n // 4, 1), evaluate=False)
        torch.cuda.synchronize(); t_train += time.time() - t0
        done += n
        tr.model.eval()
        poses = tr.model.sample(eval_batch, seed=1)
        solved, counts = checker(poses, return_counts=True)
        frac = float(solved.float().mean())
        frac4 = float(checker4(tr.model.sample(eval4, seed=2)).float().mean())
        free = poses[~eval_batch.mask.bool().cuda()]
        finite = float((counts[:, 0] >= 0).float().mean())
        print(f'[fixture] step {done}: loss {tr.loss_log[-1][1]:.5f}  solved N=8 {frac:.3f} N=4 {frac4:.3f}  finite scenes N=8 {finite:.3f}  '
              f'collisions {(counts[:, 0] > 0).float().mean():.3f} missing {(counts[:, 1] > 0).float().mean():.3f}  '
              f'max|x| {float(free[torch.isfinite(free)].abs().max()) if bool(torch.isfinite(free).any()) else float("nan"):.2f}  train {t_train:.1f}s', flush=True)
        log['steps'].append(done); log['loss'].append(tr.loss_log[-1][1]); log['solved'].append(frac); log.setdefault('solved_n4', []).append(frac4)
        log.setdefault('finite_n8', []).append(finite); log['train_s'].append(t_train)
        score = (finite >= 0.99, frac4 + frac)
        if best is None or score > best[0]:
            best = (score, done, {k: v.detach().half().cpu() for k, v in tr.model.state_dict().items() if k.startswith('denoise_fn.')})
    os.makedirs(a.out, exist_ok=True)
    (fin_ok, sc), best_step, half = best
    print(f'[fixture] keeping the checkpoint of step {best_step} (N=8 finite: {fin_ok}, solved N=4 + N=8: {sc:.3f})')
    np.savez_compressed(os.path.join(a.out, 'denoise_fn_fp16.npz'), trained_steps=best_step, **{k: v.numpy() for k, v in half.items()})
    # the kept checkpoint under both samplers (ULA K=10 = the headline sampler; plain DDPM = EBM False)
    from diffusion_ccsp_b200.ddpm import GaussianDiffusion
    from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser
    report = {}
    for ebm in ('ULA', False):
        den = ConstraintDiffuser(dims=dims, input_mode='qualitative', device='cuda', verbose=False, math='bf16x3')
        gd = GaussianDiffusion(den, timesteps=a.timesteps, EBM=ebm, samples_per_step=10).eval()
        gd.load_state_dict({k: v.float() for k, v in half.items()}, strict=False)
        for name, b, ck in (('N=8', eval_batch, checker), ('N=4', eval4, checker4)):
            tries = []
            ok_any = torch.zeros(b.num_graphs, dtype=torch.bool, device='cuda')
            for k in range(3):
                s_k, c_k = ck(gd.sample(b, seed=100 + k), return_counts=True)
                ok_any |= s_k
                tries.append(float(s_k.float().mean()))
            report[f'{ebm}/{name}'] = dict(solved_per_try=tries, solved_top3=float(ok_any.float().mean()), finite=float((c_k[:, 0] >= 0).float().mean()))
            print(f'[fixture] sampler {ebm} {name}: solved per try {tries}  top-3 {float(ok_any.float().mean()):.3f}', flush=True)
    log['report'] = report
    log['kept_step'] = best_step
    log['loss_log'] = tr.loss_log
    log['ms_per_step'] = t_train / max(done, 1) * 1e3
    with open(os.path.join(a.out, 'train_log.json'), 'w') as f:
        json.dump(log, f)
    print(json.dumps(dict(steps=done, kept_step=best_step, final_loss=log['loss'][-1], solved=log['solved'], solved_n4=log['solved_n4'],
                          ms_per_train_step=log['ms_per_step'], report=report)))


if __name__ == '__main__':
    main()
