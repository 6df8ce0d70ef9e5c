"""Trainer / load_trainer / evaluate_model — the callers of the sampling path, mirrored at the API level.

The reference's Trainer (networks/ddpm.py:394-904) also owns training, rendering, wandb and the CPU
success checker (trimesh + python-fcl); all of that is OUT OF SCOPE for this repo.
What is kept is what a user of `solve_csp.evaluate_model` needs around the accelerated path:
  * `Trainer.model` IS the GaussianDiffusion (train_utils.py:299-300) and `Trainer.load/save` use the
    reference's checkpoint layout `logs/<run>/model-<milestone>.pt` -> {'step', 'model': state_dict}
    (ddpm.py:496-514);
  * `Trainer.evaluate` iterates scene batches, calls `self.model.sample(batch, ...)` `tries` times
    (ddpm.py:607-614), clamps to [-1, 1] and re-assembles full feature rows (ddpm.py:620, 807-821), and
    logs `model_ave_sample_time` (ddpm.py:830-836).  A `checker(world_rows, batch) -> list[bool]` callback
    stands in for the CPU constraint checker.
"""
from __future__ import annotations

import json
import os
import time
from pathlib import Path
from typing import Callable, Dict, Iterable, Optional, Sequence

import torch

from .ddpm import GaussianDiffusion
from .denoise_fn import ConstraintDiffuser
from .scenes import SceneBatch
from .synthetic import dims_for


class Trainer(object):
    def __init__(self, denoise_fn: GaussianDiffusion, train_dataset=None, test_datasets: Optional[Dict] = None,
                 render_dir: str = './renders', *, results_folder: str = './results', EBM=False, eval_only=True,
                 input_mode=None, **kwargs):
        self.model = denoise_fn
        self.dims = denoise_fn.dims
        self.input_mode = denoise_fn.input_mode
        self.EBM = EBM
        self.test_datasets = test_datasets or {}      # {n_objects: iterable of SceneBatch}
        self.eval_kwargs = dict(tries=(10, 0))
        self.render_dir = render_dir
        self.results_folder = Path(results_folder)
        self.step = 0

    # ---- checkpoints (ddpm.py:496-514) --------------------------------------------------------------
    def save(self, milestone):
        self.results_folder.mkdir(parents=True, exist_ok=True)
        torch.save({'step': self.step, 'model': self.model.state_dict()}, str(self.results_folder / f'model-{milestone}.pt'))

    def load(self, milestone):
        data = torch.load(str(self.results_folder / f'model-{milestone}.pt'), map_location='cpu')
        self.step = data['step']
        self.model.load_state_dict(data['model'])

    # ---- ddpm.py:807-821 ----------------------------------------------------------------------------
    def get_all_features(self, all_features, batch):
        return torch.cat([batch.x[:, :self.dims[-1][1]].cpu(), all_features.detach().cpu(),
                          batch.x[:, self.dims[-1][2]:].cpu()], dim=1)

    def train(self):
        raise NotImplementedError('training is the "next" row N2 of SURVEY.md §8f')

    # ---- ddpm.py:558-805, sampling + bookkeeping only -------------------------------------------------
    def evaluate(self, json_name='eval', tries=(10, 0), render=False, save_log=True, run_all=False, run_only=False,
                 resume_eval=False, return_history=False, checker: Optional[Callable] = None, **kwargs):
        """ddpm.py:558-805 — sampling, success accounting and the JSON log.  The per-graph CPU check of ddpm.py:633-713
        (trimesh + python-fcl) runs on the GPU for the 2-D box worlds (`checker.SolvedChecker`, one launch per sample,
        poses stay on the device); other worlds need a `checker(rows, batch) -> list[bool]` callback (none = success
        rates are reported as None)."""
        assert not self.model.training, 'call .eval() first (ddpm.py:328)'
        from .checker import SolvedChecker, world_kind_for
        use_gpu_checker = checker is None and world_kind_for(self.input_mode) is not None
        log = {}
        for i, batches in self.test_datasets.items():
            count = 0
            solved = set()
            first_round = {}
            self.model.sample_loop_time = []
            t0 = time.time()
            for data in batches:
                base = count
                count += data.num_graphs
                gpu_checker = SolvedChecker(data, self.dims, self.input_mode, self.model.denoise_fn.device) if use_gpu_checker else None
                for k in range(tries[0]):
                    batch = data.clone()                                            # ddpm.py:607
                    result = self.model.sample(batch, return_history=return_history)  # ddpm.py:612  <- the hot path
                    poses = result[0] if return_history else result
                    if gpu_checker is not None:
                        flags = gpu_checker(poses).cpu().tolist()                   # clamp (ddpm.py:620) + check, on the device
                    elif checker is not None:
                        rows = self.get_all_features(poses.clamp(-1, 1), batch)     # ddpm.py:620-621
                        flags = checker(rows, batch)
                    else:
                        flags = []
                    for j, ok in enumerate(flags):
                        if ok and (base + j) not in solved:
                            solved.add(base + j); first_round[base + j] = k
                    if len(solved) == count and not run_all:
                        break
            n_samples = max(len(self.model.sample_loop_time), 1)
            checked = use_gpu_checker or checker is not None
            log[i] = {
                'success_rate': round(len([s for s in first_round.values() if s == 0]) / max(count, 1), 3) if checked else None,
                'success_rate_top3': round(len(solved) / max(count, 1), 3) if checked else None,
                'success_rounds': {str(k): v for k, v in sorted(first_round.items())} if checked else None,
                'model_ave_sample_time': sum(self.model.sample_loop_time) / n_samples / max(count, 1),   # ddpm.py:830
                'scenes': count, 'wall_time': time.time() - t0,
            }
        if save_log:
            os.makedirs(self.render_dir, exist_ok=True)
            with open(os.path.join(self.render_dir, f'denoised_{json_name}.json'), 'w') as f:
                json.dump(log, f, indent=2)
        return log


def create_trainer(input_mode='qualitative', timesteps=1000, EBM='ULA', samples_per_step=10, step_sizes='2*self.betas',
                   hidden_dim=256, normalize=True, train_task='', test_datasets=None, results_folder='./logs/run',
                   render_dir='./renders/run', device='cuda', math='bf16x3', **kwargs) -> Trainer:
    """train_utils.py:185-313 reduced to the model/diffusion/trainer factory (no datasets, no wandb)."""
    dims = dims_for(input_mode, 'Triangular' in train_task)
    denoise_fn = ConstraintDiffuser(dims=dims, hidden_dim=hidden_dim, EBM=EBM, input_mode=input_mode, normalize=normalize,
                                    energy_wrapper=False, device=device, verbose=False, math=math)
    diffusion = GaussianDiffusion(denoise_fn, timesteps=timesteps, EBM=EBM, samples_per_step=samples_per_step,
                                  step_sizes=step_sizes).eval()
    return Trainer(diffusion, None, test_datasets, render_dir, results_folder=results_folder, EBM=EBM, input_mode=input_mode)


def load_trainer(run_id, milestone, logs_dir='./logs', **kwargs) -> Trainer:
    """train_utils.py:340-354: build the trainer for `run_id` and load `logs/<run_id>/model-<milestone>.pt`.
    The reference recovers the run's flags from wandb/<run>/files/config.yaml; here they are passed as kwargs
    (or read from logs/<run_id>/config.json when present)."""
    cfg_path = os.path.join(logs_dir, str(run_id), 'config.json')
    cfg = json.load(open(cfg_path)) if os.path.exists(cfg_path) else {}
    cfg.update(kwargs)
    trainer = create_trainer(results_folder=os.path.join(logs_dir, str(run_id)), **cfg)
    trainer.load(milestone)
    return trainer


def evaluate_model(run_id, milestone, tries=(10, 0), json_name='eval', save_log=True, run_all=False, render=True,
                   run_only=False, resume_eval=False, render_name_extra=None, return_history=False, **kwargs):
    """solve_csp.py:19-28."""
    trainer = load_trainer(run_id, milestone, **kwargs)
    if render_name_extra is not None:
        trainer.render_dir += f'_{render_name_extra}'
    return trainer.evaluate(json_name, tries=tries, render=render, save_log=save_log, run_all=run_all, run_only=run_only,
                            resume_eval=resume_eval, return_history=return_history)
