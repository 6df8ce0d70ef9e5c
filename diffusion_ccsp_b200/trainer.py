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
from .scenes import SceneBatch, SceneLoader
from .synthetic import dims_for


class Trainer(object):
    def __init__(self, denoise_fn: GaussianDiffusion, train_dataset=None, test_datasets: Optional[Dict] = None,
                 render_dir: str = './renders', *, train_batch_size=32, train_lr=2e-3, train_num_steps=100000,
                 gradient_accumulate_every=2, save_and_sample_every=10000, results_folder: str = './results', EBM=False,
                 eval_only=True, input_mode=None, loader_seed=0, **kwargs):
        """ddpm.py:395-490.  `train_dataset`: a SceneBatch pool (or list of single-scene SceneBatch objects) standing in for the
        GraphDataset; `test_datasets`: {n_objects: iterable of SceneBatch}.  EMA is disabled in the reference (ema_model = None,
        ddpm.py:426) and therefore absent here."""
        self.model = denoise_fn
        self.dims = denoise_fn.dims
        self.input_mode = denoise_fn.input_mode
        self.EBM = EBM
        self.batch_size = train_batch_size
        self.gradient_accumulate_every = gradient_accumulate_every
        self.train_num_steps = train_num_steps
        self.save_and_sample_every = save_and_sample_every
        self.train_dl = SceneLoader(train_dataset, train_batch_size, shuffle=True, seed=loader_seed) if train_dataset is not None else None
        self.test_datasets = test_datasets or {}      # {n_objects: iterable of SceneBatch}
        self.eval_kwargs = dict(tries=(10, 0))
        self.render_dir = render_dir
        self.results_folder = Path(results_folder)
        self.train_lr = train_lr
        self.opt = None                                # train.Adam, created on first use (needs the parameters on the GPU)
        self.step = 0
        self.loss_log = []
        self.lr_schedule = None                        # optional callable step -> learning rate (the reference trains at a constant rate)

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

    def _optimizer(self):
        if self.opt is None:
            from .train import Adam
            den = self.model.denoise_fn
            den.to(den.cuda_device())
            self.opt = Adam(self.model.parameters(), lr=self.train_lr, on_step=den.mark_weights_dirty)   # ddpm.py:466
        return self.opt

    def train(self, log_every: int = 1000, evaluate: bool = True):
        """ddpm.py:519-556: gradient accumulation, Adam step, periodic save + evaluate.  The loss and every gradient come
        from the CUDA training step (`self.model(data, debug=False, tag='EBM')` -> ccsp_train_step)."""
        if self.train_dl is None:
            raise ValueError('Trainer.train needs a train_dataset')
        opt = self._optimizer()
        self.model.train()
        dl_iter = iter(self.train_dl)
        window = []
        while self.step < self.train_num_steps:
            for _ in range(self.gradient_accumulate_every):
                try:
                    data = next(dl_iter)
                except StopIteration:
                    dl_iter = iter(self.train_dl)
                    data = next(dl_iter)
                loss = self.model(data, debug=False, tag='EBM')
                window.append(loss.detach())
                (loss / self.gradient_accumulate_every).backward()                  # ddpm.py:534
            if (self.step + 1) % log_every == 0:
                mean = float(torch.stack(window).mean())
                self.loss_log.append((self.step, mean))
                print(f"Step: {self.step} | lr: {opt.param_groups[0]['lr']}\tloss={mean:.6f}")
                window = []
            if self.lr_schedule is not None:
                opt.param_groups[0]['lr'] = float(self.lr_schedule(self.step))
            opt.step()
            opt.zero_grad()
            if self.step % self.save_and_sample_every == (self.save_and_sample_every - 1):
                milestone = self.step // self.save_and_sample_every
                self.save(milestone)
                if evaluate and self.test_datasets:
                    self.model.eval()
                    self.evaluate(milestone, **self.eval_kwargs)
                    self.model.train()
            self.step += 1
        self.model.eval()
        print('training completed')

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
                   hidden_dim=256, normalize=True, train_task='', train_dataset=None, test_datasets=None, train_num_steps=300000,
                   train_batch_size=128, train_lr=5e-4, results_folder='./logs/run', render_dir='./renders/run', device='cuda',
                   math='bf16x3', loss_type='l2', **kwargs) -> Trainer:
    """train_utils.py:185-313 reduced to the model / diffusion / trainer factory (datasets are handed in; no wandb).
    Training configuration as at train_utils.py:216-219, 299-313: batch 128, lr 5e-4, gradient_accumulate_every=1."""
    dims = dims_for(input_mode, 'Triangular' in train_task)
    denoise_fn = ConstraintDiffuser(dims=dims, hidden_dim=hidden_dim, EBM=EBM, input_mode=input_mode, normalize=normalize,
                                    energy_wrapper=False, device=device, verbose=False, math=math)
    diffusion = GaussianDiffusion(denoise_fn, timesteps=timesteps, loss_type=loss_type, EBM=EBM, samples_per_step=samples_per_step,
                                  step_sizes=step_sizes).eval()
    return Trainer(diffusion, train_dataset, test_datasets, render_dir, train_batch_size=train_batch_size, train_lr=train_lr,
                   train_num_steps=train_num_steps, gradient_accumulate_every=1,
                   save_and_sample_every=kwargs.pop('save_and_sample_every', 10000),
                   results_folder=results_folder, EBM=EBM, input_mode=input_mode, **kwargs)


def load_trainer(run_id, milestone, logs_dir='./logs', wandb_roots=('wandb', 'wandb2'), test_datasets=None, data_root='data',
                 **kwargs) -> Trainer:
    """train_utils.py:340-354: recover the run's flags from wandb/<run>/files/config.yaml (`data_io.get_args_from_run_id`),
    build the trainer and load logs/<run_id>/model-<milestone>.pt (ddpm.py:503-514).  Keyword arguments override the
    recovered flags (the reference allows input_mode / train_task / test_tasks / train_num_steps, train_utils.py:344-347);
    when there is no wandb directory for the run the flags come from the keyword arguments alone.  Test datasets named by
    `test_tasks` are read from `data_root` when present (`data_io.GraphDataset`)."""
    from . import data_io
    try:
        args = vars(data_io.get_args_from_run_id(str(run_id), wandb_roots))
    except FileNotFoundError:
        args = dict(data_io.ARG_DEFAULTS, run_id=str(run_id), input_mode=kwargs.get('input_mode', 'qualitative'),
                    EBM=kwargs.get('EBM', 'ULA'))
    args.update(kwargs)
    if args.get('model', 'Diffusion-CCSP') != 'Diffusion-CCSP' or args.get('energy_wrapper'):
        raise NotImplementedError('StructDiffusion / energy-wrapper runs are outside the accelerated path (SURVEY.md §2 #3, #6)')
    input_mode = args['input_mode']
    if test_datasets is None and args.get('test_tasks'):
        test_datasets = {}
        for k, task in args['test_tasks'].items():
            if input_mode not in task:
                task = task.replace('_test', f'_{input_mode}_test')                 # train_utils.py:243-250
            if os.path.isdir(os.path.join(data_root, task, 'raw')):
                ds = data_io.GraphDataset(task, input_mode, root=data_root)
                from .scenes import SceneLoader
                test_datasets[k] = list(SceneLoader(ds.scenes, 100, shuffle=False))      # ddpm.py:446-449 (batch size 100)
    keep = ('timesteps', 'EBM', 'samples_per_step', 'step_sizes', 'hidden_dim', 'normalize', 'train_task', 'train_num_steps')
    cfg = {k: args[k] for k in keep if k in args and args[k] is not None}
    for k in ('device', 'math', 'render_dir'):
        if k in kwargs:
            cfg[k] = kwargs[k]
    trainer = create_trainer(input_mode=input_mode, test_datasets=test_datasets,
                             results_folder=os.path.join(logs_dir, str(run_id)), **cfg)
    trainer.load(milestone)
    return trainer


def evaluate_model(run_id, milestone, tries=(10, 0), json_name='eval', save_log=True, run_all=False, render=True,
                   run_only=False, resume_eval=False, render_name_extra=None, return_history=False, **kwargs):
    """solve_csp.py:19-28.  Extra keyword `checker=` (a `(rows, batch) -> list[bool]` callback) replaces the GPU solved-checker."""
    checker = kwargs.pop('checker', None)
    trainer = load_trainer(run_id, milestone, **kwargs)
    if render_name_extra is not None:
        trainer.render_dir += f'_{render_name_extra}'
    return trainer.evaluate(json_name, tries=tries, render=render, save_log=save_log, run_all=run_all, run_only=run_only,
                            resume_eval=resume_eval, return_history=return_history, checker=checker)
