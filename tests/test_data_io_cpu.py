"""CPU: on-disk formats and the batch builder.

  * data_transform_cn_diffuse_batch == the reference's own transform (networks/data_transforms.py:26-200, imported unmodified
    through oracle/ref_shim.py), bit for bit, for boxes / qualitative / triangle (theta and sin-cos) / robot rows;
  * raw `data_i.pt` files pickled as torch_geometric Data objects (PyG 1.x and 2.x layouts) are read without PyG;
  * wandb config.yaml -> flags; GraphDataset; Trainer.load / load_trainer / evaluate_model / solve_csp with a stub sampler.
"""
import json
import os
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from diffusion_ccsp_b200 import data_io, scenes, synthetic
from oracle import ref_shim

needs_ref = pytest.mark.skipif(not ref_shim.reference_available(), reason='reference sources not available')


def raw_qualitative(batch_scene, W=3.0, L=2.0):
    """invert the normalisation of a fixture scene: raw rows [typ, w, l, x, y, sn, cs] and named edges"""
    x = batch_scene.x.numpy().astype(np.float64)
    rows = [[0.0, W, L, 0.0, 0.0, 0.0, 0.0]]
    for r in x[1:]:
        rows.append([1.0, r[0] * W, r[1] * L, r[2] * W / 2, r[3] * L / 2, r[5], r[4]])       # stored pose is [x, y, cs, sn]
    names = scenes.qualitative_constraints
    edges = [(names[int(t)], int(a), int(b)) for t, a, b in zip(batch_scene.edge_attr, *batch_scene.edge_index)]
    return data_io.RawGraph(np.array(rows, np.float32), edges, torch.zeros(len(rows)))


def random_raw(kind, rng):
    n = int(rng.integers(2, 7))
    if kind == 'boxes':
        rows = [[0, 3.0, 2.0, 0, 0]] + [[1, *rng.uniform(0.2, 1.0, 2), *rng.uniform(-1, 1, 2)] for _ in range(n)]
        mode, names = 'diffuse_pairwise', scenes.puzzle_constraints
    elif kind == 'triangles_theta':
        rows = [[0, 3.0, 3.0, 0, 0, 0, 0]] + [[1, *rng.uniform(0.2, 1.0, 3), *rng.uniform(-1, 1, 2), rng.uniform(-3, 3)] for _ in range(n)]
        mode, names = 'diffuse_pairwise', scenes.puzzle_constraints
    elif kind == 'triangles_sincos':
        rows = [[0, 3.0, 3.0, 0, 0, 0, 0, 0]] + [[1, *rng.uniform(0.2, 1.0, 3), *rng.uniform(-1, 1, 4)] for _ in range(n)]
        mode, names = 'diffuse_pairwise', scenes.puzzle_constraints
    elif kind == 'stability':
        rows = [[0, 3.0, 2.0, 0, 0, 0, 0]] + [[1, *rng.uniform(0.2, 1.0, 2), *rng.uniform(-1, 1, 4)] for _ in range(n)]
        mode, names = 'stability_flat', scenes.stability_constraints
    else:
        rows = [[0] + list(rng.uniform(0.1, 1, 21))] + [[1] + list(rng.uniform(-1, 1, 21)) for _ in range(n)]
        mode, names = 'robot_box', scenes.robot_constraints
    edges = [(names[0], i, 0) for i in range(1, n + 1)] + [(names[-1], i, j) for i in range(1, n + 1) for j in range(i + 1, n + 1)]
    return data_io.RawGraph(np.array(rows, np.float32), edges, torch.zeros(n + 1)), mode


def assert_same_as_reference(raw, idx, mode):
    ref_shim.load_reference()
    import data_transforms as ref_dt
    torch.manual_seed(0)
    ref = ref_dt.data_transform_cn_diffuse_batch(raw, idx, mode)[0]
    got = data_io.data_transform_cn_diffuse_batch(raw, idx, mode)
    assert torch.equal(got.x, ref.x) and got.x.dtype == ref.x.dtype
    assert torch.equal(got.edge_index, ref.edge_index) and torch.equal(got.edge_attr, ref.edge_attr)
    assert torch.equal(got.mask, ref.mask) and got.mask.dtype == ref.mask.dtype
    assert torch.equal(got.x_extract, ref.x_extract) and torch.equal(got.edge_extract, ref.edge_extract)
    assert tuple(got.world_dims[0]) == tuple(ref.world_dims)
    return got


@needs_ref
def test_transform_matches_reference_on_qualitative_fixtures():
    pool = scenes.qualitative_batch(32, 8)
    for i in range(0, 32, 5):
        sc = pool.select_scenes(i, i + 1)
        got = assert_same_as_reference(raw_qualitative(sc), i, 'qualitative')
        # and the round trip lands back on the fixture rows (up to the float32 division)
        assert np.allclose(got.x.numpy(), sc.x.numpy(), atol=1e-6)


@needs_ref
@pytest.mark.parametrize('kind', ['boxes', 'triangles_theta', 'triangles_sincos', 'stability', 'robot'])
def test_transform_matches_reference_other_row_layouts(kind):
    rng = np.random.default_rng(3)
    for idx in range(5):
        raw, mode = random_raw(kind, rng)
        assert_same_as_reference(raw, idx, mode)


def _fake_pyg_pickle(path, layout, x, edge_index, y):
    """write a pickle that names torch_geometric classes, with PyG absent afterwards"""
    mods = {}
    for name in ('torch_geometric', 'torch_geometric.data', 'torch_geometric.data.data', 'torch_geometric.data.storage'):
        mods[name] = types.ModuleType(name)

    class Data:
        pass

    class GlobalStorage:
        pass

    Data.__module__, Data.__qualname__ = 'torch_geometric.data.data', 'Data'
    GlobalStorage.__module__, GlobalStorage.__qualname__ = 'torch_geometric.data.storage', 'GlobalStorage'
    mods['torch_geometric.data.data'].Data = Data
    mods['torch_geometric.data.storage'].GlobalStorage = GlobalStorage
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        d = Data()
        if layout == 'v1':
            d.__dict__.update(x=x, edge_index=edge_index, y=y)
        else:
            st = GlobalStorage()
            st.__dict__['_mapping'] = dict(x=x, edge_index=edge_index, y=y)
            d.__dict__['_store'] = st
        torch.save(d, path)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.parametrize('layout', ['v1', 'v2'])
def test_reference_written_raw_files_are_read_without_pyg(tmp_path, layout):
    assert 'torch_geometric' not in sys.modules or not hasattr(sys.modules['torch_geometric'], '__path__')
    x = torch.tensor([[0, 3, 2, 0, 0, 0, 0], [1, 0.5, 0.4, 0.1, -0.2, 0.0, 1.0]], dtype=torch.float)
    ei = [('in', 1, 0), ('left-in', 1, 0)]
    p = str(tmp_path / 'data_0.pt')
    _fake_pyg_pickle(p, layout, x, ei, torch.zeros(2))
    raw = data_io.load_raw_graph(p)
    assert torch.equal(raw.x, x) and raw.edge_index == ei
    sc = data_io.data_transform_cn_diffuse_batch(raw, 0, 'qualitative')
    assert sc.x.shape == (2, 6) and sc.edge_attr.tolist() == [0.0, 2.0]


def test_graph_dataset_and_loader(tmp_path):
    pool = scenes.qualitative_batch(6, 4)
    name = 'RandomSplitQualitativeWorld(6)_qualitative_test_4_split'
    os.makedirs(tmp_path / name / 'raw')
    for i in range(6):
        raw = raw_qualitative(pool.select_scenes(i, i + 1))
        data_io.save_raw_graph(str(tmp_path / name / 'raw' / f'data_{i}.pt'), raw.x.numpy(), raw.edge_index, None)
    ds = data_io.GraphDataset(name, 'qualitative', root=str(tmp_path))
    assert len(ds) == 6 and ds.length == 6
    batch = scenes.collate(list(ds))
    assert batch.num_graphs == 6 and np.allclose(batch.x.numpy(), pool.x.numpy(), atol=1e-6)
    assert torch.equal(batch.edge_index, pool.edge_index) and torch.equal(batch.edge_attr, pool.edge_attr)
    assert batch.x_extract.tolist() == pool.x_extract.tolist()


def test_wandb_config_is_parsed_like_the_reference(tmp_path):
    run = tmp_path / 'wandb' / 'run-20230101_000000-qsd3ju74' / 'files'
    os.makedirs(run)
    (run / 'config.yaml').write_text(
        'wandb_version: 1\n_wandb:\n  desc: null\n  value:\n    cli_version: 0.13\n'
        'EBM:\n  desc: null\n  value: ULA\ntimesteps:\n  desc: null\n  value: 1000\ninput_mode:\n  desc: null\n  value: qualitative\n'
        'samples_per_step:\n  desc: null\n  value: 10\ntrain_lr:\n  desc: null\n  value: 0.123\nstep_sizes:\n  desc: null\n  value: 2*self.betas\n'
        'train_task:\n  desc: null\n  value: RandomSplitQualitativeWorld(30000)_qualitative_train\n')
    args = data_io.get_args_from_run_id('qsd3ju74', wandb_roots=(str(tmp_path / 'wandb'),))
    assert args.EBM == 'ULA' and args.timesteps == 1000 and args.input_mode == 'qualitative' and args.samples_per_step == 10
    assert not hasattr(args, 'train_lr')                    # train_batch_size / train_lr are not taken from the run (train_utils.py:325-326)
    assert args.normalize is True and args.hidden_dim == 256 and args.wandb_version == 1 if hasattr(args, 'wandb_version') else True
    # the reference's per-run patches (train_utils.py:329-336)
    run2 = tmp_path / 'wandb' / 'run-x-4xt8u4n7' / 'files'
    os.makedirs(run2)
    (run2 / 'config.yaml').write_text('EBM:\n  value: ULA\ninput_mode:\n  value: robot_box\n')
    assert data_io.get_args_from_run_id('4xt8u4n7', wandb_roots=(str(tmp_path / 'wandb'),)).normalize is False
    with pytest.raises(FileNotFoundError):
        data_io.get_args_from_run_id('nope', wandb_roots=(str(tmp_path / 'wandb'),))


# ---------------------------------------------------------------------------------------------------------------------
# Trainer / load_trainer / evaluate_model / solve_csp around a stub sampler (no GPU)
# ---------------------------------------------------------------------------------------------------------------------
def _stub_sampling(monkeypatch, trainer_mod):
    """GaussianDiffusion.sample -> ground-truth poses for every even try, zeros otherwise; checker -> CPU callback"""
    from diffusion_ccsp_b200.ddpm import GaussianDiffusion
    calls = dict(n=0)

    def fake_sample(self, batch, **kw):
        calls['n'] += 1
        self.sample_loop_time.append(0.01)
        p0, p1 = self.dims[-1][1], self.dims[-1][2]
        out = batch.x[:, p0:p1].clone() if calls['n'] % 2 == 0 else torch.zeros(batch.x.shape[0], p1 - p0)
        return (out, [out]) if kw.get('return_history') else out

    monkeypatch.setattr(GaussianDiffusion, 'sample', fake_sample)
    return calls


def _oracle_checker(rows, batch):
    from oracle import checker_oracle as chk
    s, _, _ = chk.check_batch(rows[:, 2:6].numpy(), batch, (2, 6))
    return s.tolist()


def test_trainer_evaluate_save_load_and_evaluate_model(tmp_path, monkeypatch):
    import importlib.util
    spec = importlib.util.spec_from_file_location('ccsp_b200_solve_csp', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'solve_csp.py'))
    solve_csp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(solve_csp)          # the repo's entry script (the reference's own solve_csp.py may be on sys.path)
    from diffusion_ccsp_b200 import trainer as tr
    calls = _stub_sampling(monkeypatch, tr)
    test_sets = {4: [scenes.qualitative_batch(8, 4)], 3: [scenes.qualitative_batch(4, 3)]}
    t = tr.create_trainer('qualitative', timesteps=10, test_datasets=test_sets, results_folder=str(tmp_path / 'logs' / 'abc123'),
                          render_dir=str(tmp_path / 'renders'), device='cpu')
    # checkpoints in the reference's layout (ddpm.py:496-514)
    t.step = 7
    t.save(3)
    ck = data_io.load_checkpoint(str(tmp_path / 'logs' / 'abc123' / 'model-3.pt'))
    assert ck['step'] == 7 and 'denoise_fn.mlps.12.0.weight' in ck['model'] and 'betas' in ck['model']
    log = t.evaluate('unit', tries=(4, 0), checker=_oracle_checker, save_log=True)
    # try 0 returns zeros (everything collides), try 1 the ground truth: solved in the second round
    assert log[4]['success_rate'] == 0.0 and log[4]['success_rate_top3'] == 1.0 and log[4]['scenes'] == 8
    assert log[3]['success_rate_top3'] == 1.0 and set(log[4]['success_rounds'].values()) == {1}
    assert os.path.isfile(tmp_path / 'renders' / 'denoised_unit.json')
    assert log[4]['model_ave_sample_time'] > 0
    # load_trainer without a wandb directory: flags from keyword arguments; weights from logs/<run>/model-<k>.pt
    t2 = tr.load_trainer('abc123', 3, logs_dir=str(tmp_path / 'logs'), wandb_roots=(str(tmp_path / 'none'),), input_mode='qualitative',
                         timesteps=10, device='cpu', test_datasets=test_sets, render_dir=str(tmp_path / 'r2'))
    assert t2.step == 7
    for (k, a), (_, b) in zip(t.model.state_dict().items(), t2.model.state_dict().items()):
        assert torch.equal(a, b), k
    # solve_csp.evaluate_model (solve_csp.py:19-28) drives the same objects
    monkeypatch.setattr(solve_csp, 'load_trainer', lambda run_id, milestone, **kw: t2, raising=False)
    monkeypatch.setattr(tr, 'load_trainer', lambda run_id, milestone, **kw: t2)
    out = solve_csp.evaluate_model('abc123', 3, tries=(2, 0), json_name='em', render_name_extra='x', checker=_oracle_checker)
    assert solve_csp.evaluate_model is tr.evaluate_model
    assert set(out) == {4, 3} and t2.render_dir.endswith('_x')
    assert calls['n'] > 0


def test_load_trainer_reads_wandb_flags_and_test_datasets(tmp_path, monkeypatch):
    from diffusion_ccsp_b200 import trainer as tr
    run = tmp_path / 'wandb' / 'run-1-zzz999' / 'files'
    os.makedirs(run)
    (run / 'config.yaml').write_text('EBM:\n  value: ULA+\ntimesteps:\n  value: 8\ninput_mode:\n  value: qualitative\n'
                                     'samples_per_step:\n  value: 3\n')
    name = 'RandomSplitQualitativeWorld(3)_qualitative_test_4_split'
    os.makedirs(tmp_path / 'data' / name / 'raw')
    pool = scenes.qualitative_batch(3, 4)
    for i in range(3):
        raw = raw_qualitative(pool.select_scenes(i, i + 1))
        data_io.save_raw_graph(str(tmp_path / 'data' / name / 'raw' / f'data_{i}.pt'), raw.x.numpy(), raw.edge_index)
    src = tr.create_trainer('qualitative', timesteps=8, EBM='ULA+', samples_per_step=3, results_folder=str(tmp_path / 'logs' / 'zzz999'),
                            device='cpu')
    src.save(0)
    t = tr.load_trainer('zzz999', 0, logs_dir=str(tmp_path / 'logs'), wandb_roots=(str(tmp_path / 'wandb'),), data_root=str(tmp_path / 'data'),
                        test_tasks={4: 'RandomSplitQualitativeWorld(3)_test_4_split'}, device='cpu')
    assert t.model.EBM == 'ULA+' and t.model.num_timesteps == 8 and t.model.samples_per_step == 3
    assert list(t.test_datasets) == [4] and t.test_datasets[4][0].num_graphs == 3
