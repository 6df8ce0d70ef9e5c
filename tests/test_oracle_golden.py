"""CPU: pin oracle/ccsp_oracle.py against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Tolerances are relative to max(1, max|x_ref|)."""
import numpy as np
import pytest

from oracle import ccsp_oracle as orc
from diffusion_ccsp_b200 import synthetic
from tests.util import case_model, golden_names, load_golden, rel_err

SCHED_TOL = 0.0          # same float64 formulas, same cast -> bit-exact
FWD_TOL = 2e-6           # one denoiser evaluation (BLAS summation order only)
TRAJ_TOL = 1e-5          # full trajectories, T <= 100


@pytest.mark.parametrize('T', [100, 1000])
def test_schedule_tables_bit_exact(T):
    z, _ = load_golden(f'schedule_T{T}')
    s = orc.make_schedule(T)
    for k, ref in z.items():
        assert s[k].dtype == np.float32
        assert np.array_equal(s[k], ref), k


def test_time_embedding():
    z, _ = load_golden('time_embedding')
    dims = synthetic.DIMS['qualitative']
    sd = synthetic.make_state_dict(dims, 'qualitative', seed=int(z['weight_seed']))
    den = orc.OracleDenoiser({k: v.numpy() for k, v in sd.items()}, dims, 'qualitative')
    pe = orc.sinusoidal_pos_emb(z['t'], 256)
    assert np.max(np.abs(pe - z['pos_emb'])) < 2e-4      # sin/cos of arguments up to 999 rad in f32
    te = den.time_mlp(z['t'])
    assert rel_err(te, z['time_mlp']) < 1e-4


@pytest.mark.parametrize('name', golden_names('forward_'))
def test_forward(name):
    z, batch = load_golden(name)
    mode, dims, sd = case_model(z)
    den = orc.OracleDenoiser({k: v.numpy() for k, v in sd.items()}, dims, mode)
    for t, ref in zip(z['t'], z['out']):
        out = den.forward(z['poses_in'], batch, int(t))
        assert out.shape == ref.shape
        assert rel_err(out, ref) < FWD_TOL, (name, t, rel_err(out, ref))
        m = batch.mask.numpy().astype(bool)
        assert np.array_equal(out[m], batch.x.numpy()[:, -dims[-1][0]:][m])


@pytest.mark.parametrize('name', golden_names('traj_'))
def test_trajectory(name):
    z, batch = load_golden(name)
    mode, dims, sd = case_model(z)
    T, K = int(z['T']), int(z['K'])
    EBM = 'ULA' if K > 0 else False
    den = orc.OracleDenoiser({k: v.numpy() for k, v in sd.items()}, dims, mode)
    gd = orc.OracleDiffusion(den, timesteps=T, EBM=EBM, samples_per_step=K)
    noise = synthetic.make_noise(T, K, batch.num_nodes, dims[-1][0], seed=int(z['noise_seed'])).numpy()
    assert noise.shape[0] == orc.num_noise_draws(T, K)
    out, hist = gd.p_sample_loop(batch, noise, return_history=True)
    assert len(hist) == T + 1
    assert rel_err(out, z['out']) < TRAJ_TOL, (name, rel_err(out, z['out']))
    assert rel_err(np.stack(hist), z['history']) < TRAJ_TOL
    m = batch.mask.numpy().astype(bool)
    gt = batch.x.numpy()[:, dims[-1][1]:dims[-1][2]]
    for h in hist:
        assert np.array_equal(h[m], gt[m])


def test_ula_plus_schedule():
    """EBM='ULA+': 4/8/12/16 ULA steps per quarter of the schedule (ddpm.py:297-299)."""
    z, batch = load_golden('ulaplus_qualitative_T8')
    mode, dims, sd = case_model(z)
    den = orc.OracleDenoiser({k: v.numpy() for k, v in sd.items()}, dims, mode)
    gd = orc.OracleDiffusion(den, timesteps=8, EBM='ULA+')
    noise = np.random.default_rng(int(z['noise_seed'])).standard_normal((1 + 8 + 2 * 40, batch.num_nodes, 4), dtype=np.float32)
    out, hist = gd.p_sample_loop(batch, noise, return_history=True)
    assert rel_err(out, z['out']) < TRAJ_TOL and rel_err(np.stack(hist), z['history']) < TRAJ_TOL


def test_trained_regime_T100():
    """Realistic O(1) regime: weights partially trained with the reference's own loss
    (tests/golden/make_trained_fixture.py); oracle vs the reference's trajectory."""
    z, batch = load_golden('trained_traj_qualitative_T100')
    dims = synthetic.DIMS['qualitative']
    sd = synthetic.make_trained_state_dict()
    den = orc.OracleDenoiser({k: v.numpy() for k, v in sd.items()}, dims, 'qualitative')
    gd = orc.OracleDiffusion(den, timesteps=100, EBM='ULA', samples_per_step=10)
    noise = synthetic.make_noise(100, 10, batch.num_nodes, 4, seed=int(z['noise_seed'])).numpy()
    out, hist = gd.p_sample_loop(batch, noise, return_history=True)
    assert np.abs(z['out']).max() < 2.0                      # the regime really is O(1)
    assert rel_err(out, z['out']) < TRAJ_TOL, rel_err(out, z['out'])
    assert rel_err(np.stack(hist)[::int(z['history_every'])], z['history']) < TRAJ_TOL
