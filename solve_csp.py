"""solve_csp.py — same entry point as the reference's solve_csp.py:19-28, over the B200 sampling path.

    from solve_csp import evaluate_model
    evaluate_model(run_id, milestone, tries=(10, 0), input_mode='qualitative', test_datasets={8: [batch]})
"""
from diffusion_ccsp_b200.trainer import evaluate_model, load_trainer  # noqa: F401

if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('run_id'); ap.add_argument('milestone', type=int)
    ap.add_argument('-input_mode', default='qualitative'); ap.add_argument('-timesteps', type=int, default=1000)
    a = ap.parse_args()
    print(evaluate_model(a.run_id, a.milestone, input_mode=a.input_mode, timesteps=a.timesteps))
