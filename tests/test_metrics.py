"""ADE/FDE evaluator (host glue): known answers on hand-made trajectories."""
import numpy as np

from desire_b200.utils.metrics import displacement_errors, evaluate


def _case():
    B, N, K, T = 1, 2, 4, 3
    true = np.zeros((B, N, T, 2))
    true[..., 0] = np.arange(1, T + 1)                       # moves +1 in x per frame
    pred = np.repeat(true[:, :, None], K, 2).copy()
    for k in range(K):
        pred[:, :, k, :, 1] += k                             # sample k is off by k in y at every frame
    scores = np.array([[[0.1, 0.9, 0.5, 0.2], [0.3, 0.2, 0.1, 0.9]]])      # agent 0 ranks k=1 first, agent 1 ranks k=3 first
    mask = np.array([[True, True]])
    return pred, true, scores, mask


def test_displacement_errors_known_answer():
    pred, true, _, _ = _case()
    ade, fde = displacement_errors(pred, true)
    assert np.allclose(ade[0, 0], [0, 1, 2, 3]) and np.allclose(fde[0, 1], [0, 1, 2, 3])


def test_evaluate_rankings_and_mask():
    pred, true, scores, mask = _case()
    r = evaluate(pred, true, scores, mask, top_frac=0.5)      # top-2
    assert r["top_k"] == 2 and r["n_agents"] == 2
    assert np.isclose(r["ade_best"], 0.0)
    assert np.isclose(r["ade_top1"], (1 + 3) / 2)             # k=1 for agent 0, k=3 for agent 1
    assert np.isclose(r["ade_topk"], (1 + 0) / 2)             # agent 0: best of {1,2} -> 1; agent 1: best of {3,0} -> 0
    assert np.isclose(r["fde_mean"], 1.5)
    r2 = evaluate(pred, true, scores, np.array([[True, False]]), top_frac=0.1)
    assert r2["n_agents"] == 1 and r2["top_k"] == 1 and np.isclose(r2["ade_topk"], 1.0)
