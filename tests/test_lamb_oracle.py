"""CPU checks of the LAMB oracle: C restatement vs independent numpy restatement vs closed-form
known answers derived by hand from multi_tensor_lamb.cu:121-157,274-282."""
import copy

import numpy as np
import pytest

from oracle import lamb_oracle as LO


def make_groups(rng, shapes, wd=(0.01, 0.0), lr=1e-3, gscale=1e-3):
    groups = []
    for gi, shp_list in enumerate(shapes):
        ps = [rng.standard_normal(s).astype(np.float32) * 0.02 for s in shp_list]
        groups.append(dict(params=ps,
                           grads=[rng.standard_normal(p.shape).astype(np.float32) * gscale for p in ps],
                           exp_avg=[np.zeros_like(p) for p in ps], exp_avg_sq=[np.zeros_like(p) for p in ps],
                           lr=lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=wd[gi], step=0,
                           bias_correction=True, grad_averaging=True))
    return groups


def test_c_matches_numpy_multi_step():
    rng = np.random.default_rng(0)
    a = make_groups(rng, [[(33, 17), (128,), (5, 5, 3)], [(7,), (2,)]])
    b = copy.deepcopy(a)
    for it in range(4):
        for ga, gb in zip(a, b):
            for i in range(len(ga["grads"])):
                g = rng.standard_normal(ga["grads"][i].shape).astype(np.float32) * (10.0 if it == 2 else 1e-3)
                ga["grads"][i] = g.copy(); gb["grads"][i] = g.copy()
        ra = LO.lamb_step(a, scale=1024.0 if it % 2 else 1.0)
        rb = LO.lamb_step_numpy(b, scale=1024.0 if it % 2 else 1.0)
        assert ra["found_inf"] == rb["found_inf"] is False
        assert ra["global_grad_norm"] == pytest.approx(rb["global_grad_norm"], rel=1e-6)
    for ga, gb in zip(a, b):
        assert ga["step"] == gb["step"] == 4
        for k in ("params", "exp_avg", "exp_avg_sq"):
            for x, y in zip(ga[k], gb[k]):
                np.testing.assert_allclose(x, y, rtol=2e-6, atol=1e-9)


def test_known_answer_first_step():
    """step 1, m=v=0, no clipping: m = (1-b1) g, v = (1-b2) g^2, bias-corrected => u = g/(|g|+eps) + wd*p,
    p_new = p - lr*|p|/|u| * u."""
    p = np.array([3.0, -4.0], np.float32)            # |p| = 5
    g = np.array([0.1, -0.2], np.float32)            # |g| < max_grad_norm => no clip
    grp = dict(params=[p.copy()], grads=[g.copy()], exp_avg=[np.zeros(2, np.float32)],
               exp_avg_sq=[np.zeros(2, np.float32)], lr=0.5, betas=(0.9, 0.999), eps=1e-6,
               weight_decay=0.01, step=0, bias_correction=True, grad_averaging=True)
    r = LO.lamb_step([grp])
    u = g / (np.abs(g) + 1e-6) + 0.01 * p            # ~ [1.03, -1.04]
    want = p - 0.5 * (5.0 / np.linalg.norm(u)) * u
    np.testing.assert_allclose(grp["params"][0], want, rtol=1e-5)
    np.testing.assert_allclose(grp["exp_avg"][0], 0.1 * g, rtol=1e-6)
    np.testing.assert_allclose(grp["exp_avg_sq"][0], 0.001 * g * g, rtol=1e-4)
    assert grp["step"] == 1 and r["param_norms"][0] == pytest.approx(5.0)


def test_clip_and_unscale():
    """grads scaled by 2^10 with true norm 4 > max_grad_norm 1: effective grad = g_true/4."""
    rng = np.random.default_rng(1)
    gt = rng.standard_normal(64).astype(np.float32)
    gt *= 4.0 / np.linalg.norm(gt)
    mk = lambda grads: dict(params=[np.ones(64, np.float32)], grads=[grads], exp_avg=[np.zeros(64, np.float32)],
                            exp_avg_sq=[np.zeros(64, np.float32)], lr=1e-2, betas=(0.9, 0.999), eps=1e-6,
                            weight_decay=0.0, step=0, bias_correction=True, grad_averaging=True)
    a, b = mk(gt * 1024.0), mk(gt / 4.0)
    LO.lamb_step([a], scale=1024.0)
    LO.lamb_step([b], scale=1.0)
    np.testing.assert_allclose(a["exp_avg"][0], b["exp_avg"][0], rtol=1e-6)
    np.testing.assert_allclose(a["params"][0], b["params"][0], rtol=1e-6)


def test_no_decay_group_uses_plain_lr():
    """decay == 0 and not nvlamb => ratio = lr (multi_tensor_lamb.cu:274-282)."""
    p = np.full(8, 2.0, np.float32)
    g = np.full(8, 1e-3, np.float32)
    grp = dict(params=[p.copy()], grads=[g], exp_avg=[np.zeros(8, np.float32)], exp_avg_sq=[np.zeros(8, np.float32)],
               lr=0.1, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, step=0, bias_correction=True, grad_averaging=True)
    LO.lamb_step([grp])
    np.testing.assert_allclose(grp["params"][0], p - 0.1 * (1e-3 / (1e-3 + 1e-6)), rtol=1e-5)


def test_overflow_skips_everything():
    rng = np.random.default_rng(2)
    a = make_groups(rng, [[(16,)], [(4,)]])
    a[0]["grads"][0][3] = np.inf
    before = copy.deepcopy(a)
    r = LO.lamb_step(a)
    assert r["found_inf"]
    for ga, gb in zip(a, before):
        assert ga["step"] == gb["step"] == 0
        for k in ("params", "exp_avg", "exp_avg_sq"):
            for x, y in zip(ga[k], gb[k]):
                np.testing.assert_array_equal(x, y)


def test_poly_warmup_schedule():
    # schedulers.py:131-136 with warmup 0.1, total 100, base 1.0
    assert LO.poly_warmup_lr(0, 100, 0.1, 1.0) == pytest.approx(0.1, rel=1e-6)
    assert LO.poly_warmup_lr(8, 100, 0.1, 1.0) == pytest.approx(0.9, rel=1e-6)
    assert LO.poly_warmup_lr(49, 100, 0.1, 1.0) == pytest.approx((1 - 0.5) ** 0.5, rel=1e-6)


def test_adam_oracle_matches_torch_adamw():
    """Pins oracle.adam_step_numpy (the SQuAD FusedAdam/clip restatement) to torch.optim.AdamW (decoupled decay, bias correction)."""
    import torch
    rng = np.random.default_rng(3)
    shapes = [(17, 5), (64,), (3, 3, 3)]
    ps = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    tps = [torch.nn.Parameter(torch.from_numpy(p.copy())) for p in ps]
    opt = torch.optim.AdamW(tps, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    grp = dict(params=[p.copy() for p in ps], grads=None, exp_avg=[np.zeros_like(p) for p in ps], exp_avg_sq=[np.zeros_like(p) for p in ps],
               lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, step=0, bias_correction=True)
    for it in range(4):
        gs = [rng.standard_normal(s).astype(np.float32) * 0.1 for s in shapes]
        for tp, g in zip(tps, gs):
            tp.grad = torch.from_numpy(g.copy())
        opt.step()
        grp["grads"] = gs
        LO.adam_step_numpy([grp], max_grad_norm=0.0)
    for tp, p in zip(tps, grp["params"]):
        np.testing.assert_allclose(tp.detach().numpy(), p, rtol=2e-5, atol=1e-7)


def test_adam_oracle_clip_matches_squad_clipper():
    """GradientClipper: grads scaled by max/(norm+1e-6) when that is < 1 (run_squad.py:721-724)."""
    g = np.full(16, 2.0, np.float32)                      # norm 8 > max 1
    mk = lambda grads: dict(params=[np.zeros(16, np.float32)], grads=[grads], exp_avg=[np.zeros(16, np.float32)],
                            exp_avg_sq=[np.zeros(16, np.float32)], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, step=0,
                            bias_correction=False)
    a, b = mk(g.copy()), mk(g * np.float32(1.0 / (8.0 + 1e-6)))
    LO.adam_step_numpy([a], max_grad_norm=1.0)
    LO.adam_step_numpy([b], max_grad_norm=0.0)
    np.testing.assert_allclose(a["exp_avg"][0], b["exp_avg"][0], rtol=1e-6)
    np.testing.assert_allclose(a["params"][0], b["params"][0], rtol=1e-5)
