"""Optimizer step after the path: oracle/optim_oracle.py against the reference's own AdamW +
clip_grad_norm_ run (tests/golden/adamw.npz, oracle/make_golden.py:golden_adamw), the host side of
vlbert_b200.optim.FusedAdamW, and -- on the GPU -- the two kernels behind it against the same fixture."""
import os

import numpy as np
import pytest
import torch

import optim_oracle as oo
from synth import adamw_case


def _oracle_run(G=None):
    params, grads, groups, lr_scale = adamw_case()
    p = [t.numpy().copy() for t in params]
    m = [np.zeros_like(a) for a in p]
    v = [np.zeros_like(a) for a in p]
    norms, snaps = [], []
    for k in range(4):
        g = [t.numpy() for t in grads[k]]
        coef, total = oo.clip_coef(g, 1.0)
        norms.append(total)
        for grp in groups:
            for i in grp["idx"]:
                oo.adamw_step(p[i], (g[i] * np.float32(coef)).astype(np.float32), m[i], v[i], k + 1, grp["lr"] * lr_scale[k],
                              weight_decay=grp["weight_decay"])
        snaps.append([a.copy() for a in p])
    return norms, snaps, m, v


def test_oracle_matches_reference_adamw_fixture(golden_dir):
    """fp32 both sides, same operation order: 2e-6 of the tensor's max (torch fuses some multiply-adds)."""
    G = np.load(os.path.join(golden_dir, "adamw.npz"))
    norms, snaps, m, v = _oracle_run()
    for k in range(4):
        assert abs(norms[k] / float(G["norm%d" % k]) - 1) < 1e-6
        for i in range(3):
            ref = G["p%d_step%d" % (i, k)]
            assert np.abs(snaps[k][i] - ref).max() <= 2e-6 * np.abs(ref).max(), (k, i)
    for i in range(3):
        assert np.abs(m[i] - G["m%d" % i]).max() <= 2e-6 * np.abs(G["m%d" % i]).max()
        assert np.abs(v[i] - G["v%d" % i]).max() <= 2e-6 * np.abs(G["v%d" % i]).max()


def test_fused_adamw_host_side():
    """constructor validation, param-group / state layout of the reference class; no CPU path"""
    import vlbert_b200
    from vlbert_b200.optim import FusedAdamW
    w = torch.nn.Parameter(torch.randn(4, 3))
    b = torch.nn.Parameter(torch.randn(3))
    opt = FusedAdamW([dict(params=[w], weight_decay=0.01), dict(params=[b], weight_decay=0.0, lr=5e-4)], lr=1e-3, max_grad_norm=1.0)
    assert [g["lr"] for g in opt.param_groups] == [1e-3, 5e-4]
    assert set(opt.param_groups[0].keys()) >= {"lr", "betas", "eps", "weight_decay", "correct_bias"}
    assert opt.defaults["eps"] == 1e-6 and opt.defaults["betas"] == (0.9, 0.999)          # optimization.py:116
    with pytest.raises(ValueError):
        FusedAdamW([w], lr=-1.0)
    with pytest.raises(ValueError):
        FusedAdamW([w], betas=(1.0, 0.999))
    opt.step()                                   # no gradients yet: nothing to do, no error
    w.grad = torch.zeros_like(w)
    with pytest.raises(RuntimeError):            # CPU tensors: loud failure, no fallback
        opt.step()


@pytest.mark.gpu
def test_fused_adamw_against_reference_fixture(golden_dir):
    """kernels vlb_grad_sqnorm + vlb_adamw_step through FusedAdamW: parameters after each of 4 steps, moments and the clip norm
    against the reference's run.  fp32 with the reference's operation order (no FMA contraction): 4e-6 of max."""
    from vlbert_b200.optim import FusedAdamW
    G = np.load(os.path.join(golden_dir, "adamw.npz"))
    params, grads, groups, lr_scale = adamw_case()
    ps = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    opt = FusedAdamW([dict(params=[ps[i] for i in g["idx"]], lr=g["lr"], weight_decay=g["weight_decay"]) for g in groups],
                     lr=1e-3, betas=(0.9, 0.999), eps=1e-6, max_grad_norm=1.0)
    base = [g["lr"] for g in groups]
    static = [torch.empty_like(p) for p in ps]   # static gradient buffers (as GraphedStep provides): the descriptor table is built once
    for p, s in zip(ps, static):
        p.grad = s
    for k in range(4):
        for grp, b in zip(opt.param_groups, base):
            grp["lr"] = b * lr_scale[k]
        for s, gr in zip(static, grads[k]):
            s.copy_(gr)
        opt.step()
        assert abs(float(opt.last_total_norm) / float(G["norm%d" % k]) - 1) < 1e-5
        for i, p in enumerate(ps):
            ref = G["p%d_step%d" % (i, k)]
            assert np.abs(p.detach().cpu().numpy() - ref).max() <= 4e-6 * np.abs(ref).max(), (k, i)
        assert all(torch.equal(s.cpu(), gr) for s, gr in zip(static, grads[k]))          # gradients are not modified by the clip
    for i, p in enumerate(ps):
        assert np.abs(opt.state[p]["exp_avg"].cpu().numpy() - G["m%d" % i]).max() <= 4e-6 * np.abs(G["m%d" % i]).max()
        assert np.abs(opt.state[p]["exp_avg_sq"].cpu().numpy() - G["v%d" % i]).max() <= 4e-6 * np.abs(G["v%d" % i]).max()
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and sd["state"][0]["step"] == 4


@pytest.mark.gpu
def test_dropout_contract_on_the_device():
    """vlb_dropout_mask / vlb_dropout against oracle/philox.py: bit-exact masks, kept values scaled by 1/(1-p)"""
    import philox
    import vlbert_b200
    L = vlbert_b200._lib
    st = torch.cuda.current_stream().cuda_stream
    for (n, p, seed, site, step) in ((1003, 0.1, 1234567890123, 5, 17), (6464 * 768, 0.1, 42, 13, 3), (7, 0.5, 1, 0, 0)):
        keep = torch.empty(n, dtype=torch.uint8, device="cuda")
        L.check(L.lib().vlb_dropout_mask(keep.data_ptr(), n, p, seed, site, step, st))
        ref = philox.keep_mask((n,), p, seed, site, step)
        assert np.array_equal(keep.cpu().numpy().astype(bool), ref)
        x = torch.randn(n, device="cuda")
        y = torch.empty_like(x)
        L.check(L.lib().vlb_dropout(x.data_ptr(), y.data_ptr(), n, 0, p, seed, site, step, st))
        want = torch.where(torch.from_numpy(ref).cuda(), x * (1.0 / (1.0 - np.float32(p))), torch.zeros_like(x))
        assert torch.allclose(y, want, rtol=1e-6, atol=0)
        xb = x.to(torch.bfloat16)
        yb = torch.empty_like(xb)
        L.check(L.lib().vlb_dropout(xb.data_ptr(), yb.data_ptr(), n, 1, p, seed, site, step, st))
        assert torch.equal(yb == 0, ~torch.from_numpy(ref).cuda() | (xb == 0))
