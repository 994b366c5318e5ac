"""Seeded synthetic inputs / weights shared by the tests, bench.py and smoke().

`synth_vlbert_inputs` and `seeded_state_dict` must stay call-for-call identical to
oracle/make_golden.py (which produced tests/golden/*.npz from the real reference)."""
import torch


def synth_vlbert_inputs(B, T, R, H, vocab, seed, ragged=True):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(min(1000, vocab // 2), vocab, (B, T), generator=g)
    types = torch.randint(0, 2, (B, T), generator=g)
    tvis = torch.randn(B, T, H, generator=g)
    ovl = torch.randn(B, R, 2 * H, generator=g)
    tmask = torch.ones(B, T, dtype=torch.bool)
    omask = torch.ones(B, R, dtype=torch.bool)
    if ragged:
        for b in range(B):
            tl = int(torch.randint(max(1, T // 2), T + 1, (1,), generator=g))
            ol = int(torch.randint(1, R + 1, (1,), generator=g))
            if b == 0:
                tl, ol = T, R
            tmask[b, tl:] = False
            omask[b, ol:] = False
        ids = ids * tmask
    return ids, types, tvis, tmask, ovl, omask


def seeded_state_dict(model, seed, std=0.02):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        if "LayerNorm.weight" in k or k.startswith("visual_ln") and k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.02 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = std * torch.randn(v.shape, generator=g)
    return sd


def vlbert_loss(layers, pooled, seed):
    """The scalar the golden gradients were taken of (oracle/make_golden.py:run_vlbert)."""
    g = torch.Generator().manual_seed(seed + 1)
    gw = [torch.randn(l.shape, generator=g).to(l.device) for l in layers]
    gp = torch.randn(pooled.shape, generator=g).to(pooled.device)
    loss = sum((l.float() * w).sum() for l, w in zip(layers, gw)) * 0.5 + (pooled.float() * gp).sum()
    return loss + (layers[-1].float() * gw[-1]).sum() * 0.5
