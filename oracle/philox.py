"""TEST INFRASTRUCTURE ONLY -- Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) in numpy,
and the dropout-mask contract the fused kernels are specified against (DESIGN.md "Dropout contract").

The reference draws its dropout masks from torch's global generator (nn.Dropout in modeling.py:283,316,334,379 and
common/visual_linguistic_bert.py:75), so no implementation can reproduce the reference's masks; what CAN be pinned is that
the library drops exactly the elements this counter-based stream says, scales the kept ones by 1/(1-p), and uses the same
mask in forward and backward.  Contract (one Philox call covers four consecutive elements of a row-major tensor):

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (group & 0xffffffff, group >> 32, site, step)       group = linear_element_index // 4
    keep[4*group + j] = philox(counter, key)[j] >= floor(float32(p) * 2^32)     (j = 0..3)

2-D form (what every fused site uses): element (r, c) of a row-major [rows, cols] tensor belongs to
    group = r * ceil(cols / 4) + c // 4,  word = c % 4
which IS the linear contract whenever cols % 4 == 0 (all hidden-size tensors); the attention probabilities
(rows = (b, head, query), cols = S keys) pad every row to a multiple of four so that a row never shares a Philox call.

`site` numbers the dropout call sites of one step: 0 = embedding output [B*S, H], then per layer l: 1+3l = attention
probabilities [B*heads*S, S], 2+3l = self-output dense [B*S, H], 3+3l = output dense [B*S, H]; FastRCNN's obj_downsample
input [B*R, 4096] uses site 1000 with the FastRCNN module's own (seed, step).
`step` is the training step counter, so the masks differ every step and the backward regenerates them from (seed, step).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(counter, key, rounds=10):
    """counter: uint32 array [..., 4]; key: (k0, k1) -> uint32 array [..., 4]"""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    for _ in range(rounds):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c[3] ^ np.uint64(k1)) & MASK, lo0]
        with np.errstate(over="ignore"):
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return np.stack([x.astype(np.uint32) for x in c], -1)


def keep_mask(shape, p, seed, site, step):
    """boolean keep-mask of a row-major tensor of `shape` under the contract above"""
    n = int(np.prod(shape))
    groups = (n + 3) // 4
    g = np.arange(groups, dtype=np.uint64)
    ctr = np.stack([(g & MASK).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
                    np.full(groups, site, np.uint32), np.full(groups, step, np.uint32)], -1)
    r = philox4x32(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(-1)[:n]
    thresh = np.uint64(min(int(float(np.float32(p)) * 4294967296.0), 0xFFFFFFFF))   # p is a float32 in the C ABI
    return (r.astype(np.uint64) >= thresh).reshape(shape)


def dropout(x, p, seed, site, step):
    """x: numpy array; returns (y, keep) with y = x * keep / (1 - p)"""
    if p <= 0.0:
        return x, np.ones(x.shape, bool)
    keep = keep_mask(x.shape, p, seed, site, step)
    return (x * keep / np.float32(1.0 - p)).astype(x.dtype), keep


def keep_mask_2d(rows, cols, p, seed, site, step):
    """boolean keep-mask [rows, cols] under the 2-D contract (rows padded to a multiple of four columns)"""
    gpr = (cols + 3) // 4
    full = keep_mask((rows * gpr * 4,), p, seed, site, step).reshape(rows, gpr * 4)
    return full[:, :cols]


def dropout_2d(x2d, p, seed, site, step):
    """x2d: numpy [rows, cols] -> x * keep / (1 - p) under the 2-D contract"""
    if p <= 0.0:
        return x2d
    keep = keep_mask_2d(x2d.shape[0], x2d.shape[1], p, seed, site, step)
    return (x2d * keep * (np.float32(1.0) / np.float32(1.0 - np.float32(p)))).astype(x2d.dtype)
