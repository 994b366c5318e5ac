"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/task_glue.npz by executing the UNMODIFIED reference task-module helpers
(/root/reference via oracle/ref_shim.py) on seeded synthetic inputs.  Run in the build container:

    python oracle/make_golden_glue.py

  vqa_*   ResNetVLBERT.prepare_text_from_qa      vqa/modules/resnet_vlbert_for_vqa.py:141-167
  vcr_*   ResNetVLBERT.prepare_text_from_qa      vcr/modules/resnet_vlbert_for_vcr.py:135-164
  obj_*   ResNetVLBERT._collect_obj_reps         vqa/modules/resnet_vlbert_for_vqa.py:122-139
  pad_*   common.utils.pad_sequence.pad_sequence common/utils/pad_sequence.py:4-17
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")
CLS, SEP = 101, 102


class _Tok(object):
    def convert_tokens_to_ids(self, toks):
        return [{"[CLS]": CLS, "[SEP]": SEP}[t] for t in toks]


class _Self(object):
    tokenizer = _Tok()


def glue_inputs(seed, B=5, C=4, Lq=9, La=7):
    """Shared with tests/test_task_glue.py: masked, left-aligned token tensors with ragged lengths (incl. an empty answer)."""
    g = torch.Generator().manual_seed(seed)
    q_len = torch.randint(1, Lq + 1, (B,), generator=g)
    a_len = torch.randint(0, La + 1, (B, C), generator=g)
    a_len[0, 0] = 0
    a_len[1, 1] = La
    q_len[1] = Lq
    question = torch.randint(1000, 30000, (B, Lq), generator=g)
    q_tags = torch.randint(-1, 6, (B, Lq), generator=g)
    q_mask = torch.arange(Lq)[None, :] < q_len[:, None]
    question = question * q_mask
    answers = torch.randint(1000, 30000, (B, C, La), generator=g)
    a_tags = torch.randint(-1, 6, (B, C, La), generator=g)
    a_mask = torch.arange(La)[None, None, :] < a_len[:, :, None]
    answers = answers * a_mask
    return question, q_tags, q_mask, answers, a_tags, a_mask


def main():
    ref_shim.install()
    import importlib
    vqa = importlib.import_module("vqa.modules.resnet_vlbert_for_vqa")
    vcr = importlib.import_module("vcr.modules.resnet_vlbert_for_vcr")
    from common.utils.pad_sequence import pad_sequence
    out = {}
    q, qt, qm, a, at, am = glue_inputs(11)
    r = vqa.ResNetVLBERT.prepare_text_from_qa(_Self(), q, qt, qm, a[:, 0], at[:, 0], am[:, 0])
    for n, t in zip(("ids", "types", "tags", "mask", "ans_pos"), r):
        out["vqa_" + n] = t.numpy()
    qt3 = qt.repeat(1, a.shape[1]).view(qt.shape[0], a.shape[1], -1)          # as vcr/modules/resnet_vlbert_for_vcr.py:266
    r = vcr.ResNetVLBERT.prepare_text_from_qa(_Self(), q, qt3, qm, a, at, am)
    for n, t in zip(("ids", "types", "tags", "mask"), r):
        out["vcr_" + n] = t.numpy()
    g = torch.Generator().manual_seed(5)
    tags = torch.randint(-1, 6, (5, 4, 11), generator=g)
    reps = torch.randn(5, 6, 8, generator=g)
    out["obj_tags"], out["obj_reps"] = tags.numpy(), reps.numpy()
    out["obj_out"] = vqa.ResNetVLBERT._collect_obj_reps(_Self(), tags, reps).numpy()
    lengths = [3, 0, 5, 1]
    seq = torch.randn(sum(lengths), 2, 3, generator=g)
    out["pad_seq"], out["pad_lengths"] = seq.numpy(), np.asarray(lengths)
    out["pad_out"] = pad_sequence(seq, lengths).numpy()
    np.savez_compressed(os.path.join(GOLD, "task_glue.npz"), **out)
    print("wrote task_glue.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
