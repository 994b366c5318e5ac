"""Device-side task glue: the tensor bookkeeping the reference's task modules do around the hot path,
restated without boolean-mask assignments, `.item()` / `.tolist()` reads or Python loops over samples, so that with a static
`max_len` a whole step can be captured into a CUDA graph.  Index arithmetic only (cumulative sums, `scatter_`, `gather`): the
results are bit-identical to the reference functions; with `max_len=None` the width is computed like the reference does (one
host read) and the returned tensors have exactly the reference's shapes.

  pack_question_answer      <- ResNetVLBERT.prepare_text_from_qa   (vqa/modules/resnet_vlbert_for_vqa.py:141-167,
                                                                    vcr/modules/resnet_vlbert_for_vcr.py:135-164)
  pad_sequence              <- common/utils/pad_sequence.py:4-17   (called with `box_mask.sum(1).tolist()`, common/fast_rcnn.py:178)
  collect_obj_reps          <- ResNetVLBERT._collect_obj_reps       (vqa/...:122-139; already free of host reads, restated for callers)

`install(module_class, max_len=None)` swaps the methods on a reference task-module class, including the VCR module's
`prepare_text_from_aq` / `prepare_text_from_qa_onesent` variants (`dropin.install` leaves the task modules themselves untouched).
"""
import torch


def _scatter_tokens(dst, src, mask, first_col, dump_col):
    """dst[r, first_col[r] + k] = k-th element of src[r] whose mask is set; elements without mask go to the dump column."""
    rank = mask.long().cumsum(1) - 1
    col = torch.where(mask.bool(), first_col + rank, torch.full_like(rank, dump_col))
    dst.scatter_(1, col, src.to(dst.dtype))


def pack_question_answer(question, question_tags, question_mask, answer, answer_tags, answer_mask, cls_id, sep_id, max_len=None,
                         one_sentence=False):
    """[CLS] question [SEP] answer [SEP] per row, left-packed from masked token tensors
    (one_sentence: [CLS] question answer [SEP] with token type 0 throughout, the QA_ONE_SENT form).

    question / question_tags / question_mask: [R, Lq]; answer / answer_tags / answer_mask: [R, La] (R = any flattened batch).
    Returns (input_ids [R, W], input_type_ids [R, W], text_tags [R, W], input_mask [R, W] uint8, a_end [R]) with
    W = max_len, or max(question length + answer length) + 3 when max_len is None (the reference's width; one host read).
    A static max_len must be an upper bound of that value (e.g. Lq + La + 3): the extra columns carry input_mask = 0."""
    R = question.shape[0]
    dev = question.device
    q_len = question_mask.long().sum(1, keepdim=True)
    a_len = answer_mask.long().sum(1, keepdim=True)
    gap = 0 if one_sentence else 1                      # the [SEP] between the two segments
    if max_len is None:
        max_len = int((q_len + a_len).max()) + 2 + gap
    W = int(max_len)
    q_end = 1 + q_len
    a_end = q_end + gap + a_len
    j = torch.arange(W, device=dev)[None, :]
    input_mask = (j <= a_end).to(torch.uint8)
    input_type_ids = (((j > q_end) & (j <= a_end)) if not one_sentence else torch.zeros_like(j.expand(R, W), dtype=torch.bool)).to(question.dtype)
    # one spare column (index W) absorbs the writes of masked-out source tokens
    ids = torch.zeros((R, W + 1), dtype=question.dtype, device=dev)
    tags = torch.zeros((R, W + 1), dtype=question.dtype, device=dev)
    ids[:, 0] = cls_id
    sep = torch.full((R, 1), sep_id, dtype=question.dtype, device=dev)
    if not one_sentence:
        ids.scatter_(1, q_end.clamp(max=W), sep)
    ids.scatter_(1, a_end.clamp(max=W), sep)
    _scatter_tokens(ids, question, question_mask, torch.ones_like(q_end), W)
    _scatter_tokens(ids, answer, answer_mask, q_end + gap, W)
    _scatter_tokens(tags, question_tags, question_mask, torch.ones_like(q_end), W)
    _scatter_tokens(tags, answer_tags, answer_mask, q_end + gap, W)
    return ids[:, :W], input_type_ids, tags[:, :W], input_mask, a_end.squeeze(1)


def prepare_text_from_qa(question, question_tags, question_mask, answer, answer_tags, answer_mask, cls_id, sep_id, max_len=None):
    """VQA form (vqa/modules/resnet_vlbert_for_vqa.py:141-167): 2-D inputs, returns
    (input_ids, input_type_ids, text_tags, input_mask, a_end - 1)."""
    ids, types, tags, mask, a_end = pack_question_answer(question, question_tags, question_mask, answer, answer_tags, answer_mask,
                                                         cls_id, sep_id, max_len)
    return ids, types, tags, mask, a_end - 1


def prepare_text_from_qa_choices(question, question_tags, question_mask, answers, answers_tags, answers_mask, cls_id, sep_id,
                                 max_len=None, order="qa"):
    """VCR form (vcr/modules/resnet_vlbert_for_vcr.py:135-164): question* [B, Lq] (question_tags may already be [B, C, Lq]),
    answers* [B, C, La]; every answer choice gets its own copy of the question.  Returns four [B, C, W] tensors.
    order: "qa" (default, :135-164), "aq" = answer first (`prepare_text_from_aq`, :195-225, NETWORK.ANSWER_FIRST),
    "qa_onesent" = no separator between question and answer (`prepare_text_from_qa_onesent`, :166-193, NETWORK.QA_ONE_SENT)."""
    B, C, La = answers.shape
    Lq = question.shape[-1]

    def rows(t):
        if t.dim() == 2:
            t = t[:, None, :].expand(B, C, t.shape[-1])
        return t.reshape(B * C, t.shape[-1])

    q = (rows(question), rows(question_tags), rows(question_mask))
    a = (answers.reshape(B * C, La), answers_tags.reshape(B * C, La), answers_mask.reshape(B * C, La))
    first, second = (a, q) if order == "aq" else (q, a)
    ids, types, tags, mask, _ = pack_question_answer(*first, *second, cls_id, sep_id, max_len, one_sentence=(order == "qa_onesent"))
    W = ids.shape[1]
    return ids.view(B, C, W), types.view(B, C, W), tags.view(B, C, W), mask.view(B, C, W)


def pad_sequence(sequence, lengths, max_len=None):
    """[sum b, ...] -> [len(lengths), max b, ...], zero padded (common/utils/pad_sequence.py:4-17).  `lengths` may be a tensor
    (no host read when max_len is given) or the list the reference passes."""
    if not torch.is_tensor(lengths):
        if max_len is None:
            max_len = max(lengths) if len(lengths) else 0
        lengths = torch.as_tensor(lengths, dtype=torch.long, device=sequence.device)
    lengths = lengths.long()
    if max_len is None:
        max_len = int(lengths.max()) if lengths.numel() else 0
    W = int(max_len)
    n = lengths.numel()
    if n == 0 or W == 0 or sequence.shape[0] == 0:
        return sequence.new_zeros((n, W) + tuple(sequence.shape[1:]))
    start = lengths.cumsum(0) - lengths
    j = torch.arange(W, device=sequence.device)[None, :]
    valid = j < lengths[:, None]
    idx = (start[:, None] + j).clamp(max=sequence.shape[0] - 1)
    out = sequence[idx.reshape(-1)].view((n, W) + tuple(sequence.shape[1:]))
    return torch.where(valid.view((n, W) + (1,) * (sequence.dim() - 1)), out, torch.zeros((), dtype=out.dtype, device=out.device))


def collect_obj_reps(span_tags, object_reps):
    """object_reps[b, max(span_tags[b, ...], 0)] (vqa/modules/resnet_vlbert_for_vqa.py:122-139)."""
    tags = span_tags.clamp(min=0)
    B = tags.shape[0]
    row = torch.arange(B, device=tags.device).view((B,) + (1,) * (tags.dim() - 1)).expand_as(tags)
    return object_reps[row.reshape(-1), tags.reshape(-1)].view(tuple(tags.shape) + (object_reps.shape[-1],))


def install(module_class, max_len=None):
    """Replace `prepare_text_from_qa` and `_collect_obj_reps` of a reference task-module class (VQA or VCR `ResNetVLBERT`)
    by the device-side versions; `max_len`: optional static width (CUDA-graph capture)."""
    import inspect
    choices = "answers" in inspect.signature(module_class.prepare_text_from_qa).parameters

    def prepare(self, question, question_tags, question_mask, answer, answer_tags, answer_mask):
        cls_id, sep_id = self.tokenizer.convert_tokens_to_ids(['[CLS]', '[SEP]'])
        fn = prepare_text_from_qa_choices if choices else prepare_text_from_qa
        return fn(question, question_tags, question_mask, answer, answer_tags, answer_mask, cls_id, sep_id, max_len)

    module_class.prepare_text_from_qa = prepare
    if choices:
        for name, order in (("prepare_text_from_aq", "aq"), ("prepare_text_from_qa_onesent", "qa_onesent")):
            if hasattr(module_class, name):
                def variant(self, question, question_tags, question_mask, answers, answers_tags, answers_mask, _order=order):
                    cls_id, sep_id = self.tokenizer.convert_tokens_to_ids(['[CLS]', '[SEP]'])
                    return prepare_text_from_qa_choices(question, question_tags, question_mask, answers, answers_tags, answers_mask,
                                                        cls_id, sep_id, max_len, order=_order)
                setattr(module_class, name, variant)
    module_class._collect_obj_reps = lambda self, span_tags, object_reps: collect_obj_reps(span_tags, object_reps)
    return module_class
