"""Model-level GPU parity AT THE BENCHMARKED WIDTHS (BASELINE.json configs[1]/[2]: H=1024, A=16, I=4096, V=30528) against
(a) golden vectors produced by the reference's own modeling.py (tests/golden/bert_large2_golden.pt, make_golden.py) and
(b) the CPU oracle run in-test on the same bf16-representable weights: forward logits, loss through the BENCHED criterion
(training.BertPretrainingCriterion), and every parameter gradient.  Two encoder layers (24 would repeat the same shapes),
B=2, ragged attention mask, S=512 and S=128.

Metrics: north_star's "within 1e-2 rel on bf16 logits" is applied to the relative L2 error
||got - want|| / ||want||; the max-norm error max|got - want| / max|want| is bounded at 2e-2 because one bf16 ulp (2^-8 = 0.39 %)
on a logit near the maximum already costs 0.4-0.8 % in that metric.  Index gathers are bit exact (tests/test_pointwise_gpu.py).
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
bf = torch.bfloat16


def _rel_l2(got, want):
    return ((got.float() - want.float()).norm() / want.float().norm().clamp_min(1e-30)).item()


def _rel_max(got, want):
    return ((got.float() - want.float()).abs().max() / want.float().abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def large2(golden_dir):
    from oracle import bert_oracle as O
    gold = torch.load(os.path.join(golden_dir, "bert_large2_golden.pt"), weights_only=False)
    sd = O.bf16_representable_params(gold["cfg"], seed=gold["param_seed"], std=gold["param_std"])
    return gold, sd


def _model(cfg, sd):
    from deeplearningexamples_b200 import modeling
    m = modeling.BertForPreTraining(modeling.BertConfig.from_dict(dict(cfg)), sequence_output_is_dense=True)
    full = dict(sd)
    full["cls.predictions.decoder.weight"] = sd["bert.embeddings.word_embeddings.weight"]
    missing, unexpected = m.load_state_dict(full, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return m.cuda().to(bf).train()


@pytest.mark.parametrize("S", [512, 128])
def test_bert_large_widths_forward_loss_backward_vs_reference_golden_and_oracle(large2, S):
    from deeplearningexamples_b200.training import BertPretrainingCriterion
    from oracle import bert_oracle as O
    gold, sd = large2
    cfg, case = gold["cfg"], gold["cases"][S]
    batch = O.synthetic_batch(2, S, cfg["vocab_size"], case["max_pred"], seed=case["batch_seed"], full_mask=False)

    # ---- CPU oracle on the same weights and batch (fp32 autograd = backward oracle)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss_o, scores_o, nsp_o, seq_o = O.forward_loss(sdo, cfg, batch)
    loss_o.backward()
    # the oracle itself is pinned to the reference here too (fp32 vs fp32)
    assert abs(loss_o.item() - case["loss"].item()) < 1e-4 * case["loss"].item()
    assert _rel_l2(scores_o.detach()[:, ::16], case["scores_strided"]) < 2e-3          # golden slice is stored in fp16

    # ---- product path
    m = _model(cfg, sd)
    crit = BertPretrainingCriterion(cfg["vocab_size"], sequence_output_is_dense=True)
    b = {k: v.cuda() for k, v in batch.items()}
    scores, nsp = m(b["input_ids"], b["token_type_ids"], b["attention_mask"], b["labels"])
    assert scores.dtype == bf and tuple(scores.shape) == tuple(scores_o.shape)
    sc = scores.float().cpu()
    # logits: vs the reference golden (strided columns + full-vocabulary logsumexp) and vs the oracle (every element)
    assert _rel_l2(sc[:, ::16], case["scores_strided"]) < 1e-2
    assert _rel_l2(sc, scores_o.detach()) < 1e-2 and _rel_max(sc, scores_o.detach()) < 2e-2
    lse = torch.logsumexp(sc, -1)
    assert (lse - case["scores_lse"]).abs().max().item() < 1e-2 * case["scores_lse"].abs().max().item()
    assert _rel_max(nsp.float().cpu(), case["nsp"]) < 2e-2
    # loss through the benched criterion (fp32 cross-entropy on the bf16 logits)
    loss = crit(scores, nsp, b["labels"], b["next_sentence_labels"])
    assert loss.dtype == torch.float32
    assert abs(loss.item() - case["loss"].item()) < 5e-3 * case["loss"].item(), (loss.item(), case["loss"].item())
    # ---- gradients: every parameter vs the oracle; norms + strided slices vs the reference golden
    loss.backward()
    named = dict(m.named_parameters())
    worst = {}
    for k, p in named.items():
        if k == "cls.predictions.decoder.weight":
            continue                                   # tied: same tensor as the word-embedding table
        want = sdo[k].grad
        assert p.grad is not None and want is not None, k
        got = p.grad.float().cpu()
        if k.endswith("key.bias"):
            continue                                   # analytically zero (softmax shift invariance): both sides are rounding noise
        cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
        worst[k] = (1 - cos, _rel_l2(got, want))
        assert cos > 0.999, (k, cos)
        assert _rel_l2(got, want) < 3e-2, (k, _rel_l2(got, want))
    for k, n in case["grad_norms"].items():
        if k == "cls.predictions.decoder.weight" or k.endswith("key.bias"):
            continue
        gn = named[k].grad.float().norm().item()
        assert abs(gn - n.item()) <= 3e-2 * n.item() + 1e-7, (k, gn, n.item())
    for k, g in case["grads_strided"].items():
        got = named[k].grad.float().cpu().reshape(-1)[::(1 if named[k].numel() <= 4096 else 97)]
        assert _rel_l2(got, g) < 3e-2, (k, _rel_l2(got, g))


def test_benched_criterion_equals_oracle_loss_and_ignores_padding_slots(large2):
    """training.BertPretrainingCriterion (the criterion bench.py and the driver use) vs oracle.pretraining_loss on the same
    logits: fp32 cross-entropy, same rows in the same order; with the static-size gather the surplus slots (index -1) are ignored
    and a full batch raises nothing while an overflowing one is reported."""
    from deeplearningexamples_b200 import _lib as L
    from deeplearningexamples_b200.training import BertPretrainingCriterion
    from oracle import bert_oracle as O
    gold, sd = large2
    cfg = gold["cfg"]
    case = gold["cases"][128]
    batch = O.synthetic_batch(2, 128, cfg["vocab_size"], case["max_pred"], seed=case["batch_seed"], full_mask=False)
    m = _model(cfg, sd)
    crit = BertPretrainingCriterion(cfg["vocab_size"], sequence_output_is_dense=True)
    b = {k: v.cuda() for k, v in batch.items()}
    with torch.no_grad():
        s_dyn, n_dyn = m(b["input_ids"], b["token_type_ids"], b["attention_mask"], b["labels"])
        want = O.pretraining_loss(s_dyn.float().cpu(), n_dyn.float().cpu(), batch["labels"], batch["next_sentence_labels"])
        got = crit(s_dyn, n_dyn, b["labels"], b["next_sentence_labels"])
        assert abs(got.item() - want.item()) < 1e-5 * abs(want.item())
        m.cls.static_masked_count = 2 * 32                     # 40 masked positions -> 24 padding slots
        s_st, n_st = m(b["input_ids"], b["token_type_ids"], b["attention_mask"], b["labels"])
        assert s_st.shape[0] == 64 and torch.equal(s_st[:40], s_dyn)
        got_st = crit(s_st, n_st, b["labels"], b["next_sentence_labels"])
        assert abs(got_st.item() - want.item()) < 1e-5 * abs(want.item())
        m.cls.check_mlm_overflow()                             # no overflow so far
        m.cls.static_masked_count = 16                         # fewer slots than masked positions: must be reported, not silent
        m(b["input_ids"], b["token_type_ids"], b["attention_mask"], b["labels"])
        with pytest.raises(L.DleError):
            m.cls.check_mlm_overflow()


def test_static_gather_backward_has_no_aliasing_race(large2):
    """surplus slots used to alias flat token 0 (duplicate scatter indices).  With -1 padding slots the gradient of the
    sequence output equals the dynamic-gather gradient bit for bit, also when token 0 itself is a masked position."""
    from deeplearningexamples_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(256, 1024, generator=g, device="cuda").to(bf).requires_grad_(True)
    labels = torch.full((256,), -1, device="cuda")
    labels[[0, 5, 77, 200]] = 3
    dy = torch.randn(8, 1024, generator=g, device="cuda").to(bf)
    idx_dyn = torch.nonzero(labels != -1).squeeze(-1)
    idx_st = torch.nonzero_static(labels != -1, size=8, fill_value=-1).squeeze(-1)
    y_dyn = ops.GatherRowsFn.apply(x, idx_dyn)
    y_dyn.backward(dy[:4])
    g_dyn = x.grad.clone()
    x.grad = None
    y_st = ops.GatherRowsFn.apply(x, idx_st)
    assert torch.equal(y_st[:4], y_dyn) and torch.count_nonzero(y_st[4:]) == 0
    y_st.backward(dy)
    assert torch.equal(x.grad, g_dyn)
    ops.check_device_errors()


def test_checkpoint_activations_replays_dropout_masks(large2):
    """BertEncoder.checkpointed_forward (reference modeling.py:471-489, --checkpoint_activations) with dropout ON: the recomputation
    in backward must regenerate the forward's masks (host seeds replayed, device step counter unchanged), so outputs and gradients
    equal those of the plain path run from the same RNG state."""
    from deeplearningexamples_b200 import modeling, ops
    from oracle import bert_oracle as O
    cfg = dict(hidden_size=256, num_hidden_layers=4, num_attention_heads=4, intermediate_size=1024, vocab_size=1024,
               max_position_embeddings=128, type_vocab_size=2, hidden_act="gelu", initializer_range=0.02,
               hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    sd = O.bf16_representable_params(cfg, seed=3)
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 128, cfg["vocab_size"], 10, seed=4, full_mask=False).items()}
    res = []
    for ckpt in (False, True):
        ops.manual_seed(99)
        m = _model(cfg, sd)
        m.checkpoint_activations(ckpt)
        assert m.bert.encoder._checkpoint_activations is ckpt
        scores, nsp = m(batch["input_ids"], batch["token_type_ids"], batch["attention_mask"], batch["labels"])
        flat = batch["labels"].view(-1)
        lf = torch.nn.CrossEntropyLoss(ignore_index=-1)
        loss = lf(scores.float(), flat[flat != -1]) + lf(nsp.float(), batch["next_sentence_labels"])
        loss.backward()
        res.append((scores.detach().clone(), {k: p.grad.detach().float().clone() for k, p in m.named_parameters()}))
    assert torch.equal(res[0][0], res[1][0])                    # same masks in forward
    for k in res[0][1]:
        if k.endswith("key.bias"):
            continue
        a, c = res[0][1][k], res[1][1][k]
        # identical arithmetic; only fp32 atomics of the split-K weight gradients may reorder
        assert (a - c).abs().max().item() <= 1e-2 * c.abs().max().item() + 1e-6, k
    # sanity: dropout really was on (eval differs)
    m.eval()
    with torch.no_grad():
        s_eval, _ = m(batch["input_ids"], batch["token_type_ids"], batch["attention_mask"], batch["labels"])
    assert not torch.equal(s_eval, res[1][0])
