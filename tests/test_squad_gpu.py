"""SQuAD fine-tuning step (BASELINE.json configs[3]) through deeplearningexamples_b200.squad at seq 384:
loss and every gradient against the CPU oracle (run_squad.py:1062-1071 loss over oracle.bert_model + the QA head, modeling.py:1366-1369),
then one full step (clip + FusedAdam + device-side linear schedule) eagerly and as a replayed CUDA graph."""
import pytest
import torch

pytestmark = pytest.mark.gpu
bf = torch.bfloat16
CFG = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=1024, vocab_size=1024,
           max_position_embeddings=512, type_vocab_size=2, hidden_act="gelu", initializer_range=0.02,
           hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)


def _setup(B=4, S=384):
    from deeplearningexamples_b200 import squad as SQ
    from oracle import bert_oracle as O
    sd = O.bf16_representable_params(CFG, seed=5)
    g = torch.Generator().manual_seed(6)
    qa_w = (torch.randn(2, 256, generator=g) * 0.05).to(bf).float()
    qa_b = (torch.randn(2, generator=g) * 0.05).to(bf).float()
    full = {k: v for k, v in sd.items() if k.startswith("bert.")}
    full["qa_outputs.weight"], full["qa_outputs.bias"] = qa_w, qa_b
    batch = SQ.synthetic_squad_batch(B, S, CFG["vocab_size"], seed=8, full_mask=False)
    batch["start_positions"][0] = S + 5                          # out-of-span answer: clamped to S and ignored (run_squad.py:1064-1068)
    return SQ, O, sd, full, qa_w, qa_b, batch


def test_squad_loss_and_gradients_vs_oracle():
    SQ, O, sd, full, qa_w, qa_b, batch = _setup()
    B, S = batch["input_ids"].shape
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    w, b_ = qa_w.clone().requires_grad_(True), qa_b.clone().requires_grad_(True)
    seq, _ = O.bert_model(sdo, CFG, batch["input_ids"], batch["segment_ids"], batch["input_mask"])
    logits = torch.nn.functional.linear(seq, w, b_)
    s_ref, e_ref = logits[..., 0], logits[..., 1]
    lf = torch.nn.CrossEntropyLoss(ignore_index=S)
    loss_ref = (lf(s_ref, batch["start_positions"].clamp(0, S)) + lf(e_ref, batch["end_positions"].clamp(0, S))) / 2
    loss_ref.backward()
    model, opt, sched = SQ.prepare_squad_model_and_optimizer(CFG, torch.device("cuda", 0), state_dict=full, total_steps=100)
    model.train()
    bd = {k: v.cuda() for k, v in batch.items()}
    s_log, e_log = model(bd["input_ids"], bd["segment_ids"], bd["input_mask"])
    rel = lambda a, c: ((a.float().cpu() - c.float()).norm() / c.float().norm().clamp_min(1e-20)).item()
    assert rel(s_log, s_ref.detach()) < 1e-2 and rel(e_log, e_ref.detach()) < 1e-2
    loss = SQ.squad_loss(s_log, e_log, bd["start_positions"], bd["end_positions"])
    assert abs(loss.item() - loss_ref.item()) < 5e-3 * loss_ref.item()
    loss.backward()
    for k, p in model.named_parameters():
        if "pooler" in k:
            assert p.grad is None
            continue
        want = w.grad if k == "qa_outputs.weight" else b_.grad if k == "qa_outputs.bias" else sdo[k].grad
        # analytically zero gradients (both sides are rounding noise): key biases (softmax shift invariance), and -- because every row of
        # d(loss)/d(logits) sums to zero (softmax minus one-hot) -- the QA bias and the bias of the last LayerNorm in front of the linear head
        if k.endswith("key.bias") or want is None or k in ("qa_outputs.bias", "bert.encoder.layer.%d.output.LayerNorm.bias" % (CFG["num_hidden_layers"] - 1)):
            continue
        got = p.grad.float().cpu()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
        assert cos > 0.999 and rel(got, want) < 3e-2, (k, cos, rel(got, want))


@pytest.mark.parametrize("graphs", [False, True])
def test_squad_training_step_runs_eagerly_and_as_a_graph(graphs):
    from deeplearningexamples_b200 import ops, training as T
    SQ, O, sd, full, qa_w, qa_b, batch = _setup()
    cfg = dict(CFG, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    ops.manual_seed(3)
    model, opt, sched = SQ.prepare_squad_model_and_optimizer(cfg, torch.device("cuda", 0), state_dict=full, total_steps=50, learning_rate=1e-4)
    model.train()
    bd = {k: v.cuda() for k, v in batch.items()}
    loss_acc = torch.zeros((), device="cuda")
    step = lambda: SQ.squad_training_step(model, opt, sched, bd, loss_acc)
    n = 6
    if graphs:
        g = T.capture_step_graph(step, warmup_iters=3)
        loss_acc.zero_()
        for _ in range(n):
            g.replay()
        done = 3 + n
    else:
        for _ in range(n):
            step()
        done = n
    torch.cuda.synchronize()
    assert int(opt.param_groups[0]["step"].item()) == done and opt._found_inf.item() == 0.0
    lr = float(opt.param_groups[0]["lr"].item())
    progress = done / 50                                                       # the schedule saw step counter done-1 -> progress done/total
    factor = progress / 0.1 if progress < 0.1 else max((progress - 1.0) / (0.1 - 1.0), 0.0)
    assert abs(lr - 1e-4 * factor) < 1e-9 + 1e-3 * lr
    assert 0 < loss_acc.item() / n < 7.0
    ops.check_device_errors()
