"""Pin the CPU oracle (oracle/bert_oracle.py) against vectors produced by the reference's own
modeling.py (tests/golden/make_golden.py) and the reference repo's only numeric KAT."""
import os

import pytest
import torch

from oracle import bert_oracle as O


def test_gelu_known_answer():
    # TensorFlow2/LanguageModeling/BERT/official/modeling/activations/gelu_test.py:29-32
    x = torch.tensor([[0.25, 0.0, -0.25], [-1.0, -2.0, 3.0]])
    want = torch.tensor([[0.14967535, 0.0, -0.10032465], [-0.15880796, -0.04540223, 2.9963627]])
    torch.testing.assert_close(O.gelu_tanh(x), want, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(torch.nn.functional.gelu(x, approximate="tanh"), want, rtol=1e-6, atol=1e-7)


@pytest.fixture(scope="module")
def tiny(golden_dir):
    return torch.load(os.path.join(golden_dir, "bert_tiny_golden.pt"), weights_only=False)


def test_tiny_forward_matches_reference(tiny):
    cfg, sd, batch = tiny["cfg"], tiny["state_dict"], tiny["batch"]
    loss, scores, nsp, seq = O.forward_loss(sd, cfg, batch)
    torch.testing.assert_close(scores, tiny["scores"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(nsp, tiny["nsp"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(loss, tiny["loss"], rtol=1e-6, atol=1e-6)


def test_tiny_per_layer_activations(tiny):
    cfg, sd, batch = tiny["cfg"], tiny["state_dict"], tiny["batch"]
    acts = tiny["acts"]
    emb = O.embeddings(sd, batch["input_ids"], batch["token_type_ids"])
    torch.testing.assert_close(emb, acts["embeddings"], rtol=1e-5, atol=1e-5)
    ext = O.extended_mask(batch["attention_mask"], emb.dtype)
    h = emb.transpose(0, 1)
    for i in range(cfg["num_hidden_layers"]):
        p = f"bert.encoder.layer.{i}."
        ctx = O.self_attention(sd, p + "attention.self.", h, ext, cfg["num_attention_heads"])
        torch.testing.assert_close(ctx, acts[f"layer{i}.ctx"], rtol=1e-5, atol=1e-5)
        h = O.bert_layer(sd, p, h, ext, cfg["num_attention_heads"])
        torch.testing.assert_close(h, acts[f"layer{i}.out"], rtol=1e-5, atol=2e-5)


def test_tiny_embedding_gather_bit_exact(tiny):
    """Index gathers are integer work: the gathered rows must be bit-identical."""
    sd, batch = tiny["state_dict"], tiny["batch"]
    w = sd["bert.embeddings.word_embeddings.weight"]
    rows = w[batch["input_ids"]]
    assert torch.equal(rows, torch.nn.functional.embedding(batch["input_ids"], w))


def test_tiny_backward_matches_reference(tiny):
    cfg, batch = tiny["cfg"], tiny["batch"]
    sd = {k: v.clone().requires_grad_(True) for k, v in tiny["state_dict"].items()}
    loss, *_ = O.forward_loss(sd, cfg, batch)
    loss.backward()
    for k, g in tiny["grads"].items():
        if k == "cls.predictions.decoder.weight":
            continue   # tied: same tensor as the word-embedding table
        torch.testing.assert_close(sd[k].grad, g, rtol=2e-4, atol=2e-6, msg=lambda m, k=k: f"{k}: {m}")


def test_base_config1_matches_reference(golden_dir):
    """BASELINE.json configs[0] shape: BERT-base, B=4, S=128, fp32, CPU."""
    gold = torch.load(os.path.join(golden_dir, "bert_base_golden.pt"), weights_only=False)
    cfg = gold["cfg"]
    sd = O.init_params(cfg, seed=gold["seed"])
    batch = O.synthetic_batch(4, 128, cfg["vocab_size"], 20, seed=gold["batch_seed"], full_mask=True)
    with torch.no_grad():
        loss, scores, nsp, _ = O.forward_loss(sd, cfg, batch)
    torch.testing.assert_close(loss, gold["loss"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(scores[:8, :64], gold["scores_slice"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(nsp, gold["nsp"], rtol=1e-4, atol=1e-4)


def test_param_inventory_bert_large():
    """336 232 258 parameters in 398 tensors (V padded to 30528)."""
    shp = O.param_shapes(O.BERT_LARGE)
    n = sum(int(torch.tensor(s).prod()) for s in shp.values())
    assert len(shp) == 398 and n == 336_232_258


def test_small_kernel_config_matches_reference(golden_dir):
    """The config the GPU parity tests use (H=256, A=4 -> head 64, S=128): oracle == reference."""
    gold = torch.load(os.path.join(golden_dir, "bert_small_golden.pt"), weights_only=False)
    cfg = gold["cfg"]
    sd = {k: v.clone().requires_grad_(True) for k, v in O.bf16_representable_params(cfg, seed=gold["param_seed"]).items()}
    batch = O.synthetic_batch(2, 128, cfg["vocab_size"], 10, seed=gold["batch_seed"], full_mask=False)
    loss, scores, nsp, seq = O.forward_loss(sd, cfg, batch)
    torch.testing.assert_close(loss, gold["loss"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(scores, gold["scores"].float(), rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(seq, gold["seq_out"].float(), rtol=2e-3, atol=2e-3)
    loss.backward()
    for k, g in gold["grads"].items():
        torch.testing.assert_close(sd[k].grad, g.float(), rtol=5e-3, atol=1e-5, msg=lambda m, k=k: f"{k}: {m}")
    for k, n in gold["grad_norms"].items():
        if k != "cls.predictions.decoder.weight":
            assert float(sd[k].grad.norm()) == pytest.approx(float(n), rel=1e-3, abs=1e-6), k   # key.bias grad is analytically 0
