"""Host-side checks of the driver mirror (no GPU): the reference's full flag surface parses, derived values follow
run_pretraining.py:315-321,363, and the stand-in loader yields the lddl batch format, sharded by rank."""
import pytest
import torch

REFERENCE_FLAGS = ["--input_dir", "--config_file", "--output_dir", "--vocab_file", "--init_checkpoint", "--max_seq_length",
                   "--max_predictions_per_seq", "--train_batch_size", "--learning_rate", "--num_train_epochs", "--max_steps",
                   "--warmup_proportion", "--local_rank", "--seed", "--gradient_accumulation_steps", "--fp16", "--amp", "--loss_scale",
                   "--log_freq", "--checkpoint_activations", "--resume_from_checkpoint", "--resume_step", "--num_steps_per_checkpoint",
                   "--skip_checkpoint", "--phase2", "--resume_phase2", "--allreduce_post_accumulation",
                   "--allreduce_post_accumulation_fp16", "--phase1_end_step", "--init_loss_scale", "--do_train", "--json-summary",
                   "--use_env", "--disable_progress_bar", "--steps_this_run", "--profile", "--profile-start", "--num_workers",
                   "--no_dense_sequence_output", "--disable_jit_fusions", "--cuda_graphs"]     # run_pretraining.py:145-313


def test_every_reference_flag_is_accepted():
    from deeplearningexamples_b200 import run_pretraining as rp
    known = set()
    import argparse
    # introspect our parser
    try:
        rp.parse_arguments(["--help"])
    except SystemExit:
        pass
    args = rp.parse_arguments(["--config_file", "c.json", "--output_dir", "o", "--amp", "--max_steps", "10", "--do_train",
                                "--allreduce_post_accumulation", "--allreduce_post_accumulation_fp16", "--json-summary", "x.json",
                                "--profile-start", "3", "--phase2", "--resume_from_checkpoint", "--disable_jit_fusions"])
    assert args.fp16 and args.steps_this_run == 10 and args.json_summary == "x.json" and args.profile_start == 3
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), pytest.raises(SystemExit):
        rp.parse_arguments(["--help"])
    helptext = buf.getvalue()
    for flag in REFERENCE_FLAGS:
        assert flag in helptext, flag


def test_synthetic_loader_is_lddl_shaped_and_rank_sharded():
    from deeplearningexamples_b200.run_pretraining import SyntheticPretrainLoader
    a = SyntheticPretrainLoader(4, 128, 20, 30528, steps_per_epoch=5, base_seed=42, rank=0)
    b = SyntheticPretrainLoader(4, 128, 20, 30528, steps_per_epoch=5, base_seed=42, rank=1)
    assert len(a) == 5
    batches = list(a)
    assert len(batches) == 5
    for bt in batches:
        assert set(bt) == {"input_ids", "token_type_ids", "attention_mask", "labels", "next_sentence_labels"}
        assert all(v.dtype == torch.int64 for v in bt.values())
        assert bt["input_ids"].shape == (4, 128) and bt["next_sentence_labels"].shape == (4,)
        assert ((bt["labels"] != -1).sum(1) == 20).all() and (bt["input_ids"][:, 0] == 101).all()
    assert not torch.equal(next(iter(a))["input_ids"], next(iter(b))["input_ids"])


def test_configs_match_reference_values():
    import json, os
    from deeplearningexamples_b200 import modeling
    d = os.path.join(os.path.dirname(modeling.__file__), "bert_configs")
    large = modeling.BertConfig.from_json_file(os.path.join(d, "large.json"))
    assert (large.hidden_size, large.num_hidden_layers, large.num_attention_heads, large.intermediate_size, large.vocab_size) == (1024, 24, 16, 4096, 30522)
    assert json.loads(large.to_json_string())["max_position_embeddings"] == 512


def test_from_pretrained_maps_legacy_keys_and_prefixes(tmp_path):
    """reference modeling.py:655-786: directory / .tar.gz archive with bert_config.json + pytorch_model.bin, legacy key names
    (LayerNorm gamma/beta, intermediate.dense., pooler.dense.), `bert.` prefix dropped for the bare BertModel, (model, config) returned."""
    import json
    import tarfile
    import torch
    from deeplearningexamples_b200 import modeling as M
    cfg = dict(vocab_size_or_config_json_file=96, hidden_size=64, num_hidden_layers=2, num_attention_heads=1, intermediate_size=128,
               hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=32,
               type_vocab_size=2, initializer_range=0.02)
    torch.manual_seed(3)
    src = M.BertForPreTraining(M.BertConfig(**cfg))
    want = {k: v.clone() for k, v in src.state_dict().items()}

    def legacy(k):                                   # what a checkpoint written by the older reference code calls this tensor
        if "LayerNorm.weight" in k:
            return k.replace("LayerNorm.weight", "LayerNorm.gamma")
        if "LayerNorm.bias" in k:
            return k.replace("LayerNorm.bias", "LayerNorm.beta")
        return k.replace("intermediate.dense_act.", "intermediate.dense.").replace("pooler.dense_act.", "pooler.dense.")
    old = {legacy(k): v for k, v in want.items()}
    assert any("gamma" in k for k in old) and any("intermediate.dense." in k for k in old) and any("pooler.dense." in k for k in old)
    d = tmp_path / "ckpt"
    d.mkdir()
    cfg_file = dict(cfg); cfg_file["vocab_size"] = cfg_file.pop("vocab_size_or_config_json_file")
    (d / "bert_config.json").write_text(json.dumps(cfg_file))
    torch.save(old, d / "pytorch_model.bin")

    model, config = M.BertForPreTraining.from_pretrained(str(d))
    assert isinstance(model, M.BertForPreTraining) and config.hidden_size == 64 and config.vocab_size == 96
    got = model.state_dict()
    assert set(got) == set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k

    tar = tmp_path / "ckpt.tar.gz"                   # archive form
    with tarfile.open(tar, "w:gz") as t:
        t.add(d / "bert_config.json", arcname="bert_config.json")
        t.add(d / "pytorch_model.bin", arcname="pytorch_model.bin")
    bare, _ = M.BertModel.from_pretrained(str(tar))  # pretraining checkpoint into the bare encoder: `bert.` prefix dropped, heads unused
    for k, v in bare.state_dict().items():
        assert torch.equal(v, want["bert." + k]), k

    qa, _ = M.BertForQuestionAnswering.from_pretrained(str(d), state_dict=dict(old))      # explicit state_dict; qa_outputs stays initialised
    assert torch.equal(qa.state_dict()["bert.embeddings.word_embeddings.weight"], want["bert.embeddings.word_embeddings.weight"])
    assert M.BertModel.from_pretrained("bert-large-uncased") is None                       # names need a download: error logged, None
    bad = dict(old); bad["bert.embeddings.word_embeddings.weight"] = torch.zeros(5, 64)
    with pytest.raises(RuntimeError):
        M.BertForPreTraining.from_pretrained(str(d), state_dict=bad)
