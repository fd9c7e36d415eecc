"""Import-compatible stand-in for NVIDIA/lddl (absent offline; the reference pins no version: `pip install git+https://github.com/NVIDIA/lddl.git`,
PyTorch/LanguageModeling/BERT/Dockerfile:32).  Only the call site the pretraining driver uses exists: lddl.torch.get_bert_pretrain_data_loader
(run_pretraining.py:557-570)."""
