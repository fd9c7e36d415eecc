"""Install recipe for the reference arm: copy the UNMODIFIED reference files of the pretraining path from /root/reference into
baseline/_ref/BERT/ (git-ignored, travels to the GPU box with the snapshot; never committed).  Run in the authoring container
(`python baseline/install_ref.py`, also called by __graft_entry__.build()).  The reference's CUDA extension is built separately by
oracle/build_ref.py.  Files (relative to PyTorch/LanguageModeling/BERT/):
    run_pretraining.py modeling.py schedulers.py utils.py file_utils.py lamb_amp_opt/fused_lamb/{__init__,fused_lamb}.py
    bert_configs/*.json run_squad.py tokenization.py tokenization_utils.py optimization.py
`pip install /root/reference` does not apply: the reference tree is not a Python package (no setup.py at its root; the BERT directory is
run from source inside NVIDIA's container, Dockerfile:34-36)."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/PyTorch/LanguageModeling/BERT"
DST = os.path.join(HERE, "_ref", "BERT")
FILES = ["run_pretraining.py", "modeling.py", "schedulers.py", "utils.py", "file_utils.py", "run_squad.py", "tokenization.py",
         "tokenization_utils.py", "optimization.py", "lamb_amp_opt/fused_lamb/__init__.py", "lamb_amp_opt/fused_lamb/fused_lamb.py",
         "bert_configs/large.json", "bert_configs/base.json"]


def installed():
    return os.path.exists(os.path.join(DST, "run_pretraining.py"))


def install(force=False):
    if not os.path.isdir(SRC):
        return installed()
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if force or not os.path.exists(d):
            shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    return True


if __name__ == "__main__":
    print("installed" if install(force="--force" in sys.argv) else "reference tree not present")
