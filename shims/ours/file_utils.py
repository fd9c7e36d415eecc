"""`from file_utils import PYTORCH_PRETRAINED_BERT_CACHE` (run_pretraining.py:45): the constant only (the reference module's S3 / HTTP
download helpers need boto3 and network access and are outside the hot path)."""
import os
from pathlib import Path

PYTORCH_PRETRAINED_BERT_CACHE = Path(os.getenv("PYTORCH_PRETRAINED_BERT_CACHE", Path.home() / ".pytorch_pretrained_bert"))
CONFIG_NAME = "config.json"
WEIGHTS_NAME = "pytorch_model.bin"
