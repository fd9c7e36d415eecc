"""`from utils import is_main_process, format_step, get_world_size, get_rank` (run_pretraining.py:46): rank helpers with the
reference's semantics (utils.py:21-60), needed because the driver is run from a scratch copy without its sibling files."""
from pathlib import Path

import torch.distributed as dist


def get_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def is_main_process():
    return get_rank() == 0


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def format_step(step):
    if isinstance(step, str):
        return step
    names = ("Training Epoch", "Training Iteration", "Validation Iteration")
    return "".join("{}: {} ".format(n, s) for n, s in zip(names, step))


def mkdir(path):
    Path(path).mkdir(parents=True, exist_ok=True)


def mkdir_by_main_process(path):
    if is_main_process():
        mkdir(path)
    barrier()
