"""N>1 host logic on CPU (gloo, world_size 2): the data-parallel path shards the batch by rank, SUM-allreduces/averages the
gradients (DDP) and applies the identical LAMB update on every rank, so an N-rank step on per-rank micro-batches must equal a
1-rank step on the concatenated batch.  The model here is the CPU oracle wrapped in an nn.Module -- the CUDA
kernels cannot run here; what is under test is the partitioning / reduction / timing-aggregation logic bench.py relies on."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

TINY = dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128, vocab_size=256,
            max_position_embeddings=32, type_vocab_size=2, hidden_act="gelu", initializer_range=0.02)


class OracleModule(torch.nn.Module):
    def __init__(self, cfg, seed):
        super().__init__()
        from oracle import bert_oracle as O
        self.cfg, self.O = cfg, O
        sd = O.init_params(cfg, seed=seed, std=0.1)
        self.keys = list(sd.keys())
        self.params = torch.nn.ParameterList([torch.nn.Parameter(sd[k]) for k in self.keys])

    def forward(self, batch):
        sd = dict(zip(self.keys, self.params))
        return self.O.forward_loss(sd, self.cfg, batch)[0]


def _lamb(model, step):
    from oracle import lamb_oracle as LO
    g = dict(params=[p.detach().numpy() for p in model.params], grads=[p.grad.numpy() for p in model.params],
             exp_avg=[np.zeros(tuple(p.shape), np.float32) for p in model.params],
             exp_avg_sq=[np.zeros(tuple(p.shape), np.float32) for p in model.params], lr=1e-2, betas=(0.9, 0.999), eps=1e-6,
             weight_decay=0.01, step=step, bias_correction=True, grad_averaging=True)
    LO.lamb_step([g])


def _batches(world):
    from deeplearningexamples_b200 import training as T
    return [T.synthetic_batch(2, 32, TINY["vocab_size"], 4, seed=T.rank_seed(42, r)) for r in range(world)]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deeplearningexamples_b200 import training as T
    model = OracleModule(TINY, seed=1)
    ddp = torch.nn.parallel.DistributedDataParallel(model, bucket_cap_mb=1, gradient_as_bucket_view=True)
    batch = _batches(world)[rank]
    loss = ddp(batch)
    loss.backward()
    grads = [p.grad.detach().clone() for p in model.params]
    _lamb(model, 0)
    slow = T.max_over_ranks(10.0 + 5.0 * rank)             # the slowest rank defines the step time
    torch.save(dict(params=[p.detach().clone() for p in model.params], grads=grads, slow=slow, ids=batch["input_ids"]),
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_step_equals_single_rank_on_concatenated_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    # every rank ends with identical parameters, and saw a different shard
    for a, b in zip(outs[0]["params"], outs[1]["params"]):
        assert torch.equal(a, b)
    assert not torch.equal(outs[0]["ids"], outs[1]["ids"])
    assert outs[0]["slow"] == outs[1]["slow"] == 15.0
    # single process, concatenated global batch (equal masked-token counts per rank => mean of means == global mean)
    model = OracleModule(TINY, seed=1)
    bs = _batches(world)
    cat = {k: torch.cat([b[k] for b in bs], 0) for k in bs[0]}
    model(cat).backward()
    ref_grads = [p.grad.detach().clone() for p in model.params]
    _lamb(model, 0)
    for g_ddp, g_ref in zip(outs[0]["grads"], ref_grads):               # DDP average == gradient of the global batch
        torch.testing.assert_close(g_ddp, g_ref, rtol=1e-4, atol=1e-7)
    for a, b, g in zip(outs[0]["params"], model.params, ref_grads):
        if g.abs().min() > 1e-4:      # first LAMB step is sign-like: u = g/(|g|+eps) amplifies float noise where |g| <~ eps
            torch.testing.assert_close(a, b.detach(), rtol=1e-4, atol=1e-5)


def test_throughput_formula_matches_reference_definition():
    from deeplearningexamples_b200 import training as T
    # run_pretraining.py:748: train_batch_size * gpu_count * steps / seconds
    assert T.global_throughput(32, 8, 10, 2000.0) == pytest.approx(32 * 8 * 10 / 2.0)
    assert T.rank_seed(42, 3) == 45
