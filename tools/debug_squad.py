"""Debug: last-layer LayerNorm bias gradient in the SQuAD path (tests/test_squad_gpu.py failure): compare p.grad with sum_t dL/dy."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_squad_gpu as t
SQ, O, sd, full, qa_w, qa_b, batch = t._setup()
model, opt, sched = SQ.prepare_squad_model_and_optimizer(t.CFG, torch.device("cuda", 0), state_dict=full, total_steps=100)
model.train()
bd = {k: v.cuda() for k, v in batch.items()}
enc, _ = model.bert(bd["input_ids"], bd["segment_ids"], bd["input_mask"])
seq = enc[-1]
seq.retain_grad()
logits = model.qa_outputs(seq)
s_log, e_log = (x.squeeze(-1) for x in logits.split(1, dim=-1))
loss = SQ.squad_loss(s_log, e_log, bd["start_positions"], bd["end_positions"])
loss.backward()
named = dict(model.named_parameters())
for k in ("bert.encoder.layer.1.output.LayerNorm.bias", "bert.encoder.layer.1.output.LayerNorm.weight", "bert.encoder.layer.0.output.LayerNorm.bias"):
    print(k, "grad norm", named[k].grad.float().norm().item(), "dtype", named[k].grad.dtype)
print("sum_t dy norm", seq.grad.float().sum((0, 1)).norm().item(), "dy shape", tuple(seq.grad.shape), seq.grad.stride(), seq.grad.is_contiguous())
print("diff", (named["bert.encoder.layer.1.output.LayerNorm.bias"].grad.float() - seq.grad.float().sum((0, 1))).abs().max().item())
