"""GPU parity of the fused softmax cross-entropy (csrc/loss.cu) against torch's CrossEntropyLoss(ignore_index=-1) evaluated in fp32 on
the same bf16 logits -- what the reference criterion computes under autocast (run_pretraining.py:85-95)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
bf = torch.bfloat16


@pytest.mark.parametrize("rows,V", [(160, 30528), (37, 1024), (10240, 30528)])
def test_softmax_ce_forward_backward_vs_torch_fp32(rows, V):
    from deeplearningexamples_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + V)
    logits = (torch.randn(rows, V, generator=g, device="cuda") * 3.0).to(bf)
    labels = torch.randint(0, V, (rows,), generator=g, device="cuda")
    labels[torch.rand(rows, generator=g, device="cuda") < 0.25] = -1                 # ignored rows (padding slots of the static gather)
    labels[0] = 5
    x = logits.clone().requires_grad_(True)
    loss = ops.SoftmaxCrossEntropyFn.apply(x, labels, -1)
    loss.backward(torch.tensor(1.7, device="cuda"))
    xr = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(xr, labels, ignore_index=-1)
    ref.backward(torch.tensor(1.7, device="cuda"))
    assert loss.dtype == torch.float32
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item()) + 1e-6
    got, want = x.grad.float(), xr.grad
    assert x.grad.dtype == bf
    # bf16 rounding of the gradient only: relative L2 and elementwise bound of one bf16 ulp of the largest entry per row
    assert ((got - want).norm() / want.norm()).item() < 4e-3
    assert (got - want).abs().max().item() <= 2 ** -8 * want.abs().max().item() + 1e-12
    assert torch.count_nonzero(got[labels == -1]) == 0
    ops.check_device_errors()


def test_softmax_ce_reports_out_of_range_labels():
    from deeplearningexamples_b200 import _lib as L
    from deeplearningexamples_b200 import ops
    logits = torch.zeros(4, 64, device="cuda", dtype=bf)
    labels = torch.tensor([1, 64, -1, 3], device="cuda")
    ops.SoftmaxCrossEntropyFn.apply(logits, labels, -1)
    with pytest.raises(L.DleError):
        ops.check_device_errors()
