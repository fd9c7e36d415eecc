"""GPU parity: LayerNorm family, bias-GELU, embeddings, gathers vs plain torch fp32 references."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
bf = torch.bfloat16


def _k():
    from deeplearningexamples_b200 import kernels
    return kernels


def _rand(shape, scale=1.0, seed=0, dtype=bf):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


@pytest.mark.parametrize("T,H", [(7, 256), (1000, 768), (4096, 1024), (3, 512)])
def test_add_ln_fwd_bwd(T, H):
    k = _k()
    x, res, bias = _rand((T, H), seed=1), _rand((T, H), seed=2), _rand((H,), 0.1, seed=3)
    gamma, beta = (1 + _rand((H,), 0.1, seed=4).float()).to(bf), _rand((H,), 0.1, seed=5)
    y, z, mean, rstd = k.add_ln_fwd(x, gamma, beta, bias=bias, residual=res)
    z_ref = (x.float() + bias.float() + res.float())
    torch.testing.assert_close(z.float(), z_ref, rtol=1e-2, atol=2e-2)
    zf = z.float().requires_grad_(True)
    gf, bfl = gamma.float().requires_grad_(True), beta.float().requires_grad_(True)
    y_ref = F.layer_norm(zf, (H,), gf, bfl, eps=1e-12)
    torch.testing.assert_close(y.float(), y_ref, rtol=1e-2, atol=2e-2)
    torch.testing.assert_close(mean, z.float().mean(-1), rtol=1e-4, atol=1e-4)
    dy = _rand((T, H), seed=6)
    y_ref.backward(dy.float())
    dz, dx, dgamma, dbeta, dbias = k.add_ln_bwd(dy, z, mean, rstd, gamma)
    assert dx is dz
    torch.testing.assert_close(dz.float(), zf.grad, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(dgamma, gf.grad, rtol=1e-2, atol=1e-2 * gf.grad.abs().max().item())
    torch.testing.assert_close(dbeta, bfl.grad, rtol=1e-2, atol=1e-2 * bfl.grad.abs().max().item())
    torch.testing.assert_close(dbias, dz.float().sum(0), rtol=1e-3, atol=1e-2)


def test_plain_ln_no_z():
    k = _k()
    x = _rand((300, 1024), seed=7)
    gamma, beta = _rand((1024,), seed=8), _rand((1024,), seed=9)
    y, z, mean, rstd = k.add_ln_fwd(x, gamma, beta)
    assert z is x
    torch.testing.assert_close(y.float(), F.layer_norm(x.float(), (1024,), gamma.float(), beta.float(), eps=1e-12), rtol=1e-2, atol=2e-2)


def test_add_ln_dropout_consistency():
    """fwd and bwd regenerate the same mask: dx == dz * mask / (1-p), mask inferred from forward."""
    k = _k()
    T, H, p = 512, 1024, 0.1
    x = (_rand((T, H), seed=10).float().abs() + 1.0).to(bf)   # never 0, so z != 0 <=> kept
    gamma, beta = torch.ones(H, device="cuda", dtype=bf), torch.zeros(H, device="cuda", dtype=bf)
    y, z, mean, rstd = k.add_ln_fwd(x, gamma, beta, dropout_p=p, seed=99, dropout_stream=5)
    mask = z.float() != 0
    assert 0.88 < mask.float().mean().item() < 0.92
    torch.testing.assert_close(z.float()[mask], (x.float() / (1 - p))[mask], rtol=1e-2, atol=1e-2)
    dy = _rand((T, H), seed=11)
    dz, dx, *_ = k.add_ln_bwd(dy, z, mean, rstd, gamma, dropout_p=p, seed=99, dropout_stream=5)
    torch.testing.assert_close(dx.float(), dz.float() * mask / (1 - p), rtol=1e-2, atol=1e-3)
    y2, z2, *_ = k.add_ln_fwd(x, gamma, beta, dropout_p=p, seed=99, dropout_stream=6)
    assert not torch.equal(z, z2)          # different stream id => different mask


def test_colsum():
    k = _k()
    for T, N in [(5, 256), (4096, 3072), (1000, 4096), (33, 1032)]:
        x = _rand((T, N), seed=12)
        torch.testing.assert_close(k.colsum(x), x.float().sum(0), rtol=1e-4, atol=1e-2)


def test_bias_gelu_standalone():
    k = _k()
    x, bias = _rand((777, 4096), 2.0, seed=13), _rand((4096,), seed=14)
    y, u = k.bias_gelu_fwd(x, bias)
    torch.testing.assert_close(u.float(), x.float() + bias.float(), rtol=1e-2, atol=2e-2)
    torch.testing.assert_close(y.float(), F.gelu(u.float(), approximate="tanh"), rtol=1e-2, atol=1e-2)
    dy = _rand((777, 4096), seed=15)
    uf = u.float().requires_grad_(True)
    F.gelu(uf, approximate="tanh").backward(dy.float())
    torch.testing.assert_close(k.bias_gelu_bwd(dy, u).float(), uf.grad, rtol=2e-2, atol=1e-2)
    # reference repo's only numeric KAT (TF2 gelu_test.py:29-32)
    kat = torch.tensor([[0.25, 0.0, -0.25, -1.0, -2.0, 3.0, 0.0, 0.0]], device="cuda", dtype=bf)
    want = torch.tensor([0.14967535, 0.0, -0.10032465, -0.15880796, -0.04540223, 2.9963627, 0.0, 0.0], device="cuda")
    torch.testing.assert_close(k.bias_gelu_fwd(kat)[0].float()[0], want, rtol=1e-2, atol=2e-3)


def test_embedding_gather_bit_exact_and_ln():
    k = _k()
    B, S, H, V = 4, 128, 1024, 30528
    word, pos, typ = _rand((V, H), 0.02, seed=16), _rand((512, H), 0.02, seed=17), _rand((2, H), 0.02, seed=18)
    gamma, beta = (1 + _rand((H,), 0.1, seed=19).float()).to(bf), _rand((H,), 0.1, seed=20)
    g = torch.Generator(device="cuda").manual_seed(21)
    ids = torch.randint(0, V, (B, S), generator=g, device="cuda")
    tt = torch.randint(0, 2, (B, S), generator=g, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    y, z, mean, rstd = k.embed_ln_fwd(ids, tt, word, pos, typ, gamma, beta, err_flag=err)
    # the pre-LN sum is the bf16 rounding of the exact fp32 sum of three bit-exactly gathered rows
    z_ref = (word[ids].float() + pos[torch.arange(S, device="cuda")].float().unsqueeze(0) + typ[tt].float()).to(bf).view(B * S, H)
    assert torch.equal(z, z_ref)
    assert err.item() == 0
    y_ref = F.layer_norm(z.float(), (H,), gamma.float(), beta.float(), eps=1e-12)
    torch.testing.assert_close(y.float(), y_ref, rtol=1e-2, atol=2e-2)
    # pure gather bit-exactness: zero the other two tables
    y0, z0, *_ = k.embed_ln_fwd(ids, tt, word, torch.zeros_like(pos), torch.zeros_like(typ), gamma, beta)
    assert torch.equal(z0, word[ids].view(B * S, H))
    # backward
    dy = _rand((B * S, H), seed=22)
    wf, pf, tf_ = word.float().requires_grad_(True), pos.float().requires_grad_(True), typ.float().requires_grad_(True)
    gf, bfl = gamma.float().requires_grad_(True), beta.float().requires_grad_(True)
    e = wf[ids] + pf[torch.arange(S, device="cuda")].unsqueeze(0) + tf_[tt]
    F.layer_norm(e, (H,), gf, bfl, eps=1e-12).view(B * S, H).backward(dy.float())
    dword, dpos, dtyp, dgamma, dbeta = k.embed_ln_bwd(dy, z, mean, rstd, gamma, ids, tt, V, 512, 2)
    torch.testing.assert_close(dword, wf.grad, rtol=2e-2, atol=2e-2 * wf.grad.abs().max().item())
    torch.testing.assert_close(dpos, pf.grad, rtol=2e-2, atol=2e-2 * pf.grad.abs().max().item())
    torch.testing.assert_close(dtyp, tf_.grad, rtol=2e-2, atol=2e-2 * tf_.grad.abs().max().item())
    torch.testing.assert_close(dgamma, gf.grad, rtol=2e-2, atol=2e-2 * gf.grad.abs().max().item())
    torch.testing.assert_close(dbeta, bfl.grad, rtol=2e-2, atol=2e-2 * bfl.grad.abs().max().item())


def test_embedding_out_of_range_sets_flag():
    k = _k()
    H, V = 256, 100
    word, pos, typ = _rand((V, H), seed=23), _rand((16, H), seed=24), _rand((2, H), seed=25)
    gamma, beta = _rand((H,), seed=26), _rand((H,), seed=27)
    ids = torch.tensor([[1, 2, 100, 4]], device="cuda")
    tt = torch.zeros_like(ids)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    k.embed_ln_fwd(ids, tt, word, pos, typ, gamma, beta, err_flag=err)
    assert err.item() == 1


def test_gather_scatter_rows_bit_exact():
    k = _k()
    x = _rand((16384, 1024), seed=28)
    g = torch.Generator(device="cuda").manual_seed(29)
    idx = torch.randperm(16384, generator=g, device="cuda")[:2560].sort().values
    out = k.gather_rows(x, idx)
    assert torch.equal(out, x.index_select(0, idx))
    dx = k.scatter_rows(out, idx, 16384)
    ref = torch.zeros_like(x); ref[idx] = out
    assert torch.equal(dx, ref)
    assert k.gather_rows(x, idx[:0]).shape == (0, 1024)


def test_casts_round_trip():
    k = _k()
    x = _rand((1000003,), seed=30, dtype=torch.float32)
    y = k.cast_f32_to_bf16(x)
    assert torch.equal(y, x.to(bf))
    assert torch.equal(k.cast_bf16_to_f32(y), y.float())
