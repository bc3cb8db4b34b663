"""GPU numerics of the training-step kernels against plain PyTorch fp32 autograd of
the same op on the same bf16 inputs.  Tolerances (bf16 storage, fp32 accumulation) are stated per test."""
import math

import pytest
import torch
import torch.nn.functional as F

from gpt4roi_b200 import train_ops

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


@pytest.mark.parametrize('M,V', [(1, 8), (37, 1000), (706, 32006), (64, 32008)])
def test_cross_entropy_matches_torch(M, V):
    """loss = CrossEntropyLoss()(logits.float(), targets) with ignore_index -100; dlogits = autograd.
    Tolerance: loss 1e-5 relative (fp32 both sides), dlogits rel-L2 <= 4e-3 (bf16 output)."""
    torch.manual_seed(M + V)
    logits = (torch.randn(M, V, device=DEV) * 3).to(BF)
    tg = torch.randint(0, V, (M,), device=DEV)
    tg[torch.rand(M, device=DEV) < 0.3] = -100
    if M > 1:
        tg[0] = 5
    else:
        tg[0] = 3
    loss, cnt, dl = train_ops.cross_entropy(logits, tg, grad_scale=1.0)
    lf = logits.float().requires_grad_()
    want = F.cross_entropy(lf, tg, ignore_index=-100)
    want.backward()
    assert int(cnt.item()) == int((tg >= 0).sum().item())
    assert abs(loss.item() - want.item()) <= 1e-5 * max(1.0, abs(want.item()))
    assert rel(dl, lf.grad) < 4e-3
    assert torch.all(dl[tg < 0] == 0)
    loss2, _, dl2 = train_ops.cross_entropy(logits, tg, grad_scale=1.0)
    assert torch.equal(dl, dl2) and loss.item() == loss2.item()      # reproducible
    # all rows ignored -> loss 0, zero gradient (no NaN)
    l0, c0, d0 = train_ops.cross_entropy(logits, torch.full_like(tg, -100))
    assert l0.item() == 0.0 and c0.item() == 0.0 and torch.all(d0 == 0)


@pytest.mark.parametrize('M,D', [(5, 256), (706, 4096), (1000, 1024)])
def test_rmsnorm_bwd_matches_autograd(M, D):
    """LlamaRMSNorm backward; dx rel-L2 <= 6e-3, dw rel-L2 <= 6e-3 (fp32 reference without the bf16 rounding
    of x_hat that the forward applies)."""
    torch.manual_seed(D)
    x = (torch.randn(M, D, device=DEV) * 2).to(BF)
    w = (torch.rand(D, device=DEV) + 0.5).to(BF)
    dy = torch.randn(M, D, device=DEV).to(BF)
    xf, wf = x.float().requires_grad_(), w.float().requires_grad_()
    y = wf * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6))
    (y * dy.float()).sum().backward()
    dx, dw = train_ops.rmsnorm_bwd(x, w, dy, 1e-6)
    assert rel(dx, xf.grad) < 6e-3
    assert rel(dw, wf.grad) < 6e-3
    dx2, dw2 = train_ops.rmsnorm_bwd(x, w, dy, 1e-6)
    assert torch.equal(dx, dx2) and torch.equal(dw, dw2)
    dres = torch.randn(M, D, device=DEV).to(BF)
    dx3, _ = train_ops.rmsnorm_bwd(x, w, dy, 1e-6, dres=dres)
    assert rel(dx3, xf.grad + dres.float()) < 6e-3


def test_swiglu_fwd_bwd_match_autograd():
    torch.manual_seed(0)
    M, Fd = 300, 11008
    gu = torch.randn(M, 2 * Fd, device=DEV).to(BF)
    df = torch.randn(M, Fd, device=DEV).to(BF)
    guf = gu.float().requires_grad_()
    f = F.silu(guf[:, 0::2]) * guf[:, 1::2]
    (f * df.float()).sum().backward()
    assert rel(train_ops.swiglu_fwd(gu), f) < 4e-3
    assert rel(train_ops.swiglu_bwd(gu, df), guf.grad) < 4e-3


@pytest.mark.parametrize('g_dtype', [BF, torch.float32])
def test_adamw_matches_torch(g_dtype):
    """Three steps of torch.optim.AdamW on fp32 parameters; identical up to fp32 rounding (rtol 1e-5)."""
    torch.manual_seed(1)
    n = 100003
    p0 = torch.randn(n, device=DEV)
    ref = p0.clone().requires_grad_()
    opt = torch.optim.AdamW([ref], lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    p16 = torch.empty(n, device=DEV, dtype=BF)
    for step in range(1, 4):
        g = torch.randn(n, device=DEV).to(g_dtype)
        ref.grad = g.float() * 0.5
        opt.step()
        train_ops.adamw_step(p, g, m, v, p16, 2e-3, (0.9, 0.999), 1e-8, 0.1, step, grad_scale=0.5)
        torch.testing.assert_close(p, ref.detach(), rtol=1e-5, atol=1e-6)
        assert torch.equal(p16, p.to(BF))


def _ref_attention(qkv, B, L, H, D, causal, scale):
    q, k, v = [t.reshape(B, L, H, D).permute(0, 2, 1, 3) for t in qkv.float().split(H * D, dim=1)]
    s = q @ k.transpose(-1, -2) * scale
    if causal:
        s = s.masked_fill(torch.ones(L, L, device=qkv.device, dtype=torch.bool).triu(1), float('-inf'))
    p = torch.softmax(s, -1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * L, H * D), torch.logsumexp(s, -1)


@pytest.mark.parametrize('B,L,H,D,causal', [(1, 64, 2, 128, True), (2, 200, 3, 128, True), (1, 706, 4, 128, True),
                                             (2, 130, 2, 64, False), (1, 77, 2, 64, True)])
def test_attention_fwd_lse_and_bwd_match_autograd(B, L, H, D, causal):
    """Flash-style backward vs autograd of softmax(QK^T*scale [+causal]) V in fp32 on the same bf16 q,k,v,dO.
    Tolerance: out / lse as the forward test; dQ, dK, dV rel-L2 <= 1.5e-2 (P and dS rounded to bf16)."""
    torch.manual_seed(L + D)
    hid = H * D
    qkv = torch.randn(B * L, 3 * hid, device=DEV).to(BF)
    dout = torch.randn(B * L, hid, device=DEV).to(BF)
    scale = D ** -0.5
    out, lse = train_ops.attention_fwd_lse(qkv, B, L, H, D, causal, scale)
    qf = qkv.float().requires_grad_()
    want, want_lse = _ref_attention(qf, B, L, H, D, causal, scale)
    (want * dout.float()).sum().backward()
    assert rel(out, want) < 8e-3
    torch.testing.assert_close(lse, want_lse, rtol=1e-4, atol=2e-4)
    dqkv = train_ops.attention_bwd(qkv, out, dout, lse, B, L, H, D, causal, scale)
    for name, sl in (('dq', slice(0, hid)), ('dk', slice(hid, 2 * hid)), ('dv', slice(2 * hid, 3 * hid))):
        e = rel(dqkv[:, sl], qf.grad[:, sl])
        assert e < 1.5e-2, (name, e)
    assert torch.equal(dqkv, train_ops.attention_bwd(qkv, out, dout, lse, B, L, H, D, causal, scale))  # no atomics


@pytest.mark.parametrize('n,H,Cin,Cout', [(2, 12, 256, 128), (1, 24, 1024, 1024), (3, 14, 256, 256)])
def test_conv3x3_backward_matches_autograd(n, H, Cin, Cout):
    """grad_x (implicit-GEMM conv on flipped/transposed weights) and grad_W (one padded-layout GEMM with the nine
    taps as column blocks) vs autograd of F.conv2d in fp32.  rel-L2 <= 6e-3 (bf16 operands, fp32 accumulation)."""
    torch.manual_seed(H + Cin)
    x = (torch.randn(n, H, H, Cin, device=DEV) * 0.5).to(BF)
    w = (torch.randn(Cout, 3, 3, Cin, device=DEV) * 0.05).to(BF)
    dz = (torch.randn(n, H, H, Cout, device=DEV) * 0.1).to(BF)
    xf = x.float().permute(0, 3, 1, 2).requires_grad_()
    wf = w.float().permute(0, 3, 1, 2).requires_grad_()
    (F.conv2d(xf, wf, padding=1) * dz.float().permute(0, 3, 1, 2)).sum().backward()
    dx, dW = train_ops.conv3x3_bwd(x, w, dz)
    assert rel(dx, xf.grad.permute(0, 2, 3, 1)) < 6e-3
    assert rel(dW, wf.grad.permute(0, 2, 3, 1)) < 6e-3
    _, dW2 = train_ops.conv3x3_bwd(x, w, dz, dw_acc=dW.clone(), need_dx=False)     # accumulation: 2x
    assert rel(dW2, 2 * wf.grad.permute(0, 2, 3, 1)) < 6e-3
    # a level slice of a stacked [Cout, L, 3, 3, Cin] weight (pconv layout) is flipped in place
    stacked = torch.stack([w, w * 2], 1).contiguous()
    assert torch.equal(train_ops.conv_weight_flip_t(stacked[:, 0]), train_ops.conv_weight_flip_t(w))


def test_gn_relu_bwd_matches_autograd():
    """GroupNorm(64)+ReLU backward on a conv output with the forward's own statistics vs autograd of
    relu(group_norm(z.float())).  dz rel-L2 <= 6e-3, dgamma/dbeta <= 2e-3; accumulation doubles the param grads."""
    from gpt4roi_b200 import dense, kernels
    torch.manual_seed(0)
    B, H, C, G = 2, 24, 1024, 64
    x = (torch.randn(B, H, H, C, device=DEV) * 0.5).to(BF)
    w = (torch.randn(C, 3, 3, C, device=DEV) * 0.02).to(BF)
    gamma = (torch.rand(C, device=DEV) + 0.5).to(BF)
    beta = (torch.randn(C, device=DEV) * 0.2).to(BF)
    st = torch.zeros((B, dense.gn_slots(H, H), G, 2), dtype=torch.float32, device=DEV)
    z = dense.conv_nhwc(x, w, gn_stats=st)
    count = H * H * (C // G)
    scale, shift = kernels.gn_finalize(st, gamma, beta, count=count)
    for dt in (torch.float32, BF):
        dA = torch.randn(B, H, H, C, device=DEV).to(dt)
        zf = z.float().permute(0, 3, 1, 2).requires_grad_()
        gf, bf = gamma.float().requires_grad_(), beta.float().requires_grad_()
        a = F.relu(F.group_norm(zf, G, gf, bf, 1e-5))
        (a * dA.float().permute(0, 3, 1, 2)).sum().backward()
        dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        dz = train_ops.gn_relu_bwd(z, dA, scale, shift, st, count, gamma, dg, db, accumulate=False)
        assert rel(dz, zf.grad.permute(0, 2, 3, 1)) < 6e-3
        assert rel(dg, gf.grad) < 2e-3 and rel(db, bf.grad) < 2e-3
        train_ops.gn_relu_bwd(z, dA, scale, shift, st, count, gamma, dg, db, accumulate=True)
        assert rel(dg, 2 * gf.grad) < 2e-3 and rel(db, 2 * bf.grad) < 2e-3


def test_fuse_gather_bwd_is_the_adjoint_of_fuse_gather():
    """<fuse_gather(a), d_in> == <a, fuse_gather_bwd(d_in)> checked through autograd of the torch restatement
    (own | resize(top[3C/4:]) | resize(down[C/2:3C/4]), align_corners bilinear).  rel-L2 <= 4e-3 per level."""
    torch.manual_seed(1)
    B, C, n = 2, 256, 4
    sizes = [32, 16, 8, 4]
    q = C // 4
    a = [torch.randn(B, C, h, h, device=DEV, requires_grad=True) for h in sizes]
    d_in = [(torch.randn(B, h, h, C, device=DEV) * 0.3).to(BF) for h in sizes]
    loss = 0
    for l in range(n):
        top, down = min(l + 1, n - 1), max(l - 1, 0)
        ft = F.interpolate(a[top][:, 3 * q:], size=(sizes[l],) * 2, mode='bilinear', align_corners=True)
        fd = F.interpolate(a[down][:, 2 * q:3 * q], size=(sizes[l],) * 2, mode='bilinear', align_corners=True)
        zin = torch.cat([a[l][:, :2 * q], ft, fd], 1)
        loss = loss + (zin * d_in[l].float().permute(0, 3, 1, 2)).sum()
    loss.backward()
    for m in range(n):
        got = train_ops.fuse_gather_bwd(d_in, m)
        assert rel(got, a[m].grad.permute(0, 2, 3, 1)) < 4e-3, m
