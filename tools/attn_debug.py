"""Bring-up diagnostics for the tcgen05 attention kernel (structured inputs, compact reports)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpt4roi_b200 import kernels
dev = 'cuda:0'
torch.manual_seed(0)


def ref(qkv, B, L, H, D, causal):
    q, k, v = (t.permute(0, 2, 1, 3).float() for t in qkv.view(B, L, 3, H, D).unbind(2))
    s = (q @ k.transpose(-1, -2)) * D ** -0.5
    if causal:
        s = s.masked_fill(torch.ones(L, L, device=dev, dtype=torch.bool).triu(1), float('-inf'))
    return (s.softmax(-1) @ v).permute(0, 2, 1, 3)


def report(name, got, want):
    err = (got.float() - want).abs()
    bad = err > 2e-2
    print('%-40s max_err %.4g mean_err %.3g bad %d/%d nan %d' % (name, err.max().item(), err.mean().item(), int(bad.sum()), bad.numel(), int(torch.isnan(got.float()).sum())))
    if bad.any():
        idx = bad.nonzero()
        print('   first bad (b,l,h,d):', idx[0].tolist(), 'last:', idx[-1].tolist(),
              '| bad rows(l):', sorted(set(idx[:, 1].tolist()))[:12], '| bad d:', sorted(set(idx[:, 3].tolist()))[:12])
        b, l, h, d = idx[0].tolist()
        print('   got ', [round(x, 3) for x in got[b, l, h, d:d + 6].float().tolist()])
        print('   want', [round(x, 3) for x in want[b, l, h, d:d + 6].tolist()])
    sys.stdout.flush()


def run(L, H, D, causal, kind):
    B = 1
    if kind == 'rand':
        qkv = (torch.randn(B * L, 3 * H * D, device=dev) * 0.7)
    elif kind == 'uniform':      # q=k=0 -> uniform softmax: tests the P.V path and V layout only
        qkv = torch.randn(B * L, 3 * H * D, device=dev)
        qkv.view(B, L, 3, H, D)[:, :, :2] = 0
    elif kind == 'v_colid':      # V[key, d] = d: O[., d] must equal d
        qkv = torch.zeros(B * L, 3 * H * D, device=dev)
        qkv.view(B, L, 3, H, D)[:, :, 2] = torch.arange(D, device=dev).float()[None, None, None, :] / 16
    elif kind == 'v_rowid':      # V[key, d] = key: O[q, :] = mean of visible keys
        qkv = torch.zeros(B * L, 3 * H * D, device=dev)
        qkv.view(B, L, 3, H, D)[:, :, 2] = (torch.arange(L, device=dev).float()[None, :, None, None] / 16)
    qkv = qkv.to(torch.bfloat16)
    want = ref(qkv, B, L, H, D, causal)
    for impl in ('tc',):
        got = kernels.attention(qkv, B, L, H, D, causal, D ** -0.5, impl=impl).view(B, L, H, D)
        torch.cuda.synchronize()
        report('L%d H%d D%d causal%d %s [%s]' % (L, H, D, int(causal), kind, impl), got, want)


for D in (128, 64):
    for kind in ('v_colid', 'v_rowid', 'uniform', 'rand'):
        run(128, 1, D, False, kind)
    run(128, 1, D, True, 'rand')
    run(256, 2, D, False, 'rand')
    run(300, 2, D, True, 'rand')
run(706, 32, 128, True, 'rand')
run(577, 16, 64, False, 'rand')
print('attn_debug done')
