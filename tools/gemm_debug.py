"""Bring-up diagnostics for the tcgen05 GEMM: tiny structured problems, compact error reports."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpt4roi_b200 import dense

dev = 'cuda:0'
torch.manual_seed(0)


def report(name, got, want):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    bad = err > (0.02 * want.abs().max() + 1e-3)
    print('%-34s max_err %.4g  bad %d/%d  |want|max %.3g  nan %d' % (
        name, err.max().item(), int(bad.sum()), bad.numel(), want.abs().max().item(), int(torch.isnan(got).sum())))
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print('   bad rows: n=%d first %s last %s | bad cols: n=%d first %s last %s' % (
            len(rows), rows[:8].tolist(), rows[-4:].tolist(), len(cols), cols[:8].tolist(), cols[-4:].tolist()))
        r, c = int(rows[0]), int(cols[0])
        print('   got[%d,%d:%d+6]  %s' % (r, c, c, [round(x, 3) for x in got[r, c:c + 6].tolist()]))
        print('   want[%d,%d:%d+6] %s' % (r, c, c, [round(x, 3) for x in want[r, c:c + 6].tolist()]))
    sys.stdout.flush()


def run(M, N, K, kind='rand'):
    if kind == 'rand':
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        b = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    elif kind == 'eyeA':   # D[i,:] = B[:, i]
        a = torch.zeros(M, K, device=dev).bfloat16()
        for i in range(min(M, K)):
            a[i, i] = 1
        b = (torch.randn(N, K, device=dev)).bfloat16()
    elif kind == 'ones':
        a = torch.ones(M, K, device=dev).bfloat16()
        b = torch.ones(N, K, device=dev).bfloat16()
    elif kind == 'rowid':  # A[i,:]=1/K * (i+1) ; B ones -> D[i,j] = i+1
        a = ((torch.arange(M, device=dev).float() + 1)[:, None].expand(M, K) / K).bfloat16()
        b = torch.ones(N, K, device=dev).bfloat16()
    elif kind == 'colid':
        a = torch.ones(M, K, device=dev).bfloat16()
        b = ((torch.arange(N, device=dev).float() % 64 + 1)[:, None].expand(N, K) / K).bfloat16()
    want = a.float() @ b.float().t()
    got = dense.linear(a, b, out_dtype=torch.float32)
    torch.cuda.synchronize()
    report('M%d N%d K%d %s' % (M, N, K, kind), got, want)


for kind in ('ones', 'rowid', 'colid', 'eyeA', 'rand'):
    run(128, 128, 16, kind)
run(128, 128, 64, 'rand')
run(128, 128, 128, 'rand')
run(128, 128, 1024, 'rand')
run(128, 256, 64, 'rand')      # BLOCK_N=128 still (few tiles)
run(256, 256, 256, 'rand')
run(100, 72, 136, 'rand')
run(4096, 4096, 512, 'rand')   # BLOCK_N=256, multi-tile persistent
run(5648, 4096, 4096, 'rand')
print('gemm_debug done')
