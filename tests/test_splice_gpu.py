"""GPU parity tests for the region-token splice kernels (bit/value-exact data movement)."""
import numpy as np
import pytest
import torch

from gpt4roi_b200.splice import splice_region_tokens
from oracle import roi_align_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _mk(rng, B, L, P, ks, tok, V):
    ids = rng.integers(3, V - 6, (B, L)).astype(np.int64)
    for b in range(B):
        ids[b, 1] = tok['start']
        ids[b, 2:2 + P] = tok['patch']
        ids[b, 2 + P] = tok['end']
        pos = rng.choice(np.arange(3 + P, L), size=ks[b], replace=False)
        ids[b, pos] = tok['bbox']
    return ids


def _u16(t):
    return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
@pytest.mark.parametrize('cfg', [dict(V=64, D=16, P=9, L=40, ks=[3, 0, 1, 5]),
                                 dict(V=32006, D=4096, P=576, L=706, ks=[8] * 8),
                                 dict(V=32006, D=4096, P=256, L=2048, ks=[16, 0, 3])])
def test_splice_matches_oracle(dtype, cfg):
    rng = np.random.default_rng(cfg['L'])
    V, D, P, L, ks = cfg['V'], cfg['D'], cfg['P'], cfg['L'], cfg['ks']
    B = len(ks)
    tok = dict(patch=V - 5, bbox=V - 4, start=V - 2, end=V - 1)
    ids = _mk(rng, B, L, P, ks, tok, V)
    if B > 1:
        ids[1, :] = rng.integers(3, V - 6, L)  # a text-only sample (spi_llava.py:104-111)
        ks = list(ks); ks[1] = 0
    emb = torch.randn(V, D).to(dtype)
    img = torch.randn(B, P, D).to(dtype)
    regs = [torch.randn(k, D).to(dtype) for k in ks]
    offs = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
    packed = torch.cat(regs, 0) if sum(ks) else torch.zeros(1, D, dtype=dtype)
    want = O.splice(ids, _u16(emb), _u16(img), _u16(packed), offs, P, tok['patch'], tok['start'],
                    tok['end'], tok['bbox'])
    got = splice_region_tokens(torch.from_numpy(ids).to(DEV), emb.to(DEV), img.to(DEV),
                               [r.to(DEV) for r in regs], P, tok['patch'], tok['start'], tok['end'],
                               tok['bbox'])
    assert np.array_equal(_u16(got), want)  # moved verbatim: bitwise equal


def test_splice_no_boxes_and_errors():
    rng = np.random.default_rng(2)
    V, D, P, L = 64, 8, 4, 20
    tok = dict(patch=V - 5, bbox=V - 4, start=V - 2, end=V - 1)
    args = (P, tok['patch'], tok['start'], tok['end'], tok['bbox'])
    emb = torch.randn(V, D, device=DEV).bfloat16()
    img = torch.randn(1, P, D, device=DEV).bfloat16()
    ids = _mk(rng, 1, L, P, [0], tok, V)
    out = splice_region_tokens(torch.from_numpy(ids).to(DEV), emb, img, None, *args)  # bboxes=None
    want = O.splice(ids, _u16(emb), _u16(img), None, None, *args)
    assert np.array_equal(_u16(out), want)
    ids2 = _mk(rng, 1, L, P, [2], tok, V)
    with pytest.raises(ValueError, match='bbox'):
        splice_region_tokens(torch.from_numpy(ids2).to(DEV), emb, img, None, *args)
    bad = ids.copy(); bad[0, 2 + P] = 5
    with pytest.raises(ValueError, match='same'):
        splice_region_tokens(torch.from_numpy(bad).to(DEV), emb, img, None, *args)
    bad[0, L - 1] = tok['end']
    with pytest.raises(ValueError, match='follow'):
        splice_region_tokens(torch.from_numpy(bad).to(DEV), emb, img, None, *args)
