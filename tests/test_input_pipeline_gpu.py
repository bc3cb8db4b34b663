"""GPU parity of the fused input-pipeline kernel (g4r_preprocess_images through gpt4roi_b200.input_pipeline) against the
oracle (integer-exact cv2 restatement + mmcv normalisation, pinned to the reference's own transforms): fp32 output
BIT-EXACT, bf16 output = the fp32 result rounded once."""
import os

import numpy as np
import pytest
import torch

from gpt4roi_b200.input_pipeline import BatchPreprocessor
from oracle import input_oracle as O
from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_reference_golden_cases_bit_exact():
    z = np.load(os.path.join(GOLDEN, 'input_pipeline_ref.npz'))
    by_size = {}
    for n in z['names']:
        by_size.setdefault(int(z[str(n) + '.params'][0]), []).append(str(n))
    for S, names in by_size.items():
        pre = BatchPreprocessor(S, DEV, out_dtype=torch.float32)
        imgs = [z[n + '.src'] for n in names]
        boxes = [z[n + '.src_boxes'] for n in names]
        shifts = [tuple(int(v) for v in z[n + '.params'][1:3]) for n in names]
        flips = [bool(z[n + '.params'][3]) for n in names]
        out, nb, idx = pre(imgs, boxes, shifts, flips)          # one launch for the whole (ragged) batch
        for i, n in enumerate(names):
            assert np.array_equal(out[i].cpu().numpy(), z[n + '.image']), n
            assert np.array_equal(nb[i].numpy(), z[n + '.boxes']), n


def test_random_sizes_vs_oracle_fp32_and_bf16():
    rng = np.random.default_rng(3)
    sizes = [(480, 640), (427, 640), (97, 131), (33, 500), (600, 45), (336, 336), (1, 1), (2, 3)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    shifts = [(0, 0), (5, -9), (-32, 31), (17, 0), (0, -3), (31, 31), (0, 0), (-1, 1)]
    flips = [False, True, True, False, True, False, True, False]
    for S in (224, 336):
        want = np.stack([O.preprocess_image(im, S, shift=s, flip=f) for im, s, f in zip(imgs, shifts, flips)])
        got32, _, _ = BatchPreprocessor(S, DEV, out_dtype=torch.float32)(imgs, None, shifts, flips)
        assert np.array_equal(got32.cpu().numpy(), want), S
        got16, _, _ = BatchPreprocessor(S, DEV, out_dtype=torch.bfloat16)(imgs, None, shifts, flips)
        assert torch.equal(got16.cpu(), torch.from_numpy(want).to(torch.bfloat16)), S


def test_feeds_the_engine_dtype_and_layout():
    """The bf16 [B,3,S,S] tensor is what PrefillEngine.forward takes as `images` (no further host work)."""
    pre = BatchPreprocessor(224, DEV)
    rng = np.random.default_rng(1)
    out, nb, _ = pre([rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)], [np.array([[10, 20, 200, 150]], np.float32)])
    assert out.shape == (1, 3, 224, 224) and out.dtype == torch.bfloat16 and out.is_contiguous()
    assert nb[0].shape == (1, 4) and float(nb[0].max()) <= 1.0
