"""Input-pipeline throughput: 64 decoded COCO-sized uint8 images (480x640) + 8 boxes each ->
[64,3,336,336] bf16 on the GPU (one fused kernel, incl. the pinned H2D copy of the 59 MB of pixels), beside the same
transforms on the host cores (cv2.resize + mmcv-style normalisation, what the reference's dataloader workers run).
usage: python tools/bench_input.py   (prints one JSON line)"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpt4roi_b200.input_pipeline import BatchPreprocessor, draw_augmentation  # noqa: E402


def main():
    B, S = 64, 336
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(B)]
    boxes = [np.sort(rng.uniform(0, 1, (8, 2, 2)), axis=1).transpose(0, 2, 1).reshape(8, 4)[:, [0, 2, 1, 3]] *
             np.array([640, 480, 640, 480]) for _ in range(B)]
    boxes = [np.stack([b[:, 0], b[:, 1], np.maximum(b[:, 2], b[:, 0] + 4), np.maximum(b[:, 3], b[:, 1] + 4)], 1).astype(np.float32)
             for b in boxes]
    shifts, flips = draw_augmentation(B, rng)
    pre = BatchPreprocessor(S, 'cuda:0')
    for _ in range(3):
        pre(imgs, boxes, shifts, flips)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        out, nb, _ = pre(imgs, boxes, shifts, flips)
    torch.cuda.synchronize()
    gpu_s = (time.perf_counter() - t0) / n
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # kernel alone: data resident
    from gpt4roi_b200 import lib as L
    d_src = torch.from_numpy(np.concatenate([a.reshape(-1) for a in imgs])).to('cuda:0')
    offs = torch.arange(B, dtype=torch.int64, device='cuda:0') * imgs[0].size
    hw = torch.tensor([[480, 640]] * B, dtype=torch.int32, device='cuda:0')
    o = torch.empty((B, 3, S, S), dtype=torch.bfloat16, device='cuda:0')
    for _ in range(3):
        L.check(L.load().g4r_preprocess_images(L.ptr(d_src), L.ptr(offs), L.ptr(hw), None, None, L.ptr(o), B, S, pre._mean, pre._std,
                                               1, L.BF16, L.stream_ptr(torch.device('cuda:0'))))
    e0.record()
    for _ in range(20):
        L.check(L.load().g4r_preprocess_images(L.ptr(d_src), L.ptr(offs), L.ptr(hw), None, None, L.ptr(o), B, S, pre._mean, pre._std,
                                               1, L.BF16, L.stream_ptr(torch.device('cuda:0'))))
    e1.record()
    torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / 20
    cpu = None
    try:
        import cv2
        mean32 = np.array([0.48145466 * 255, 0.4578275 * 255, 0.40821073 * 255], np.float32)
        sinv = 1.0 / np.array([0.26862954 * 255, 0.26130258 * 255, 0.27577711 * 255], np.float32).astype(np.float64)
        t0 = time.perf_counter()
        for im in imgs:
            r = cv2.resize(im, (S, S), interpolation=cv2.INTER_LINEAR).astype(np.float32)[..., ::-1]
            ((r - mean32).astype(np.float64) * sinv).astype(np.float32).transpose(2, 0, 1).copy()
        cpu = B / (time.perf_counter() - t0)
    except Exception as e:
        cpu = str(e)
    bytes_alg = B * 480 * 640 * 3 + B * 3 * S * S * 2
    print(json.dumps(dict(metric='input_pipeline_images_per_s', gpu_images_per_s_incl_h2d_and_host_packing=B / gpu_s,
                          kernel_ms=k_ms, kernel_images_per_s=B / (k_ms / 1e3), kernel_GBps=bytes_alg / 1e9 / (k_ms / 1e3),
                          cpu_one_core_images_per_s=cpu, batch=B, src='480x640 uint8 BGR', out='[64,3,336,336] bf16',
                          note='prefill consumes 75 images/s/GPU (600/s at 8 GPUs)')))


if __name__ == '__main__':
    main()
