"""Run one region-token prefill and a few eager decode steps (7B, 336 px) -- meant to be wrapped by ncu:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none \
      -k regex:"skinny|decode_attention|kv_append|norm_rows" --csv --log-file out.csv python tools/profile_decode.py 8
usage: python tools/profile_decode.py [batch] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synthetic_inputs  # noqa: E402
from gpt4roi_b200.engine import EngineConfig, KVCache, PrefillEngine, random_state_dicts  # noqa: E402


def main():
    dev = 'cuda:0'
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfg = EngineConfig(image_size=336)
    sd, vit_sd = random_state_dicts(cfg, dev, seed=0)
    eng = PrefillEngine(cfg, sd, vit_sd, dev)
    del sd, vit_sd
    ids, images, boxes = synthetic_inputs(cfg, B, 8, 128)
    ids, images = ids.to(dev), images.to(dev)
    cache = KVCache(cfg, B, ids.shape[1] + steps + 2, dev)
    logits = eng.forward_device(ids, images, eng.plan_boxes(boxes), last_only=True, cache=cache)
    nxt = logits[:, -1].float().argmax(-1, keepdim=True)
    for _ in range(steps):
        logits = eng.decode_step(nxt, cache)
        nxt = logits[:, -1].float().argmax(-1, keepdim=True)
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
