"""Extra smoke step for __graft_entry__.smoke(): one small end-to-end prefill through every kernel
family (tcgen05 GEMM / implicit-GEMM conv / tcgen05 attention / RoIAlign / splice / elementwise),
checked against the oracle composition when it is importable (the oracle is test infrastructure:
this module only calls it from smoke(), never from the product path)."""
import torch

from .engine import EngineConfig, PrefillEngine, random_state_dicts


def run(device='cuda:0'):
    cfg = EngineConfig(image_size=224, vit_layers=12, n_layers=1)
    sd, vit_sd = random_state_dicts(cfg, device, seed=2)
    eng = PrefillEngine(cfg, sd, vit_sd, device)
    g = torch.Generator().manual_seed(1)
    P = cfg.num_patches
    L = P + 2 + 12
    ids = torch.randint(3, 32000, (1, L), generator=g)
    ids[0, 0] = 1
    ids[0, 1] = cfg.im_start_token
    ids[0, 2:2 + P] = cfg.im_patch_token
    ids[0, 2 + P] = cfg.im_end_token
    ids[0, [P + 5, P + 9]] = cfg.bbox_token
    images = torch.randn(1, 3, 224, 224, generator=g).to(torch.bfloat16)
    boxes = [torch.tensor([[0.1, 0.2, 0.6, 0.7], [0.3, 0.05, 0.95, 0.5]])]
    logits = eng.forward(ids.to(device), images.to(device), boxes).float()
    assert logits.shape == (1, L, cfg.vocab) and torch.isfinite(logits).all(), 'engine produced non-finite logits'
    try:
        from oracle import model_oracle
    except Exception as e:  # pragma: no cover
        print('[smoke] oracle not importable (%s); engine ran, parity not checked' % e)
        return
    ref = model_oracle.forward(cfg, sd, vit_sd, ids, images.float(), boxes, device)
    err = ((logits - ref).norm() / ref.norm()).item()
    assert err < 3e-2, 'engine logits differ from the oracle: rel-L2 %.3e' % err
    print('[smoke] engine vs oracle logits rel-L2 %.3e' % err)
