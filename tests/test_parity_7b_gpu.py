"""Parity at the BASELINE configuration: the FULL 32-layer LLaMA-7B +
CLIP-ViT-L/14 (24 layers) + SPI module at 336 px, B=2, 8 RoIs per image, 128 text tokens (L = 706), driven through
the model seam (`gpt4roi_b200.spi_llava.SPILlavaMPTForCausalLM.forward`, weights loaded by reference name), against

  (a) the fp32 oracle  -- oracle/model_oracle.py: transformers CLIP/LLaMA (eager attention) + oracle/spi_oracle.py
      (pinned to the reference's own module outputs by tests/golden/*), the accuracy anchor, and
  (b) the same oracle under bf16 autocast with bf16 weights -- the reference's operating mode (train_stage2.sh
      --bf16 True) and the same-dtype figure SURVEY 8(c) asks for.

Error is recorded per stage and with depth (ViT taps, region tokens, spliced embeddings, residual stream after decoder
layers 8 / 16 / 24 / 31, logits) as rel-L2 = ||a-b|| / ||b|| and max-rel = max|a-b| / max|b|.

Two engine modes are measured: the default bf16 LLaMA residual stream (what a model cast to bf16 has; the reference's
bf16-autocast run below is this mode) and EngineConfig.llama_stream='fp32' (what the reference has in training -- fp32
parameters under autocast; more accurate, 2.4 % slower).

Stated tolerance (what bf16 GEMM operands can hold at 32 layers; north_star's "1e-3" is an fp32-vs-fp32 figure that the
reference's own bf16 mode misses by 36x at this depth: 3.6e-2):
  * fp32 stream: every stage at least as close to the fp32 anchor as the reference-under-autocast is, logits within
    3e-2 rel-L2 / 4e-2 max-rel of the anchor (measured 2.28e-2 / 2.94e-2; the reference's own bf16 mode: 3.60e-2 / 3.98e-2);
  * bf16 stream: every stage within 1.25x the reference-under-autocast's own error, logits within 4e-2 / 5e-2 (measured 3.05e-2 / 3.36e-2);
  * engine vs the bf16-autocast reference (same dtype, two independent sets of bf16 roundings) within 1.6x the
    larger of the two anchor errors;
  * fp16 engine (EngineConfig(dtype='fp16'), the demo's mode): three more mantissa bits -- logits within 8e-3 rel-L2 /
    1.2e-2 max-rel of the fp32 anchor at 32 layers, every stage closer to the anchor than the bf16-autocast reference;
  * greedy next-token agreement with the fp32 anchor no worse than the bf16-autocast reference's own agreement minus
    one point (the disagreements are near-ties of random-init logits).
Weights: seeded random init of the real architecture with the reference's init scales (no checkpoints offline)."""
import pytest
import torch

from gpt4roi_b200.engine import EngineConfig, random_state_dicts
from oracle import model_oracle
from tests.test_engine_gpu import make_inputs

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def maxrel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def test_full_7b_forward_through_the_seam_vs_fp32_and_bf16_oracles():
    from tests.test_model_seam_gpu import build_seam_model
    import gc
    gc.collect()
    torch.cuda.empty_cache()          # blocks cached by earlier tests of the same process do not count as used
    free = torch.cuda.mem_get_info()[0]
    if free < 110e9:
        pytest.skip('needs ~100 GB of free HBM (fp32 7B oracle + bf16 oracle + engine); %.0f GB free' % (free / 1e9))
    cfg = EngineConfig(image_size=336, vit_layers=24, n_layers=32)
    sd, vit_sd = random_state_dicts(cfg, DEV, seed=2024)            # bf16 weights: exactly representable in fp32
    B, K, T = 2, 8, 128
    ids, images, boxes = make_inputs(cfg, B, [K] * B, T, seed=17)
    images = images.to(BF)
    depth = (8, 16, 24, 31)

    # ---- the product, through the model seam (default mode: bf16 residual stream) -------------------------
    from gpt4roi_b200.engine import PrefillEngine
    model = build_seam_model(cfg, sd, vit_sd, dtype=BF).eval()
    with torch.no_grad():
        out = model(input_ids=ids.to(DEV), attention_mask=torch.ones_like(ids).to(DEV), images=images.to(DEV),
                    img_metas=[None] * B, bboxes=[b.to(DEV) for b in boxes])
    logits = out.logits.float()
    assert logits.shape == (B, T + cfg.num_patches + 2, cfg.vocab) and torch.isfinite(logits).all()
    eng = model._get_engine(torch.device(DEV))
    assert eng.cfg.llama_stream == 'bf16'

    def run(engine):
        stage, htaps = {}, {n: None for n in depth}
        with torch.no_grad():
            lg = engine.forward(ids.to(DEV), images.to(DEV), boxes, hidden_taps=htaps, stage_taps=stage).float()
        d = dict(logits=lg, region=stage['region'].float(), embeds=stage['embeds'].float())
        for l, t in enumerate(stage['vit_taps']):
            d['vit%d' % cfg.level_layers[l]] = t.float()
        for n in depth:
            d['h%d' % n] = htaps[n].float()
        return d
    got = run(eng)
    assert torch.equal(got['logits'], logits)                        # same engine, no atomics: bitwise repeatable
    del model, eng, out
    torch.cuda.empty_cache()
    cfg16 = EngineConfig(image_size=336, vit_layers=24, n_layers=32, llama_stream='fp32')
    eng16 = PrefillEngine(cfg16, sd, vit_sd, DEV)
    got16 = run(eng16)
    del eng16
    torch.cuda.empty_cache()
    # fp16 mode (the demo's dtype, gpt4roi/app.py:74-98): the same weights (bf16 values are fp16-representable down to
    # 6e-5; below that the absolute error is < 6e-8) through the `_f16` kernels
    eng_h = PrefillEngine(EngineConfig(image_size=336, vit_layers=24, n_layers=32, dtype='fp16'), sd, vit_sd, DEV)
    assert eng_h.dt == torch.float16
    got_h = run(eng_h)
    del eng_h
    torch.cuda.empty_cache()

    # ---- oracles ------------------------------------------------------------------------------------------
    def collect(ref_logits, inter):
        d = dict(logits=ref_logits, region=torch.cat(inter['region']), embeds=inter['embeds'])
        for l, t in enumerate(inter['vit_taps']):
            d['vit%d' % cfg.level_layers[l]] = t
        for n in depth:
            d['h%d' % n] = inter['hidden'][n]
        return d
    r32 = collect(*model_oracle.forward(cfg, sd, vit_sd, ids, images.float(), boxes, DEV, return_intermediates=True,
                                        hidden_layers=depth))
    torch.cuda.empty_cache()
    r16 = collect(*model_oracle.forward(cfg, sd, vit_sd, ids, images, boxes, DEV, autocast_bf16=True,
                                        return_intermediates=True, hidden_layers=depth))
    order = ['vit%d' % l for l in cfg.level_layers] + ['region', 'embeds'] + ['h%d' % n for n in depth] + ['logits']
    agree16 = (r16['logits'].argmax(-1) == r32['logits'].argmax(-1)).float().mean().item()
    for name, g, slack, lim in (('bf16 stream (default)', got, 1.25, (4e-2, 5e-2)), ('fp32 stream', got16, 1.0, (3e-2, 4e-2)),
                                ('fp16 engine (demo dtype)', got_h, 1.0, (8e-3, 1.2e-2))):
        print('\n[%s]\nstage      | engine vs fp32      | bf16-ref vs fp32    | engine vs bf16-ref   (rel-L2 / max-rel)' % name)
        rows = {}
        for k in order:
            rows[k] = (rel(g[k], r32[k]), maxrel(g[k], r32[k]), rel(r16[k], r32[k]), maxrel(r16[k], r32[k]),
                       rel(g[k], r16[k]), maxrel(g[k], r16[k]))
            print('%-10s | %.3e %.3e | %.3e %.3e | %.3e %.3e' % ((k,) + rows[k]))
        agree = (g['logits'].argmax(-1) == r32['logits'].argmax(-1)).float().mean().item()
        print('greedy next-token agreement with the fp32 anchor: engine %.4f, bf16-autocast reference %.4f' % (agree, agree16))
        for k in order:
            e_eng, _, e_ref, _, e_same, _ = rows[k]
            assert e_eng <= max(slack * e_ref, 2e-3), (name, k, rows[k])
            assert e_same <= 1.6 * max(e_eng, e_ref), (name, k, rows[k])
        assert rows['logits'][0] < lim[0] and rows['logits'][1] < lim[1], (name, rows['logits'])
        assert agree >= min(0.97, agree16 - 0.01), (name, agree, agree16)
