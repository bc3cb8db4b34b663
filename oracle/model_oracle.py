"""oracle/model_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

End-to-end checker for the region-token forward: the reference's composition
(gpt4roi/models/spi_llava.py:23-205 + llava/model/llava.py:203-261) restated over
  * transformers' own CLIPVisionModel / LlamaForCausalLM (the third-party arithmetic the
    reference calls, pyproject.toml:19 -- installed version on the box, eager attention),
  * oracle/spi_oracle.py for the SPI module (pinned to the reference by golden fixtures),
  * the python-loop splice semantics (spi_llava.py:99-196).
PARITY UNPINNED for ViT/LLaMA: no reference test or golden vector pins their outputs (SURVEY.md
8(c)); the oracle for those blocks is the transformers build installed on the measurement box.
Runs in fp32 (accuracy anchor) or under bf16 autocast (the reference's operating mode).
"""
import contextlib

import torch

from oracle import spi_oracle


def build_vit(cfg, vit_sd, device, dtype=torch.float32):
    from transformers import CLIPVisionConfig, CLIPVisionModel
    vc = CLIPVisionConfig(hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_mlp,
                          num_hidden_layers=cfg.vit_layers, num_attention_heads=cfg.vit_heads,
                          image_size=cfg.image_size, patch_size=cfg.patch_size, layer_norm_eps=cfg.vit_eps,
                          hidden_act='quick_gelu')
    vc._attn_implementation = 'eager'
    m = CLIPVisionModel(vc)
    missing, unexpected = m.load_state_dict({k: v.float() for k, v in vit_sd.items()}, strict=False)
    assert not unexpected, unexpected
    assert all('post_layernorm' in k or 'position_ids' in k for k in missing), missing
    return m.to(device=device, dtype=dtype).eval()


def build_llm(cfg, sd, device, dtype=torch.float32):
    from transformers import LlamaConfig, LlamaForCausalLM
    lc = LlamaConfig(hidden_size=cfg.hidden, intermediate_size=cfg.mlp, num_hidden_layers=cfg.n_layers,
                     num_attention_heads=cfg.n_heads, num_key_value_heads=cfg.n_heads, vocab_size=cfg.vocab,
                     rms_norm_eps=cfg.rms_eps, max_position_embeddings=4096, rope_theta=cfg.rope_theta,
                     tie_word_embeddings=False, attention_bias=False, mlp_bias=False)
    lc._attn_implementation = 'eager'
    with torch.device('meta'):
        m = LlamaForCausalLM(lc)
    m = m.to_empty(device=device)
    keep = {k: v for k, v in sd.items() if k.startswith('model.layers.') or k in
            ('model.embed_tokens.weight', 'model.norm.weight', 'lm_head.weight')}
    missing, unexpected = m.load_state_dict({k: v.to(device=device, dtype=dtype) for k, v in keep.items()},
                                            strict=False, assign=True)
    assert not unexpected, unexpected
    assert all('rotary' in k or 'inv_freq' in k for k in missing), missing
    m = m.to(dtype=dtype).eval()
    # buffers created on meta must be re-materialised
    if hasattr(m.model, 'rotary_emb'):
        from transformers.models.llama.modeling_llama import LlamaRotaryEmbedding
        m.model.rotary_emb = LlamaRotaryEmbedding(lc, device=device)
    return m


@torch.no_grad()
def forward(cfg, sd, vit_sd, input_ids, images, bboxes, device, autocast_bf16=False, vit=None, llm=None,
            return_intermediates=False, hidden_layers=()):
    """Returns logits [B,L,V] (fp32 tensor).  Weights: fp32 copies of the given state dicts, or bf16
    copies when autocast_bf16 (the reference's deployment mode: bf16 weights + autocast)."""
    wdt = torch.bfloat16 if autocast_bf16 else torch.float32
    vit = vit or build_vit(cfg, vit_sd, device, wdt)
    llm = llm or build_llm(cfg, sd, device, wdt)
    ctx = torch.autocast('cuda', dtype=torch.bfloat16) if autocast_bf16 else contextlib.nullcontext()
    spi_sd = {k: v.to(device=device, dtype=wdt if autocast_bf16 else torch.float32) for k, v in sd.items()
              if k.startswith('model.spi_module.') or k.startswith('model.mm_projector.')}
    input_ids = input_ids.to(device)
    B, L = input_ids.shape
    inter = {}
    with ctx:
        out = vit(images.to(device=device, dtype=wdt), output_hidden_states=True)
        hs = out.hidden_states
        sel = cfg.select_layer
        feats = hs[sel][:, 1:]
        mlvl = [h[:, 1:] for h in hs[sel::-3][::-1][-cfg.num_levels:]]          # spi_llava.py:76-82
        inter['vit_taps'] = [m.float() for m in mlvl]
        region = None
        if bboxes is not None and len(bboxes) > 0:
            bb = [b.to(device=device, dtype=torch.float32) for b in bboxes]
            region = spi_oracle.roi_query_forward(spi_sd, [m.float() for m in mlvl] if not autocast_bf16 else mlvl,
                                                  bb, cfg.image_size, tuple(cfg.strides), cfg.roi_out,
                                                  cfg.roi_sampling)
            inter['region'] = [r.float() for r in region]
        img = torch.nn.functional.linear(feats, spi_sd['model.mm_projector.weight'], spi_sd['model.mm_projector.bias'])
        inter['img_rows'] = img.float()
        emb = llm.model.embed_tokens(input_ids)
        new = []
        for b in range(B):                                                       # spi_llava.py:101-195
            ids, cur = input_ids[b], emb[b]
            if (ids == cfg.im_patch_token).sum() == 0:
                new.append(cur)
                continue
            s = int(torch.where(ids == cfg.im_start_token)[0][0])
            P = img.shape[1]
            cur = torch.cat((cur[:s + 1], img[b].to(cur.dtype), cur[s + P + 1:]), 0)
            if region is not None:
                mask = ids == cfg.bbox_token
                spi = torch.zeros_like(cur)
                spi[mask] = region[b].to(cur.dtype)
                cur = cur * (~mask).to(cur.dtype)[:, None] + spi
            new.append(cur)
        embeds = torch.stack(new, 0)
        inter['embeds'] = embeds.float()
        if hidden_layers:
            # hidden_states[n] = residual stream after decoder layer n (n < n_layers; the last entry is post-norm)
            res = llm(inputs_embeds=embeds, use_cache=False, output_hidden_states=True)
            inter['hidden'] = {n: res.hidden_states[n].float() for n in hidden_layers}
            logits = res.logits
        else:
            logits = llm(inputs_embeds=embeds, use_cache=False).logits
    if return_intermediates:
        return logits.float(), inter
    return logits.float()
