"""Training step of the LLaMA stack (first slice) against fp32 autograd of
transformers' LlamaForCausalLM (oracle/model_oracle.build_llm -- the arithmetic the reference trains through)
on identical bf16-representable weights and inputs.

Tolerances: bf16 activations/gradients with fp32 accumulation vs an fp32 reference -> loss within 2e-3
relative, every gradient tensor within rel-L2 3e-2 (2 layers, hidden 4096)."""
import contextlib

import pytest
import torch

from tests.helpers import golden_stream_mismatch

from gpt4roi_b200.engine import EngineConfig, random_state_dicts
from gpt4roi_b200.train import LAYER_KEYS, LlamaTrainStack, train_step
from oracle import model_oracle

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def _setup(n_layers=2, B=2, L=96, seed=0):
    cfg = EngineConfig(image_size=224, vit_layers=0, n_layers=n_layers)
    sd, _ = random_state_dicts(cfg, DEV, seed=seed)
    sd = {k: v.to(BF).float() for k, v in sd.items()}          # bf16-representable fp32 weights
    g = torch.Generator(device='cpu').manual_seed(seed + 1)
    x = (torch.randn(B, L, cfg.hidden, generator=g) * 0.5).to(DEV, BF)
    labels = torch.randint(0, cfg.vocab, (B, L), generator=g).to(DEV)
    labels[:, : L // 3] = -100                                   # prompt part is not supervised
    targets = torch.full_like(labels, -100)
    targets[:, :-1] = labels[:, 1:]
    return cfg, sd, x, labels, targets


def test_forward_backward_match_hf_autograd():
    cfg, sd, x, labels, targets = _setup()
    stack = LlamaTrainStack(cfg, sd, DEV)
    loss = stack.forward(x, targets)
    d_in = stack.backward()
    grads = stack.grads

    llm = model_oracle.build_llm(cfg, sd, DEV, torch.float32).train()
    xr = x.float().requires_grad_()
    out = llm(inputs_embeds=xr, labels=labels)
    out.loss.backward()
    assert abs(loss.item() - out.loss.item()) <= 2e-3 * abs(out.loss.item()), (loss.item(), out.loss.item())
    errs = {'d_inputs_embeds': rel(d_in, xr.grad)}
    for i, lyr in enumerate(llm.model.layers):
        want = {
            'ln_in': lyr.input_layernorm.weight.grad, 'ln_post': lyr.post_attention_layernorm.weight.grad,
            'wqkv': torch.cat([lyr.self_attn.q_proj.weight.grad, lyr.self_attn.k_proj.weight.grad,
                               lyr.self_attn.v_proj.weight.grad], 0),
            'wo': lyr.self_attn.o_proj.weight.grad,
            'wgu': torch.stack([lyr.mlp.gate_proj.weight.grad, lyr.mlp.up_proj.weight.grad], 1).reshape(2 * cfg.mlp, cfg.hidden),
            'wdown': lyr.mlp.down_proj.weight.grad,
        }
        for k in LAYER_KEYS:
            errs['L%d.%s' % (i, k)] = rel(grads['layers'][i][k], want[k])
    errs['norm'] = rel(grads['top']['norm'], llm.model.norm.weight.grad)
    errs['lm_head'] = rel(grads['top']['lm_head'], llm.lm_head.weight.grad)
    print('max grad rel-L2 error: %.3e (%s)' % (max(errs.values()), max(errs, key=errs.get)))
    bad = {k: v for k, v in errs.items() if not v < 3e-2}
    assert not bad, bad


def test_backward_is_bitwise_reproducible_and_step_lowers_loss():
    cfg, sd, x, labels, targets = _setup(seed=3)
    stack = LlamaTrainStack(cfg, sd, DEV, lr=1e-3)
    stack.forward(x, targets)
    d1 = stack.backward()
    g1 = stack.grads
    stack.forward(x, targets)
    d2 = stack.backward()
    g2 = stack.grads
    assert torch.equal(d1, d2)
    for a, b in zip(g1['layers'], g2['layers']):
        for k in LAYER_KEYS:
            assert torch.equal(a[k], b[k]), k
    assert torch.equal(g1['top']['lm_head'], g2['top']['lm_head'])
    # a few AdamW steps on one batch must reduce its loss (optimizer wiring, bf16 weight refresh)
    losses = []
    for _ in range(4):
        loss, _ = train_step(stack, x, targets)
        losses.append(loss.item())
    assert losses[-1] < losses[0] - 0.05, losses


def test_frontend_backward_matches_autograd():
    """Second slice: splice / embedding / mm_projector / SPI-head backward vs fp32 autograd of the same algebra
    (index_put splice, F.linear) on the tensors the forward saved.  rel-L2 <= 1.5e-2 per gradient."""
    import torch.nn.functional as F
    from gpt4roi_b200.engine import PrefillEngine
    from gpt4roi_b200.train import FrontEndTrain
    from tests.test_engine_gpu import make_inputs
    cfg = EngineConfig(image_size=224, vit_layers=24, n_layers=0)
    sd, vit_sd = random_state_dicts(cfg, DEV, seed=5)
    eng = PrefillEngine(cfg, sd, vit_sd, DEV)
    ids, images, boxes = make_inputs(cfg, 2, [3, 1], 40, seed=2)
    ids[1, -1] = ids[1, -2] = ids[0, 5] if int(ids[0, 5]) < 32000 else 17     # repeated token ids share a row
    fe = FrontEndTrain(eng, head_only=True)
    embeds = fe.forward(ids, images, boxes)
    saved = dict(fe.saved)
    torch.manual_seed(0)
    d = (torch.randn_like(embeds.float()) * 0.1).to(BF)
    got = fe.backward(d)

    # fp32 autograd reference of the same computation from the saved activations
    B, L = ids.shape
    P, Hd = cfg.num_patches, cfg.hidden
    emb = eng.embed.float().requires_grad_()
    pw, pb = eng.proj_w.float().requires_grad_(), eng.proj_b.float().requires_grad_()
    uw, ub = eng.up_w.float().requires_grad_(), eng.up_b.float().requires_grad_()
    fw, fb = eng.flat_w.float().requires_grad_(), eng.flat_b.float().requires_grad_()
    pc = saved['pc'].float().requires_grad_()
    boxes_dev = torch.cat([b for b in boxes]).to(DEV)
    from gpt4roi_b200 import kernels
    pos = kernels.pos_embed_mlp(boxes_dev.contiguous(), *eng.pos).float().requires_grad_()
    img = F.linear(saved['feat'].float(), pw, pb).view(B, P, Hd)
    t = (F.linear(pc, fw) + fb + pos)
    region = F.linear(t.detach().to(BF).float() + (t - t.detach()), uw, ub)  # bf16 rounding of t, straight-through
    codes = saved['plan'].reshape(-1).long()
    rows = []
    for r, code in enumerate(codes.tolist()):
        b = r // L
        if code & 0x40000000:
            rows.append(img[b, code & 0x1FFFFFFF])
        elif code & 0x20000000:
            rows.append(region[code & 0x1FFFFFFF])
        else:
            rows.append(emb[code])
    ref = torch.stack(rows).view(B, L, Hd)
    for name, t_ in (('embeds', embeds), ('feat', saved['feat']), ('pc', saved['pc']), ('pos', pos), ('img', img), ('region', region)):
        assert torch.isfinite(t_.float()).all(), 'non-finite values in %s' % name
    assert rel(embeds, ref) < 6e-3
    (ref * d.float()).sum().backward()
    q = 'model.spi_module.roi_align.'
    R, C = cfg.roi_out, cfg.spi_dim
    want = {'model.embed_tokens.weight': emb.grad, 'model.mm_projector.weight': pw.grad, 'model.mm_projector.bias': pb.grad,
            q + 'updims.weight': uw.grad, q + 'updims.bias': ub.grad, q + 'flatten_linear.bias': fb.grad,
            q + 'flatten_linear.weight': fw.grad.view(-1, R, R, C).permute(0, 3, 1, 2).reshape(fw.shape[0], -1),
            'd_pos': pos.grad, 'd_pconv_out': pc.grad}
    errs = {k: rel(got[k], v) for k, v in want.items()}
    print('front-end grads rel-L2:', {k.split('.')[-2] + '.' + k.split('.')[-1] if '.' in k else k: '%.2e' % v for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < 1.5e-2}
    assert not bad, bad
    # untouched embedding rows have exactly zero gradient
    used = torch.zeros(cfg.vocab, dtype=torch.bool, device=DEV)
    used[codes[codes < 0x20000000]] = True
    assert torch.all(got['model.embed_tokens.weight'][~used] == 0)


def test_spi_module_backward_matches_oracle_autograd():
    """Complete backward of the SPI module (fuse stack, RoIAlign, pconvs, head, position MLP) vs fp32 autograd
    through oracle/spi_oracle.py (the torch restatement pinned to the reference's MLVLROIQueryModule by the golden
    fixtures).  Tolerance: every parameter gradient is within rel-L2 2e-2 of the fp32 anchor, or at least as close to
    it as 1.5x the error of the same oracle run under bf16 autocast (the reference's training mode): the 5-round
    conv / GroupNorm / ReLU stack amplifies bf16 rounding (ReLU masks flip), for the reference as for this engine."""
    from gpt4roi_b200.engine import PrefillEngine
    from gpt4roi_b200.train import SpiTrain
    from oracle import spi_oracle
    cfg = EngineConfig(image_size=224, vit_layers=24, n_layers=0)
    cfg.select_index = 0
    sd, _ = random_state_dicts(cfg, DEV, seed=11)
    sd = {k: v.to(BF).float() for k, v in sd.items()}
    vit_dummy = random_state_dicts(EngineConfig(image_size=224, vit_layers=0, n_layers=0), 'cpu')[1]
    eng = PrefillEngine(cfg, sd, vit_dummy, DEV)
    g = torch.Generator().manual_seed(4)
    B, P, C = 2, cfg.num_patches, cfg.spi_dim
    toks = [(torch.randn(B, P, C, generator=g) * 0.7).to(BF).float().to(DEV) for _ in range(cfg.num_levels)]
    boxes = []
    for k in (3, 2):
        pts = torch.rand(k, 2, 2, generator=g).sort(dim=1).values
        bx = torch.cat([pts[:, 0, :], pts[:, 1, :]], 1)
        bx[:, 2:] = torch.maximum(bx[:, 2:], bx[:, :2] + 0.1).clamp(max=1.0)
        boxes.append(bx.to(DEV))
    taps = {layer: torch.cat([torch.zeros(B, 1, C, device=DEV), toks[l]], 1).contiguous()
            for l, layer in enumerate(cfg.level_layers)}
    spi = SpiTrain(eng)
    region = spi.forward(taps, eng.plan_boxes(boxes))
    torch.manual_seed(0)
    d = (torch.randn(region.shape, device=DEV) * 0.05).to(BF)
    got = spi.backward(d)

    def oracle_grads(autocast):
        ref_sd = {k: v.clone().requires_grad_() for k, v in sd.items() if k.startswith('model.spi_module.')}
        ctx = torch.autocast('cuda', dtype=BF) if autocast else contextlib.nullcontext()
        with ctx:
            out = torch.cat(spi_oracle.roi_query_forward(ref_sd, toks, boxes, cfg.image_size), 0)
        (out.float() * d.float()).sum().backward()
        return out.float(), {k: v.grad for k, v in ref_sd.items() if v.grad is not None}

    out32, g32 = oracle_grads(False)
    out16, g16 = oracle_grads(True)            # the reference's operating mode: bf16 autocast
    assert rel(region, out32) < 3e-2
    assert set(g32) == {k for k in sd if k.startswith('model.spi_module.')}   # every SPI parameter is trained
    errs, errs16 = {}, {}
    for k, v in g32.items():
        assert k in got, k
        errs[k[len('model.spi_module.'):]] = rel(got[k].reshape(v.shape), v)
        errs16[k[len('model.spi_module.'):]] = rel(g16[k].float(), v)
    rows = sorted(errs.items(), key=lambda kv: -kv[1])
    print('SPI backward rel-L2 vs fp32 autograd (ours | reference under bf16 autocast):')
    for k, v in rows:
        print('   %-40s %.2e | %.2e' % (k, v, errs16[k]))
    # at least as close to the fp32 anchor as 1.5x the reference's own bf16-autocast run, or within 2e-2 outright
    bad = {k: (v, errs16[k]) for k, v in errs.items() if not (v < 2e-2 or v < 1.5 * errs16[k])}
    assert not bad, bad


def test_stage2_step_end_to_end_matches_autograd():
    """Whole stage-2 step (frozen CLIP tower -> projector / SPI module / embeddings -> splice -> LLaMA -> loss):
    every trained parameter's gradient vs fp32 autograd through the oracle composition (spi_oracle + index splice +
    transformers LlamaForCausalLM), calibrated like the SPI test against the same composition under bf16 autocast
    (the reference's training mode).  Then one AdamW step must lower the loss of the batch."""
    import torch.nn.functional as F
    from gpt4roi_b200.train import Stage2Trainer
    from oracle import spi_oracle
    from tests.test_engine_gpu import make_inputs
    cfg = EngineConfig(image_size=224, vit_layers=24, n_layers=2)
    sd, vit_sd = random_state_dicts(cfg, DEV, seed=21)
    sd = {k: v.to(BF).float() for k, v in sd.items()}
    tr = Stage2Trainer(cfg, sd, vit_sd, DEV, lr=1e-3)
    ids, images, boxes = make_inputs(cfg, 2, [2, 1], 24, seed=6)
    labels = ids.clone()
    labels[:, : cfg.num_patches + 8] = -100
    labels[ids == cfg.bbox_token] = -100
    loss = tr.forward_backward(ids, images, boxes, labels)
    got = dict(tr.front_grads)
    for i, gl in enumerate(tr.stack.grads['layers']):
        for k in LAYER_KEYS:
            got['L%d.%s' % (i, k)] = gl[k]
    got['norm'], got['lm_head'] = tr.stack.grads['top']['norm'], tr.stack.grads['top']['lm_head']

    # ---- oracle composition (CLIP outputs are constants: the tower is frozen)
    eng = tr.eng
    taps = eng.vit(images.to(DEV, BF))
    toks = [taps[layer][:, 1:].float().contiguous() for layer in cfg.level_layers]
    feat = taps[cfg.select_index][:, 1:].float()
    B, L = ids.shape
    P, Hd = cfg.num_patches, cfg.hidden
    boxes_dev = [b.to(DEV) for b in boxes]
    fe_plan = None

    def reference(autocast):
        ref = {k: v.clone().to(DEV).requires_grad_() for k, v in sd.items()
               if k.startswith('model.spi_module.') or k.startswith('model.mm_projector.') or k == 'model.embed_tokens.weight'}
        llm = model_oracle.build_llm(cfg, sd, DEV, torch.float32).train()
        ctx = torch.autocast('cuda', dtype=BF) if autocast else contextlib.nullcontext()
        with ctx:
            region = torch.cat(spi_oracle.roi_query_forward(ref, toks, boxes_dev, cfg.image_size), 0)
            img = F.linear(feat, ref['model.mm_projector.weight'], ref['model.mm_projector.bias'])
            emb = ref['model.embed_tokens.weight']
            rows, k = [], 0
            for b in range(B):
                s0 = int((ids[b] == cfg.im_start_token).nonzero()[0])
                for t in range(L):
                    tok = int(ids[b, t])
                    if s0 < t <= s0 + P:
                        rows.append(img[b, t - s0 - 1].float())
                    elif tok == cfg.bbox_token:
                        rows.append(region[k].float())
                        k += 1
                    else:
                        rows.append(emb[tok])
            embeds = torch.stack(rows).view(B, L, Hd)
            out = llm(inputs_embeds=embeds, labels=labels.to(DEV))
        out.loss.backward()
        g = {k: v.grad for k, v in ref.items()}
        for i, lyr in enumerate(llm.model.layers):
            g['L%d.ln_in' % i] = lyr.input_layernorm.weight.grad
            g['L%d.ln_post' % i] = lyr.post_attention_layernorm.weight.grad
            g['L%d.wqkv' % i] = torch.cat([lyr.self_attn.q_proj.weight.grad, lyr.self_attn.k_proj.weight.grad, lyr.self_attn.v_proj.weight.grad], 0)
            g['L%d.wo' % i] = lyr.self_attn.o_proj.weight.grad
            g['L%d.wgu' % i] = torch.stack([lyr.mlp.gate_proj.weight.grad, lyr.mlp.up_proj.weight.grad], 1).reshape(2 * cfg.mlp, cfg.hidden)
            g['L%d.wdown' % i] = lyr.mlp.down_proj.weight.grad
        g['norm'], g['lm_head'] = llm.model.norm.weight.grad, llm.lm_head.weight.grad
        return out.loss.item(), g

    l32, g32 = reference(False)
    l16, g16 = reference(True)
    assert abs(loss.item() - l32) <= max(3e-3 * abs(l32), 1.5 * abs(l16 - l32)), (loss.item(), l32, l16)
    assert set(got) == set(g32), set(got) ^ set(g32)
    errs = {k: (rel(got[k].reshape(v.shape), v), rel(g16[k].float(), v)) for k, v in g32.items()}
    worst = sorted(errs.items(), key=lambda kv: -kv[1][0])[:8]
    print('stage-2 step, loss %.4f (fp32 oracle %.4f, bf16-autocast oracle %.4f); worst gradient rel-L2 (ours | autocast):' % (loss.item(), l32, l16))
    for k, (a, b) in worst:
        print('   %-50s %.2e | %.2e' % (k, a, b))
    bad = {k: v for k, v in errs.items() if not (v[0] < 2e-2 or v[0] < 1.5 * v[1])}
    assert not bad, bad
    tr.optimizer_step()
    loss2 = tr.forward_backward(ids, images, boxes, labels)
    assert loss2.item() < loss.item(), (loss.item(), loss2.item())


def test_right_padded_batch_needs_no_mask_in_causal_training():
    """The collator right-pads (data_modules.py:33-44) and sets labels to -100 on the padding: with causal attention a
    valid position never attends to a padded key and padded queries receive zero gradient, so the stack's gradients
    equal those of transformers run WITH the attention mask.  Same tolerances as the unpadded test."""
    cfg, sd, x, labels, targets = _setup(B=2, L=80, seed=5)
    lens = [80, 53]
    mask = torch.zeros(2, 80, dtype=torch.long, device=DEV)
    for b, n in enumerate(lens):
        mask[b, :n] = 1
    labels = labels.clone()
    labels[mask == 0] = -100
    targets = torch.full_like(labels, -100)
    targets[:, :-1] = labels[:, 1:]
    stack = LlamaTrainStack(cfg, sd, DEV)
    loss = stack.forward(x, targets)
    d_in = stack.backward()
    llm = model_oracle.build_llm(cfg, sd, DEV, torch.float32).train()
    xr = x.float().requires_grad_()
    out = llm(inputs_embeds=xr, attention_mask=mask, labels=labels)
    out.loss.backward()
    assert abs(loss.item() - out.loss.item()) <= 2e-3 * abs(out.loss.item())
    valid = mask.bool()
    assert rel(d_in[valid], xr.grad[valid]) < 3e-2
    assert torch.all(d_in[~valid] == 0)                       # padded positions: exactly zero gradient
    for i, lyr in enumerate(llm.model.layers):
        assert rel(stack.grads['layers'][i]['wo'], lyr.self_attn.o_proj.weight.grad) < 3e-2
        assert rel(stack.grads['layers'][i]['wdown'], lyr.mlp.down_proj.weight.grad) < 3e-2
    assert rel(stack.grads['top']['lm_head'], llm.lm_head.weight.grad) < 3e-2


def test_stage2_step_matches_reference_autograd_golden():
    """tests/golden/train_step_ref_224.npz = loss and probe gradients of the reference's OWN
    SPILlavaMPTForCausalLM.forward(labels=...) + loss.backward() (unmodified modules under tests/golden/ref_shims.py,
    fp32, CPU, CLIP frozen; generated by `make_golden.py --train`).  The sm_100a trainer starts from the same
    bf16-representable weights.  Tolerances (rel-L2 vs the fp32 golden): loss 3e-3; LLaMA / lm_head / projector /
    SPI-head tensors 3e-2; pconvs 8e-2; fuse / input convs and their GroupNorms 1.5e-1 (x2.5 for the 2-channel row probes
    of conv weights) -- the depth-dependent bf16
    error that the reference's own bf16-autocast mode shows against fp32 (test_spi_module_backward_...)."""
    import os
    import numpy as np
    from gpt4roi_b200.train import Stage2Trainer
    from tests.golden.make_golden import train_inputs
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'train_step_ref_224.npz'), allow_pickle=False)
    cfg, ids, images, boxes, labels = train_inputs()
    sd, vit_sd = random_state_dicts(cfg, 'cpu', seed=1234, dtype=torch.float32)
    sd = {k: v.to(BF).float() for k, v in sd.items()}
    vit_sd = {k: v.to(BF).float() for k, v in vit_sd.items()}
    if int(ids.sum()) != int(gold['ids_checksum'][0]) or \
            not np.allclose(float(sd['lm_head.weight'].double().sum()), gold['w_checksum'][0], rtol=1e-9):
        golden_stream_mismatch()
    tr = Stage2Trainer(cfg, sd, vit_sd, DEV)
    loss = tr.forward_backward(ids, images, boxes, labels).item()
    want_loss = float(gold['loss'][0])
    assert abs(loss - want_loss) <= 3e-3 * abs(want_loss), (loss, want_loss)

    H = cfg.hidden
    got = {k: v for k, v in tr.front_grads.items()}
    for i, g in enumerate(tr.stack.grads['layers']):
        q = 'model.layers.%d.' % i
        got[q + 'self_attn.q_proj.weight'], got[q + 'self_attn.k_proj.weight'], got[q + 'self_attn.v_proj.weight'] = \
            g['wqkv'][:H], g['wqkv'][H:2 * H], g['wqkv'][2 * H:]
        got[q + 'self_attn.o_proj.weight'] = g['wo']
        got[q + 'mlp.gate_proj.weight'], got[q + 'mlp.up_proj.weight'] = g['wgu'][0::2], g['wgu'][1::2]
        got[q + 'mlp.down_proj.weight'] = g['wdown']
        got[q + 'input_layernorm.weight'], got[q + 'post_attention_layernorm.weight'] = g['ln_in'], g['ln_post']
    got['model.norm.weight'], got['lm_head.weight'] = tr.stack.grads['top']['norm'], tr.stack.grads['top']['lm_head']

    def tol(name):
        t = 3e-2
        if 'mlvl_fuse' in name:
            t = 1.5e-1
        elif 'pconvs' in name:
            t = 8e-2
        elif 'pos_embedd' in name:
            t = 6e-2
        if name.endswith(']') and 'conv' in name:
            t *= 2.5   # a 2-of-1024 output-channel slice of a conv gradient: a handful of flipped ReLU masks dominate it
        return t
    names = [str(n) for n in gold['norm_names']]
    assert int(gold['n_trained_tensors'][0]) == len(names) == len(got), (len(names), len(got))
    errs = {}
    for key in gold.files:
        if key.startswith('full/'):
            k = key[5:]
            errs[k] = rel(got[k].reshape(gold[key].shape), torch.from_numpy(gold[key]).to(DEV))
        elif key.startswith('rows/'):
            k = key[5:]
            n = gold[key].shape[0]
            errs[k + '[:%d]' % n] = rel(got[k].reshape((-1,) + gold[key].shape[1:])[:n], torch.from_numpy(gold[key]).to(DEV))
    e_rows = torch.from_numpy(gold['embed_rows']).to(DEV)
    errs['model.embed_tokens.weight[rows]'] = rel(got['model.embed_tokens.weight'][torch.from_numpy(gold['embed_ids']).to(DEV)], e_rows)
    for n, want in zip(names, gold['norms']):                       # gradient norm of EVERY trained tensor
        have = got[n].float().norm().item()
        assert abs(have - float(want)) <= 2 * tol(n) * float(want) + 1e-12, (n, have, float(want))
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print('vs reference autograd golden: loss %.5f (reference %.5f); worst probes: %s'
          % (loss, want_loss, ', '.join('%s %.2e' % (k.replace('model.', ''), v) for k, v in worst)))
    bad = {k: v for k, v in errs.items() if not v < tol(k)}
    assert not bad, bad
