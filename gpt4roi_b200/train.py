"""Training step of the LLaMA stack (SURVEY.md 8(a) row 14 -- first slice).

The reference's stage-2 step (gpt4roi/train/train.py:698-712: HF Trainer over SPILlavaMPTForCausalLM under
bf16 autocast, DDP gradient all-reduce, AdamW with the parameter groups of llava_trainer.py:59-144) spends
6.48 B of its 7.04 B trainable parameters in the 32 LLaMA decoder layers.  This module is that part, built on
the sm_100a kernels only:

    inputs_embeds (from the region-token prefill front end: engine.PrefillEngine splice)
      -> 32 x [RMSNorm -> fused QKV GEMM + RoPE -> causal attention (saves LSE) -> o_proj + residual
               -> RMSNorm -> gate/up GEMM -> SwiGLU -> down_proj + residual]
      -> RMSNorm -> lm_head -> shifted cross entropy (llava/model/llava.py:238-249)
    backward: cross-entropy gradient -> lm_head -> ... every Linear through g4r_gemm_bf16_t
              (grad_x = grad_y.W, grad_W = grad_y^T.x, transposes in the UMMA descriptors), flash-attention
              backward, RoPE^T (= RoPE with -sin), SwiGLU / RMSNorm backward
    step: gradient all-reduce over NCCL (bucket = one decoder layer, overlapped with the rest of the
          backward on a side stream) + fused AdamW on fp32 master weights, bf16 copies for the next forward.

Returned beside the loss: the gradient w.r.t. inputs_embeds, which is what the not-yet-built backward of the
splice / projector / SPI module (RoIAlign backward kernels exist) will consume.

Precision: fp32 master weights and optimizer moments, bf16 weights/activations/gradients, fp32 accumulation
in every GEMM and reduction.  The reference keeps the residual stream in fp32 under autocast (fp32 embeddings,
`residual + bf16_branch` promotes); this first slice keeps it in bf16 like the inference path -- a stated
deviation, covered by the tolerance of the parity test (tests/test_train_gpu.py).
No PyTorch autograd, no torch.nn ops on the compute path: torch supplies device memory, streams and NCCL.
"""
import torch

from . import dense, kernels, train_ops

BF16 = torch.bfloat16
F32 = torch.float32

LAYER_KEYS = ('ln_in', 'wqkv', 'wo', 'ln_post', 'wgu', 'wdown')


def fuse_llama_layer(state_dict, i, device=None, dtype=F32):
    """HF LLaMA layer i -> the engine's fused layout: wqkv = [q; k; v], wgu = rows (gate_0, up_0, gate_1, up_1, ...)."""
    q = 'model.layers.%d.' % i
    cvt = lambda t: t.detach().to(device=device if device is not None else t.device, dtype=dtype).contiguous()
    wqkv = torch.cat([cvt(state_dict[q + 'self_attn.%s_proj.weight' % n]) for n in 'qkv'], 0).contiguous()
    g, u = cvt(state_dict[q + 'mlp.gate_proj.weight']), cvt(state_dict[q + 'mlp.up_proj.weight'])
    wgu = torch.stack([g, u], 1).reshape(2 * g.shape[0], g.shape[1]).contiguous()
    return dict(ln_in=cvt(state_dict[q + 'input_layernorm.weight']), wqkv=wqkv,
                wo=cvt(state_dict[q + 'self_attn.o_proj.weight']),
                ln_post=cvt(state_dict[q + 'post_attention_layernorm.weight']), wgu=wgu,
                wdown=cvt(state_dict[q + 'mlp.down_proj.weight']))


def unfuse_llama_layer(layer, i):
    """Inverse of fuse_llama_layer: fused layer dict -> HF / reference parameter names (views, no copies)."""
    q = 'model.layers.%d.' % i
    H = layer['wqkv'].shape[0] // 3
    return {q + 'self_attn.q_proj.weight': layer['wqkv'][:H], q + 'self_attn.k_proj.weight': layer['wqkv'][H:2 * H],
            q + 'self_attn.v_proj.weight': layer['wqkv'][2 * H:], q + 'self_attn.o_proj.weight': layer['wo'],
            q + 'mlp.gate_proj.weight': layer['wgu'][0::2], q + 'mlp.up_proj.weight': layer['wgu'][1::2],
            q + 'mlp.down_proj.weight': layer['wdown'], q + 'input_layernorm.weight': layer['ln_in'],
            q + 'post_attention_layernorm.weight': layer['ln_post']}


def save_checkpoint(state_dict, out_dir, max_shard_bytes=10 * 1024 ** 3, dtype=None):
    """Write `state_dict` (reference parameter names) the way the reference's trainer does
    (train.py:88-98 -> Trainer._save -> save_pretrained): `pytorch_model-XXXXX-of-YYYYY.bin` shards plus
    `pytorch_model.bin.index.json`, loadable by SPILlavaMPTForCausalLM.from_pretrained.  Returns the shard file names."""
    import json
    import os
    os.makedirs(out_dir, exist_ok=True)
    shards, cur, cur_bytes = [], {}, 0
    for k, v in state_dict.items():
        t = v.detach().to('cpu', dtype if dtype is not None else v.dtype).contiguous()
        nbytes = t.numel() * t.element_size()
        if cur and cur_bytes + nbytes > max_shard_bytes:
            shards.append(cur)
            cur, cur_bytes = {}, 0
        cur[k] = t
        cur_bytes += nbytes
    if cur:
        shards.append(cur)
    names = ['pytorch_model-%05d-of-%05d.bin' % (i + 1, len(shards)) for i in range(len(shards))]
    weight_map, total = {}, 0
    for name, shard in zip(names, shards):
        torch.save(shard, os.path.join(out_dir, name))
        for k, t in shard.items():
            weight_map[k] = name
            total += t.numel() * t.element_size()
    with open(os.path.join(out_dir, 'pytorch_model.bin.index.json'), 'w') as f:
        json.dump({'metadata': {'total_size': total}, 'weight_map': weight_map}, f, indent=2, sort_keys=True)
    return names


def load_checkpoint(ckpt_dir):
    """Read a sharded HF checkpoint directory written by save_checkpoint / save_pretrained into one state dict."""
    import json
    import os
    with open(os.path.join(ckpt_dir, 'pytorch_model.bin.index.json')) as f:
        index = json.load(f)
    out = {}
    for name in sorted(set(index['weight_map'].values())):
        out.update(torch.load(os.path.join(ckpt_dir, name), map_location='cpu'))
    return out


class LlamaTrainStack:
    """LLaMA decoder stack with explicit forward / backward / AdamW on the sm_100a kernels."""

    def __init__(self, cfg, state_dict, device, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        """cfg: engine.EngineConfig; state_dict: HF names (model.layers.N..., model.norm.weight, lm_head.weight).
        weight_decay follows llava_trainer.py:59-144: decay on matrices, none on norm weights."""
        self.cfg, self.dev = cfg, torch.device(device)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.step_count = 0
        if cfg.head_dim != 128:
            raise ValueError('LlamaTrainStack: head_dim 128 required (fused-RoPE QKV GEMM)')
        f32 = lambda t: t.detach().to(self.dev, F32).contiguous()
        self.master = [fuse_llama_layer(state_dict, i, self.dev) for i in range(cfg.n_layers)]   # fp32 masters
        self.master_top = dict(norm=f32(state_dict['model.norm.weight']), lm_head=f32(state_dict['lm_head.weight']))
        self.w = [{k: v.to(BF16) for k, v in m.items()} for m in self.master]      # bf16 compute copies
        self.w_top = {k: v.to(BF16) for k, v in self.master_top.items()}
        zeros = lambda t: torch.zeros_like(t)
        self.m1 = [{k: zeros(v) for k, v in m.items()} for m in self.master]
        self.m2 = [{k: zeros(v) for k, v in m.items()} for m in self.master]
        self.m1_top = {k: zeros(v) for k, v in self.master_top.items()}
        self.m2_top = {k: zeros(v) for k, v in self.master_top.items()}
        self._rope_cache = {}
        self.saved = None
        self.grads = None

    def state_dict(self):
        """fp32 master weights under the reference's parameter names (views of the fused tensors)."""
        out = {}
        for i, m in enumerate(self.master):
            out.update(unfuse_llama_layer(m, i))
        out['model.norm.weight'], out['lm_head.weight'] = self.master_top['norm'], self.master_top['lm_head']
        return out

    # ------------------------------------------------------------------ helpers
    def _rope(self, L):
        if L not in self._rope_cache:
            c = self.cfg
            inv = 1.0 / (c.rope_theta ** (torch.arange(0, c.head_dim, 2, dtype=torch.int64).float() / c.head_dim))
            emb = torch.cat([torch.arange(L, dtype=F32)[:, None] * inv[None, :]] * 2, -1)
            cos, sin = emb.cos().to(self.dev, BF16).contiguous(), emb.sin().to(self.dev, BF16).contiguous()
            self._rope_cache[L] = (cos, sin, (-sin).contiguous())
        return self._rope_cache[L]

    # ------------------------------------------------------------------ forward
    def forward(self, inputs_embeds, targets):
        """inputs_embeds [B, L, hidden] bf16; targets int64 [B, L]: labels shifted left by one
        (targets[:, t] = labels[:, t+1], targets[:, -1] = -100).  Returns the mean loss (0-d fp32 tensor)."""
        c = self.cfg
        B, L, Hd = inputs_embeds.shape
        M = B * L
        cos, sin, _ = self._rope(L)
        scale = c.head_dim ** -0.5
        x = inputs_embeds.reshape(M, Hd).contiguous()
        saved = []
        for w in self.w:
            h1 = kernels.rmsnorm(x, w['ln_in'], c.rms_eps)
            qkv = dense.qkv_rope(h1, w['wqkv'], cos, sin, L, 2 * c.hidden)
            a, lse = train_ops.attention_fwd_lse(qkv, B, L, c.n_heads, c.head_dim, True, scale)
            x_mid = dense.linear(a, w['wo'], residual=x)
            h2 = kernels.rmsnorm(x_mid, w['ln_post'], c.rms_eps)
            gu = dense.linear(h2, w['wgu'])
            f = train_ops.swiglu_fwd(gu)
            x_out = dense.linear(f, w['wdown'], residual=x_mid)
            saved.append(dict(x_in=x, h1=h1, qkv=qkv, a=a, lse=lse, x_mid=x_mid, h2=h2, gu=gu, f=f))
            x = x_out
        hn = kernels.rmsnorm(x, self.w_top['norm'], c.rms_eps)
        vpad = (c.vocab + 63) // 64 * 64                       # 16-byte-aligned logits rows
        logits = torch.empty((M, vpad), dtype=BF16, device=self.dev)[:, :c.vocab]
        dense.linear(hn, self.w_top['lm_head'], out=logits)
        self.saved = dict(layers=saved, x_last=x, hn=hn, logits=logits, targets=targets.reshape(M).contiguous(), B=B, L=L)
        loss, count, _ = train_ops.cross_entropy(logits, self.saved['targets'], want_grad=False)
        return loss

    # ------------------------------------------------------------------ backward
    def backward(self, loss_scale=1.0, on_layer_grads=None):
        """Gradients of loss_scale * loss.  Fills self.grads (bf16 matrices, fp32 norm weights) and returns the
        gradient w.r.t. inputs_embeds [B, L, hidden] (bf16).  on_layer_grads(i, grads_i) is called as soon as
        decoder layer i's gradients are complete (DDP bucket hook)."""
        c, s = self.cfg, self.saved
        B, L = s['B'], s['L']
        _, _, nsin = self._rope(L)
        cos = self._rope(L)[0]
        scale = c.head_dim ** -0.5
        _, _, dlogits = train_ops.cross_entropy(s['logits'], s['targets'], grad_scale=loss_scale, want_grad=True)
        g_top = {}
        g_top['lm_head'] = dense.matmul_t(dlogits, s['hn'], a_mn=True, b_mn=True)           # dW = dY^T X
        dhn = dense.matmul_t(dlogits, self.w_top['lm_head'], b_mn=True)                      # dX = dY W
        del dlogits
        dx, g_top['norm'] = train_ops.rmsnorm_bwd(s['x_last'], self.w_top['norm'], dhn, c.rms_eps)
        grads = [None] * len(self.w)
        for i in range(len(self.w) - 1, -1, -1):
            w, a = self.w[i], s['layers'][i]
            g = {}
            # x_out = x_mid + down(f)
            g['wdown'] = dense.matmul_t(dx, a['f'], a_mn=True, b_mn=True)
            df = dense.matmul_t(dx, w['wdown'], b_mn=True)
            dgu = train_ops.swiglu_bwd(a['gu'], df)
            g['wgu'] = dense.matmul_t(dgu, a['h2'], a_mn=True, b_mn=True)
            dh2 = dense.matmul_t(dgu, w['wgu'], b_mn=True)
            dx_mid, g['ln_post'] = train_ops.rmsnorm_bwd(a['x_mid'], w['ln_post'], dh2, c.rms_eps, dres=dx)
            # x_mid = x_in + o_proj(attn)
            g['wo'] = dense.matmul_t(dx_mid, a['a'], a_mn=True, b_mn=True)
            da = dense.matmul_t(dx_mid, w['wo'], b_mn=True)
            dqkv = train_ops.attention_bwd(a['qkv'], a['a'], da, a['lse'], B, L, c.n_heads, c.head_dim, True, scale)
            kernels.rope_inplace(dqkv, cos, nsin, L, 2 * c.n_heads, c.head_dim)             # RoPE^T = RoPE(-theta)
            g['wqkv'] = dense.matmul_t(dqkv, a['h1'], a_mn=True, b_mn=True)
            dh1 = dense.matmul_t(dqkv, w['wqkv'], b_mn=True)
            dx, g['ln_in'] = train_ops.rmsnorm_bwd(a['x_in'], w['ln_in'], dh1, c.rms_eps, dres=dx_mid)
            grads[i] = g
            s['layers'][i] = None                                                            # free activations
            if on_layer_grads is not None:
                on_layer_grads(i, g)
        self.grads = dict(layers=grads, top=g_top)
        self.saved = None
        return dx.view(B, L, c.hidden)

    # ------------------------------------------------------------------ optimizer
    def optimizer_step(self, grad_scale=1.0):
        """Fused AdamW on every tensor (fp32 master + moments, refreshes the bf16 compute copy)."""
        self.step_count += 1
        t = self.step_count

        def upd(master, m1, m2, w16, grad, name):
            wd = 0.0 if name.startswith('ln') or name == 'norm' else self.wd
            train_ops.adamw_step(master.view(-1), grad.reshape(-1), m1.view(-1), m2.view(-1), w16.view(-1), self.lr,
                                 self.betas, self.eps, wd, t, grad_scale)
        for i, g in enumerate(self.grads['layers']):
            for k in LAYER_KEYS:
                upd(self.master[i][k], self.m1[i][k], self.m2[i][k], self.w[i][k], g[k], k)
        for k in ('norm', 'lm_head'):
            upd(self.master_top[k], self.m1_top[k], self.m2_top[k], self.w_top[k], self.grads['top'][k], k)
        self.grads = None


class LayerBucketAllReduce:
    """DDP gradient all-reduce, one bucket per decoder layer, launched from the backward hook on a side stream so
    the collective of layer i overlaps the backward of layers < i (NCCL over NVLink; gloo in the CPU tests).
    Sums in place; the 1/world average is folded into AdamW's grad_scale."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.handles = []
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None

    def hook(self, i, grads):
        if not (self.dist.is_available() and self.dist.is_initialized()) or self.dist.get_world_size(self.group) == 1:
            return
        tensors = [grads[k] for k in LAYER_KEYS]
        if self.stream is not None and tensors[0].is_cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                for t in tensors:
                    self.handles.append(self.dist.all_reduce(t, group=self.group, async_op=True))
        else:
            for t in tensors:
                self.handles.append(self.dist.all_reduce(t, group=self.group, async_op=True))

    def reduce_now(self, tensors):
        if self.dist.is_available() and self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1:
            for t in tensors:
                self.handles.append(self.dist.all_reduce(t, group=self.group, async_op=True))

    def wait(self):
        for h in self.handles:
            h.wait()
        self.handles = []
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)


def train_step(stack, inputs_embeds, targets, reducer=None, world_size=1):
    """One optimisation step of the LLaMA stack: forward, backward (+ overlapped gradient all-reduce), AdamW.
    Returns (loss tensor, d_inputs_embeds)."""
    loss = stack.forward(inputs_embeds, targets)
    d_in = stack.backward(on_layer_grads=reducer.hook if reducer is not None else None)
    if reducer is not None:
        reducer.reduce_now([stack.grads['top']['lm_head'], stack.grads['top']['norm']])
        reducer.wait()
    stack.optimizer_step(grad_scale=1.0 / world_size)
    return loss, d_in


class FrontEndTrain:
    """Training-mode front end: the region-token forward up to `inputs_embeds` with the tensors its backward needs,
    and the backward of everything in front of the LLaMA stack that stage 2 trains: splice
    (`spi_llava.py:99-196`), `embed_tokens`, `mm_projector` and the whole SPI module (`SpiTrain`).  The CLIP tower
    is frozen in both stages (train.py:604-612) and is not differentiated.
    head_only=True stops at the pconv output / position-MLP input and returns those hand-over gradients
    (`d_pconv_out`, `d_pos`) instead of running SpiTrain (kept for the slice-level test)."""

    def __init__(self, engine, head_only=False):
        self.eng = engine
        self.head_only = head_only
        self.spi = SpiTrain(engine)
        self.saved = None

    def forward(self, input_ids, images, bboxes):
        from .roi_align import roi_align_mlvl
        from .splice import splice_region_tokens
        eng, c = self.eng, self.eng.cfg
        dev = eng.dev
        input_ids = input_ids.to(dev)
        images = images.to(dev, BF16)
        B, L = input_ids.shape
        plan_b = eng.plan_boxes(bboxes)
        taps = eng.vit(images)
        feat = kernels.cast_tokens_f32_bf16(taps[c.select_index])
        img_rows = dense.linear(feat, eng.proj_w, eng.proj_b).view(B, c.num_patches, c.hidden)
        K = plan_b['K'] if plan_b is not None else 0
        region, pc, t = None, None, None
        if K > 0 and not self.head_only:
            region = (self.spi.forward(taps, plan_b), plan_b['offs'])
        elif K > 0:
            maps, ss = eng.fuse_maps(taps)
            boxes, bidx = plan_b['boxes'], plan_b['bidx']
            rois = torch.cat([bidx[:, None], boxes * float(c.image_size)], 1).contiguous()
            scales = [float(torch.tensor(1.0 / s, dtype=F32)) for s in c.strides]
            feats = roi_align_mlvl(maps, rois, c.roi_out, scales, c.roi_sampling, True, out_dtype=BF16,
                                   gn_scale=[s for s, _ in ss], gn_shift=[b for _, b in ss])
            R = c.roi_out
            pc = dense.conv_nhwc(feats.view(c.num_levels * K, R, R, c.spi_dim), eng.pconv_w, eng.pconv_b,
                                 act='relu', levels=c.num_levels).view(K, -1)
            acc = dense.linear(pc, eng.flat_w, out_dtype=F32)
            pos = kernels.pos_embed_mlp(boxes.contiguous(), *eng.pos)
            t = kernels.add_bias_pos_cast(acc, eng.flat_b, pos)
            region = (dense.linear(t, eng.up_w, eng.up_b), plan_b['offs'])
        elif plan_b is not None:
            region = (torch.zeros((1, c.hidden), dtype=BF16, device=dev), plan_b['offs'])
        embeds, plan = splice_region_tokens(input_ids, eng.embed, img_rows, region, c.num_patches, c.im_patch_token,
                                            c.im_start_token, c.im_end_token, c.bbox_token, return_plan=True)
        self.saved = dict(feat=feat, plan=plan, pc=pc, t=t, K=K, B=B)
        return embeds

    def backward(self, d_embeds):
        """d_embeds [B,L,hidden] bf16 -> dict of gradients under the reference's parameter names."""
        from .splice import splice_backward
        eng, c, s = self.eng, self.eng.cfg, self.saved
        d_image, d_region, d_embed = splice_backward(s['plan'], d_embeds, c.num_patches, s['K'], c.vocab)
        out = {'model.embed_tokens.weight': d_embed}
        _, gw, gb = train_ops.linear_bwd(s['feat'], eng.proj_w, d_image.view(-1, c.hidden), need_dx=False)
        out['model.mm_projector.weight'], out['model.mm_projector.bias'] = gw, gb
        if s['K'] > 0 and not self.head_only:
            out.update(self.spi.backward(d_region))
        elif s['K'] > 0:
            q = 'model.spi_module.roi_align.'
            dt, out[q + 'updims.weight'], out[q + 'updims.bias'] = train_ops.linear_bwd(s['t'], eng.up_w, d_region)
            out[q + 'flatten_linear.bias'] = train_ops.colsum(dt)
            d_pc, gfw, _ = train_ops.linear_bwd(s['pc'], eng.flat_w, dt, has_bias=False)
            # engine layout of flatten_linear.weight is [out, (ph, pw, c)]; the reference's is [out, (c, ph, pw)]
            R, C = c.roi_out, c.spi_dim
            out[q + 'flatten_linear.weight'] = gfw.view(-1, R, R, C).permute(0, 3, 1, 2).reshape(gfw.shape[0], -1)
            out['d_pos'] = dt
            out['d_pconv_out'] = d_pc
        self.saved = None
        return out


class SpiTrain:
    """Forward (with saved activations) and complete backward of the SPI module -- MLVLROIQueryModule
    (gpt4roi/models/layers.py:198-236): pyramid upsampling, MLVLFuseModule (4 input 1x1 convs + 5 shared
    ConvModule(3x3, GN(64), ReLU) rounds with the channel shuffle between levels), MlvlRoIExtractor (4 RoIAligns,
    4 pconvs, flatten_linear, box position MLP, updims) -- on the sm_100a kernels, no autograd.
    Inputs: the CLIP hidden states (frozen tower) and the boxes; output: one region token per box.
    `backward(d_region)` returns gradients under the reference's parameter names and layouts (fp32 or bf16)."""

    def __init__(self, engine):
        self.eng = engine
        self.saved = None

    def forward(self, taps, plan_b):
        from .roi_align import roi_align_mlvl
        eng, c = self.eng, self.eng.cfg
        C, n = c.spi_dim, c.num_levels
        B = next(iter(taps.values())).shape[0]
        ups, maps = [], []
        for l, layer in enumerate(c.level_layers):
            H = c.level_sizes[l]
            up = kernels.upsample_tokens_coords(taps[layer], c.grid, H, eng.spi_cpad)
            ups.append(up)
            maps.append(dense.linear(up.view(-1, eng.spi_cpad), eng.in_w[l], eng.in_b[l]).view(B, H, H, C))
        zs, sts, sss = [maps], [None], [[None] * n]
        for r in range(5):
            new, stats = [], []
            prev, ss = zs[-1], sss[-1]
            for l in range(n):
                top, down = min(l + 1, n - 1), max(l - 1, 0)
                x_in = kernels.fuse_gather(prev[l], prev[top], prev[down], ss[l], ss[top], ss[down])
                st = torch.zeros((B, dense.gn_slots(c.level_sizes[l], c.level_sizes[l]), c.gn_groups, 2), dtype=F32, device=eng.dev)
                new.append(dense.conv_nhwc(x_in, eng.fuse[r]['w'], gn_stats=st))
                stats.append(st)
            zs.append(new)
            sts.append(stats)
            sss.append([kernels.gn_finalize(stats[l], eng.fuse[r]['gamma'], eng.fuse[r]['beta'],
                                            count=c.level_sizes[l] ** 2 * (C // c.gn_groups)) for l in range(n)])
        boxes, bidx = plan_b['boxes'], plan_b['bidx']
        K = boxes.shape[0]
        rois = torch.cat([bidx[:, None], boxes * float(c.image_size)], 1).contiguous()
        scales = [float(torch.tensor(1.0 / s, dtype=F32)) for s in c.strides]
        ss = sss[-1]
        feats = roi_align_mlvl(zs[-1], rois, c.roi_out, scales, c.roi_sampling, True, out_dtype=BF16,
                               gn_scale=[s for s, _ in ss], gn_shift=[b for _, b in ss])
        R = c.roi_out
        pc = dense.conv_nhwc(feats.view(n * K, R, R, C), eng.pconv_w, eng.pconv_b, act='relu', levels=n)
        acc = dense.linear(pc.view(K, -1), eng.flat_w, out_dtype=F32)
        pos = kernels.pos_embed_mlp(boxes.contiguous(), *eng.pos)
        t = kernels.add_bias_pos_cast(acc, eng.flat_b, pos)
        region = dense.linear(t, eng.up_w, eng.up_b)
        self.saved = dict(ups=ups, zs=zs, sts=sts, sss=sss, rois=rois, scales=scales, feats=feats, pc=pc, t=t,
                          boxes=boxes.contiguous(), K=K, B=B)
        return region

    def backward(self, d_region):
        from .roi_align import roi_align_mlvl_backward
        eng, c, s = self.eng, self.eng.cfg, self.saved
        C, n, R, K, B = c.spi_dim, c.num_levels, c.roi_out, s['K'], s['B']
        p = 'model.spi_module.'
        q = p + 'roi_align.'
        g = {}
        # ---- head: updims, + pos, flatten_linear (layers.py:326-335)
        dt, g[q + 'updims.weight'], g[q + 'updims.bias'] = train_ops.linear_bwd(s['t'], eng.up_w, d_region)
        g[q + 'flatten_linear.bias'] = train_ops.colsum(dt)
        for name, v in train_ops.pos_embed_mlp_bwd(s['boxes'], eng.pos, dt).items():
            g[q + 'pos_embedd.' + name] = v
        pcf = s['pc'].view(K, -1)
        d_pc, gfw, _ = train_ops.linear_bwd(pcf, eng.flat_w, dt, has_bias=False)
        g[q + 'flatten_linear.weight'] = gfw.view(-1, R, R, C).permute(0, 3, 1, 2).reshape(gfw.shape[0], -1)
        # ---- relu(sum_l pconv_l(roi_feats_l))  (layers.py:318-325)
        dZ = train_ops.relu_bwd(d_pc, pcf).view(K, R, R, C)
        gb = train_ops.colsum(dZ.view(-1, C))
        feats = s['feats'].view(n, K, R, R, C)
        d_feats = torch.empty_like(feats)
        for l in range(n):
            wl = eng.pconv_w[:, l]                                       # [Cout, 3, 3, Cin], rows n*9*Cin apart
            dx, dW = train_ops.conv3x3_bwd(feats[l], wl, dZ)
            d_feats[l] = dx
            g[q + 'pconvs.%d.weight' % l] = dW.permute(0, 3, 1, 2)
            g[q + 'pconvs.%d.bias' % l] = gb
        # ---- 4 x RoIAlign (the forward fused the last GroupNorm+ReLU into its taps)
        shapes = [tuple(z.shape) for z in s['zs'][-1]]
        dA = roi_align_mlvl_backward(d_feats, s['rois'], shapes, s['scales'], c.roi_sampling, True)
        # ---- MLVLFuseModule, rounds 4 .. 0 (layers.py:152-195)
        for r in range(4, -1, -1):
            gamma = eng.fuse[r]['gamma']
            dg, db = torch.empty(C, dtype=F32, device=eng.dev), torch.empty(C, dtype=F32, device=eng.dev)
            zr, st, ss = s['zs'][r + 1], s['sts'][r + 1], s['sss'][r + 1]
            dz = [train_ops.gn_relu_bwd(zr[l], dA[l], ss[l][0], ss[l][1], st[l], c.level_sizes[l] ** 2 * (C // c.gn_groups),
                                        gamma, dg, db, accumulate=l > 0) for l in range(n)]
            prev, pss = s['zs'][r], s['sss'][r]
            wf = train_ops.conv_weight_flip_t(eng.fuse[r]['w'])
            dW, d_in = None, []
            for l in range(n):
                top, down = min(l + 1, n - 1), max(l - 1, 0)
                x_in = kernels.fuse_gather(prev[l], prev[top], prev[down], pss[l], pss[top], pss[down])
                dx, dW = train_ops.conv3x3_bwd(x_in, eng.fuse[r]['w'], dz[l], dw_acc=dW, wf=wf)
                d_in.append(dx)
            g[p + 'mlvl_fuse.fuse_convs.%d.conv.weight' % r] = dW.permute(0, 3, 1, 2)
            g[p + 'mlvl_fuse.fuse_convs.%d.gn.weight' % r] = dg
            g[p + 'mlvl_fuse.fuse_convs.%d.gn.bias' % r] = db
            dA = [train_ops.fuse_gather_bwd(d_in, m) for m in range(n)]
            s['zs'][r + 1] = None
        # ---- input 1x1 convs on [tokens | coords] (layers.py:185-191); the CLIP tower is frozen: no grad_x
        for l in range(n):
            dm = train_ops.cast_f32_bf16(dA[l].view(-1, C))
            up = s['ups'][l].view(-1, eng.spi_cpad)
            gw = dense.matmul_t(dm, up, a_mn=True, b_mn=True)            # [C, cpad]
            g[p + 'mlvl_fuse.input_conv.%d.weight' % l] = gw[:, :C + 2].reshape(C, C + 2, 1, 1)
            g[p + 'mlvl_fuse.input_conv.%d.bias' % l] = train_ops.colsum(dm)
        self.saved = None
        return g


class Stage2Trainer:
    """One optimisation step of GPT4RoI stage 2 (scripts/train_stage2.sh -> gpt4roi/train/train.py:698-712): every
    parameter except the frozen CLIP tower is trained -- embed_tokens, mm_projector, the SPI module, the 32 LLaMA
    layers, final norm and lm_head -- with the loss of llava/model/llava.py:238-249, DDP gradient all-reduce and
    torch.optim.AdamW semantics.  Forward, backward and optimizer run on the sm_100a kernels; no autograd.

        front = FrontEndTrain(PrefillEngine without LLaMA layers)   ViT (frozen) -> projector / SPI -> splice
        stack = LlamaTrainStack                                     decoder stack -> lm_head -> cross entropy
    """

    def __init__(self, cfg, state_dict, vit_state_dict, device, lr=2e-5, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0, reducer=None, world_size=1):
        import copy
        from .engine import PrefillEngine
        self.cfg, self.dev = cfg, torch.device(device)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.reducer, self.world = reducer, world_size
        fcfg = copy.copy(cfg)
        fcfg.n_layers = 0                                   # the front-end engine holds no decoder layers
        self.eng = PrefillEngine(fcfg, state_dict, vit_state_dict, device)
        self.front = FrontEndTrain(self.eng)
        self.stack = LlamaTrainStack(cfg, state_dict, device, lr, betas, eps, weight_decay)
        names = [k for k in state_dict if k.startswith('model.spi_module.') or k.startswith('model.mm_projector.')
                 or k == 'model.embed_tokens.weight']
        self.master = {k: state_dict[k].detach().to(self.dev, F32).contiguous() for k in names}
        self.m1 = {k: torch.zeros_like(v) for k, v in self.master.items()}
        self.m2 = {k: torch.zeros_like(v) for k, v in self.master.items()}
        self.w16 = {k: v.to(BF16) for k, v in self.master.items()}
        self.front_grads = None

    @staticmethod
    def shift_labels(labels):
        """labels [B,L] (-100 = ignored) -> targets with targets[:, t] = labels[:, t+1] (llava.py:241-242)."""
        t = torch.full_like(labels, -100)
        t[:, :-1] = labels[:, 1:]
        return t

    def forward_backward(self, input_ids, images, bboxes, labels):
        """Returns the loss; gradients are left in self.stack.grads and self.front_grads (already all-reduced)."""
        embeds = self.front.forward(input_ids, images, bboxes)
        loss = self.stack.forward(embeds, self.shift_labels(labels.to(self.dev)))
        d_embeds = self.stack.backward(on_layer_grads=self.reducer.hook if self.reducer is not None else None)
        self.front_grads = {k: v for k, v in self.front.backward(d_embeds).items() if k in self.master}
        if self.reducer is not None:
            top = self.stack.grads['top']
            # every rank must contribute the same tensors to the collective: a rank whose micro-batch has no boxes
            # (no SPI gradients) contributes zeros, in the master's key order
            self.front_grads = {k: (self.front_grads[k].reshape(self.master[k].shape).contiguous() if k in self.front_grads
                                    else torch.zeros(self.master[k].shape, dtype=F32, device=self.dev)) for k in self.master}
            self.front_grads = {k: (v if v.dtype == F32 else v.float()) for k, v in self.front_grads.items()}
            self.reducer.reduce_now([top['lm_head'], top['norm']] + list(self.front_grads.values()))
            self.reducer.wait()
        return loss

    def optimizer_step(self):
        scale = 1.0 / self.world
        self.stack.optimizer_step(grad_scale=scale)
        t = self.stack.step_count
        for k, gr in self.front_grads.items():
            no_decay = k.endswith('.bias') or '.gn.' in k or 'pos_embedd.2.' in k or 'pos_embedd.5.' in k
            gr = gr.reshape(-1)
            if gr.dtype not in (BF16, F32):
                gr = gr.float()
            train_ops.adamw_step(self.master[k].view(-1), gr, self.m1[k].view(-1), self.m2[k].view(-1),
                                 self.w16[k].view(-1), self.lr, self.betas, self.eps, 0.0 if no_decay else self.wd, t, scale)
        self.front_grads = None
        # refresh the engine's bf16 tensors (its own layouts: padded 1x1 weights, KHWC convs, stacked pconvs ...)
        self.eng._prepare_spi(self.w16)
        self.eng.embed = self.w16['model.embed_tokens.weight']

    def step(self, input_ids, images, bboxes, labels):
        loss = self.forward_backward(input_ids, images, bboxes, labels)
        self.optimizer_step()
        return loss

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        """All trained parameters, fp32, reference names and layouts -- what the reference's trainer saves
        (train.py:88-98); the CLIP tower is not part of the model's state dict (llava.py:47-48)."""
        out = dict(self.master)
        out.update(self.stack.state_dict())
        return out

    def save_pretrained(self, out_dir, dtype=None, max_shard_bytes=10 * 1024 ** 3):
        return save_checkpoint(self.state_dict(), out_dir, max_shard_bytes, dtype)

    def optimizer_state(self):
        """AdamW step count and moments under reference names (resume: pass the saved weights to __init__, then
        load_optimizer_state)."""
        m1, m2 = dict(self.m1), dict(self.m2)
        for i in range(len(self.stack.master)):
            m1.update(unfuse_llama_layer(self.stack.m1[i], i))
            m2.update(unfuse_llama_layer(self.stack.m2[i], i))
        for k, name in (('norm', 'model.norm.weight'), ('lm_head', 'lm_head.weight')):
            m1[name], m2[name] = self.stack.m1_top[k], self.stack.m2_top[k]
        return dict(step=self.stack.step_count, exp_avg=m1, exp_avg_sq=m2)

    def load_optimizer_state(self, state):
        self.stack.step_count = int(state['step'])
        for src, dst_front, dst_layers, dst_top in ((state['exp_avg'], self.m1, self.stack.m1, self.stack.m1_top),
                                                    (state['exp_avg_sq'], self.m2, self.stack.m2, self.stack.m2_top)):
            for k in dst_front:
                dst_front[k].copy_(src[k])
            for i in range(len(dst_layers)):
                fused = fuse_llama_layer(src, i, self.dev)
                for k in LAYER_KEYS:
                    dst_layers[i][k].copy_(fused[k])
            dst_top['norm'].copy_(src['model.norm.weight'])
            dst_top['lm_head'].copy_(src['lm_head.weight'])
