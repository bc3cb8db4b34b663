"""Model seam: drop-in mirrors of the reference's model classes, running on the sm_100a engine.

Same class names, constructor arguments, forward keyword arguments, attributes and -- above all --
parameter names/shapes as the reference, so its checkpoints load unchanged and its train / serve
scripts can import these classes instead:

    gpt4roi/models/layers.py:96-195    MLVLFuseModule
    gpt4roi/models/layers.py:198-236   MLVLROIQueryModule
    gpt4roi/models/layers.py:239-335   MlvlRoIExtractor
    gpt4roi/models/spi_llava.py:15-205 SPILlavaLlamaModel
    gpt4roi/models/spi_llava.py:215-306 SPILlavaMPTForCausalLM
    llava/model/llava.py:36-40         LlavaConfig

The modules below are PARAMETER CONTAINERS with the reference's state-dict layout (SURVEY.md
Appendix C; mmcv's ConvModule names `.conv` / `.gn` included).  Their `forward` does not run
PyTorch ops: it hands the weights to `engine.PrefillEngine` (re-laid-out once, cached until the
weights change) which launches the hand-written kernels.  Round-1 limits, stated: inference only
(prefill `forward` + `generate()` with its own KV-cache decode loop; HF `past_key_values` plumbing and
backward through the dense blocks are not provided), right-padded attention masks only; anything else
raises NotImplementedError instead of falling back.
"""
from typing import List, Optional

import torch
import torch.nn as nn
from transformers import LlamaConfig, LlamaForCausalLM, LlamaModel
from transformers.modeling_outputs import CausalLMOutputWithPast

from .engine import EngineConfig, PrefillEngine
from .roi_align import RoIAlign

DEFAULT_IMAGE_PATCH_TOKEN = '<im_patch>'
DEFAULT_IM_START_TOKEN = '<im_start>'
DEFAULT_IM_END_TOKEN = '<im_end>'


class LlavaConfig(LlamaConfig):
    model_type = 'llava_gpt4roi_b200'  # the reference's "llava" collides with newer transformers (llava.py:329)


class _ConvModule(nn.Module):
    """Parameter layout of mmcv.cnn.ConvModule(conv -> GN -> ReLU): `.conv.weight`, `.gn.{weight,bias}`
    (mmcv-1.4.7/mmcv/cnn/bricks/conv_module.py:70-208; conv has no bias when a norm follows)."""

    def __init__(self, cin, cout, groups=64):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, stride=1, padding=1, bias=False)
        self.gn = nn.GroupNorm(groups, cout)


class MLVLFuseModule(nn.Module):
    def __init__(self, input_dims=1024, embed_dims=1024, num_levels=3, num_fuse=4):
        super().__init__()
        self.embed_dims, self.num_levels, self.num_fuse, self.input_dims = embed_dims, num_levels, num_fuse, input_dims
        self.shuffle_channles = embed_dims // 4
        self.remain_chs = embed_dims - 2 * self.shuffle_channles
        self.fuse_lvl_list = [(l, min(l + 1, num_levels - 1), max(l - 1, 0)) for l in range(num_levels)]
        self.input_conv = nn.ModuleList([nn.Conv2d(input_dims + 2, embed_dims, 1) for _ in range(num_levels)])
        self.fuse_convs = nn.ModuleList([_ConvModule(embed_dims, embed_dims) for _ in range(num_fuse)])
        self.init_weights()

    def init_weights(self):  # layers.py:146-150
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, 0, 0.01)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)


class MlvlRoIExtractor(nn.Module):
    def __init__(self, roi_layer, out_channels, featmap_strides, embed_dims=1024, stride=1, norm_init=True,
                 fuse_level=3, finest_scale=56, init_cfg=None):
        super().__init__()
        cfg = dict(roi_layer)
        assert cfg.pop('type') == 'RoIAlign'
        # mmdet BaseRoIExtractor.build_roi_layers (base_roi_extractor.py:37-60): one layer per stride
        self.roi_layers = nn.ModuleList([RoIAlign(spatial_scale=1 / s, **cfg) for s in featmap_strides])
        self.out_channels, self.featmap_strides = out_channels, featmap_strides
        self.embed_dims, self.finest_scale, self.fuse_level, self.norm_init = embed_dims, finest_scale, fuse_level, norm_init
        self.pconvs = nn.ModuleList(nn.Conv2d(embed_dims, embed_dims, 3, stride=1, padding=1) for _ in range(fuse_level))
        self.pos_embedd = nn.Sequential(nn.Linear(4, 256), nn.ReLU(inplace=True), nn.LayerNorm(256),
                                        nn.Linear(256, 1024), nn.ReLU(inplace=True), nn.LayerNorm(1024))
        self.updims = nn.Linear(1024, 4096)
        self.flatten_linear = nn.Linear(embed_dims * self.roi_layers[0].output_size[0] ** 2, 1024)
        for m in self.modules():  # layers.py:275-278
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.constant_(m.bias, 0)

    @property
    def num_inputs(self):
        return len(self.featmap_strides)


class MLVLROIQueryModule(nn.Module):
    def __init__(self, embed_dims=1024, out_dims=4096, num_levels=3):
        super().__init__()
        self.mlvl_fuse = MLVLFuseModule(input_dims=embed_dims, embed_dims=embed_dims, num_levels=num_levels, num_fuse=5)
        strids = [14 / 8, 14 / 4, 14 / 2, 14]
        assert len(strids) == num_levels
        self.roi_align = MlvlRoIExtractor(roi_layer=dict(type='RoIAlign', output_size=14, sampling_ratio=2),
                                          out_channels=embed_dims, embed_dims=embed_dims, fuse_level=num_levels,
                                          featmap_strides=strids)


class SPILlavaLlamaModel(LlamaModel):
    config_class = LlavaConfig

    def __init__(self, config):
        super().__init__(config)
        if hasattr(config, 'mm_vision_tower') and getattr(config, 'mm_vision_tower_instance', None) is not None:
            self.vision_tower = [config.mm_vision_tower_instance]  # python list: HACK for FSDP (llava.py:47-48)
        if getattr(config, 'use_mm_proj', True):
            self.mm_projector = nn.Linear(getattr(config, 'mm_hidden_size', 1024), config.hidden_size)
        self.num_level_spi_features = 4
        self.spi_module = MLVLROIQueryModule(embed_dims=1024, out_dims=4096, num_levels=4)


class SPILlavaMPTForCausalLM(LlamaForCausalLM):
    """Drop-in for gpt4roi.models.spi_llava.SPILlavaMPTForCausalLM (spi_llava.py:215-306)."""
    config_class = LlavaConfig

    def __init__(self, config):
        super(LlamaForCausalLM, self).__init__(config)
        self.model = SPILlavaLlamaModel(config)
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.post_init()
        self._engine = None
        self._engine_key = None

    def get_model(self):
        return self.model

    # ---- engine plumbing ---------------------------------------------------------------
    def set_vision_tower(self, clip_vision_model, image_size=None):
        """Attach the frozen CLIP tower (the reference keeps it in a python list outside the state dict)."""
        self.model.vision_tower = [clip_vision_model]
        if image_size is not None:
            self._image_size = image_size
        self._engine = None

    def invalidate_engine(self):
        """Call after changing parameters in place (the engine holds re-laid-out copies)."""
        self._engine = None

    def _get_engine(self, device):
        vt = getattr(self.model, 'vision_tower', None)
        if vt is None:
            raise RuntimeError('vision tower not attached (model.model.vision_tower is a list of one CLIPVisionModel)')
        vt = vt[0]
        vc = vt.config
        key = (str(device), id(vt))
        if self._engine is not None and self._engine_key == key:
            return self._engine
        tk = vc
        cfg = EngineConfig(image_size=vc.image_size, patch_size=vc.patch_size, vit_hidden=vc.hidden_size,
                           vit_heads=vc.num_attention_heads, vit_layers=vc.num_hidden_layers,
                           vit_mlp=vc.intermediate_size, vit_eps=vc.layer_norm_eps,
                           select_layer=getattr(self.config, 'mm_vision_select_layer', -1),
                           hidden=self.config.hidden_size, n_heads=self.config.num_attention_heads,
                           n_layers=self.config.num_hidden_layers, mlp=self.config.intermediate_size,
                           vocab=self.config.vocab_size, rms_eps=self.config.rms_norm_eps,
                           im_patch_token=getattr(tk, 'im_patch_token', -1), bbox_token=getattr(tk, 'bbox_token', -2),
                           im_start_token=getattr(tk, 'im_start_token', -3), im_end_token=getattr(tk, 'im_end_token', -4))
        if not getattr(tk, 'use_im_start_end', True):
            raise NotImplementedError('use_im_start_end=False branch (spi_llava.py:163-194) is not on the GPT4RoI path')
        self._engine = PrefillEngine(cfg, self.state_dict(), vt.state_dict(), device)
        self._engine_key = key
        return self._engine

    # ---- reference-facing forward (spi_llava.py:226-240 + llava.py:203-261) -------------
    def forward(self, input_ids: torch.LongTensor = None, attention_mask: Optional[torch.Tensor] = None,
                past_key_values: Optional[List[torch.FloatTensor]] = None,
                inputs_embeds: Optional[torch.FloatTensor] = None, labels: Optional[torch.LongTensor] = None,
                use_cache: Optional[bool] = None, output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None, images: Optional[torch.FloatTensor] = None,
                return_dict: Optional[bool] = None, img_metas=None, bboxes=None, **kwargs):
        if inputs_embeds is not None or past_key_values is not None or output_attentions or output_hidden_states:
            raise NotImplementedError('gpt4roi_b200 round 1 implements the prefill forward from input_ids only '
                                      '(no KV-cache decode step / inputs_embeds / attention maps yet)')
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError('training (backward through the dense blocks) is a next-round item')
        if images is None:
            raise NotImplementedError('text-only forward: use the stock LlamaForCausalLM path')
        if type(images) is list:
            raise NotImplementedError('list-of-images input is undefined in the reference SPI branch (spi_llava.py:52-64)')
        eng = self._get_engine(input_ids.device)
        logits = eng.forward(input_ids, images, bboxes, attention_mask=attention_mask)
        loss = None
        if labels is not None:  # llava.py:238-249
            shift_logits = logits[..., :-1, :].float().reshape(-1, self.config.vocab_size)
            shift_labels = labels[..., 1:].reshape(-1).to(shift_logits.device)
            loss = nn.functional.cross_entropy(shift_logits, shift_labels)
        if return_dict is False:
            return (loss, logits) if loss is not None else (logits,)
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=None, hidden_states=None, attentions=None)

    @torch.no_grad()
    def generate(self, input_ids=None, images=None, bboxes=None, max_new_tokens=32, do_sample=False,
                 temperature=1.0, stopping_criteria=None, eos_token_id=None, **kwargs):
        """Region-token prefill + KV-cache decode on the sm_100a engine.  Accepts the arguments the demo
        passes (gpt4roi/app.py:293-300); boxes may be given here or bound the way app.py does
        (`self.forward = partial(self.forward, bboxes=...)`, :286-291).  Returns ids [B, L+new] like HF."""
        if bboxes is None:
            bboxes = getattr(getattr(self, 'forward', None), 'keywords', {}).get('bboxes')
        eng = self._get_engine(input_ids.device)
        return eng.generate(input_ids, images, bboxes, max_new_tokens=max_new_tokens, do_sample=do_sample,
                            temperature=temperature, stopping_criteria=stopping_criteria, eos_token_id=eos_token_id)

    def initialize_vision_tokenizer(self, mm_use_im_start_end, tokenizer, device, tune_mm_mlp_adapter=False,
                                    pretrain_mm_mlp_adapter=None):
        """spi_llava.py:242-306 (token bookkeeping only; embedding resize is plain torch on parameters)."""
        vision_config = self.get_model().vision_tower[0].config
        vision_config.use_im_start_end = mm_use_im_start_end
        tokenizer.add_tokens([DEFAULT_IMAGE_PATCH_TOKEN], special_tokens=True)
        self.resize_token_embeddings(len(tokenizer))
        num_spi_tokens = tokenizer.add_tokens(['<bbox>', '<point>'], special_tokens=True)
        if mm_use_im_start_end:
            num_new_tokens = tokenizer.add_tokens([DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN], special_tokens=True)
            self.resize_token_embeddings(len(tokenizer))
            vision_config.im_start_token, vision_config.im_end_token = tokenizer.convert_tokens_to_ids(
                [DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN])
            num_new_tokens = num_new_tokens + num_spi_tokens
            if num_new_tokens > 0:
                inp = self.get_input_embeddings().weight.data
                out = self.get_output_embeddings().weight.data
                inp[-num_new_tokens:] = inp[:-num_new_tokens].mean(dim=0, keepdim=True)
                out[-num_new_tokens:] = out[:-num_new_tokens].mean(dim=0, keepdim=True)
        vision_config.im_patch_token = tokenizer.convert_tokens_to_ids([DEFAULT_IMAGE_PATCH_TOKEN])[0]
        vision_config.bbox_token = tokenizer.convert_tokens_to_ids(['<bbox>'])[0]
        vision_config.point_token = tokenizer.convert_tokens_to_ids(['<point>'])[0]
        for m in self.modules():
            m.tokenizer = tokenizer
        self.invalidate_engine()
