"""CPU test: the model-seam mirror exposes exactly the reference's SPI parameter names/shapes
(probed from the reference's own MLVLROIQueryModule)."""
import torch

APPENDIX_C = {
    **{'mlvl_fuse.input_conv.%d.weight' % i: (1024, 1026, 1, 1) for i in range(4)},
    **{'mlvl_fuse.input_conv.%d.bias' % i: (1024,) for i in range(4)},
    **{'mlvl_fuse.fuse_convs.%d.conv.weight' % i: (1024, 1024, 3, 3) for i in range(5)},
    **{'mlvl_fuse.fuse_convs.%d.gn.weight' % i: (1024,) for i in range(5)},
    **{'mlvl_fuse.fuse_convs.%d.gn.bias' % i: (1024,) for i in range(5)},
    **{'roi_align.pconvs.%d.weight' % i: (1024, 1024, 3, 3) for i in range(4)},
    **{'roi_align.pconvs.%d.bias' % i: (1024,) for i in range(4)},
    'roi_align.pos_embedd.0.weight': (256, 4), 'roi_align.pos_embedd.0.bias': (256,),
    'roi_align.pos_embedd.2.weight': (256,), 'roi_align.pos_embedd.2.bias': (256,),
    'roi_align.pos_embedd.3.weight': (1024, 256), 'roi_align.pos_embedd.3.bias': (1024,),
    'roi_align.pos_embedd.5.weight': (1024,), 'roi_align.pos_embedd.5.bias': (1024,),
    'roi_align.updims.weight': (4096, 1024), 'roi_align.updims.bias': (4096,),
    'roi_align.flatten_linear.weight': (1024, 200704), 'roi_align.flatten_linear.bias': (1024,),
}


def test_spi_module_state_dict_matches_reference_layout():
    from gpt4roi_b200.spi_llava import MLVLROIQueryModule
    with torch.device('meta'):
        m = MLVLROIQueryModule(embed_dims=1024, out_dims=4096, num_levels=4)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == APPENDIX_C
    assert len(got) == 43
    r = repr(m.roi_align.roi_layers[0])
    assert r == ('RoIAlign(output_size=(14, 14), spatial_scale=%s, sampling_ratio=2, pool_mode=avg, '
                 'aligned=True, use_torchvision=False)' % (1 / (14 / 8)))


def test_random_state_dict_uses_reference_names():
    from gpt4roi_b200.engine import EngineConfig, random_state_dicts
    cfg = EngineConfig(n_layers=0, vit_layers=0)
    with torch.device('meta'):
        pass
    sd, _ = random_state_dicts(cfg, 'cpu', dtype=torch.bfloat16)
    spi = {k[len('model.spi_module.'):]: tuple(v.shape) for k, v in sd.items() if k.startswith('model.spi_module.')}
    assert spi == APPENDIX_C
