"""Network-relevant keys of the shipped configuration (src/nr/configs/nrvgn_sdf.yaml:5-31), as a dict the mirror's
constructors accept; train.sh / sim_grasp.py load the YAML itself - this copy serves bench.py, tools and tests."""

NRVGN_SDF_CFG = {
    'network': 'grasp_nerf',
    'init_net_type': 'cost_volume',
    'agg_net_type': 'neus',
    'use_hierarchical_sampling': True,
    'use_depth_loss': True,
    'dist_decoder_cfg': {'use_vis': False},
    'fine_dist_decoder_cfg': {'use_vis': False},
    'ray_batch_num': 4096,
    'sample_volume': True,
    'render_rgb': True,
    'volume_type': ['sdf'],
    'volume_resolution': 40,
    'depth_sample_num': 40,
    'fine_depth_sample_num': 40,
    'agg_net_cfg': {'sample_num': 40, 'init_s': 0.3, 'fix_s': 0},
    'fine_agg_net_cfg': {'sample_num': 40, 'init_s': 0.3, 'fix_s': 0},
    'render_depth': True,
}
