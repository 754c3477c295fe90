"""Host-side checks that need no GPU: configuration values the kernels hard-code are refused by the mirror's constructor
(instead of silently computing something else), and the package's seed-0 weights equal the committed reference weights."""
import numpy as np
import pytest

from tests.helpers import load_golden
from tests.test_boundary import CFG


@pytest.mark.parametrize('patch,match', [
    ({'dist_decoder_cfg': {}}, 'use_vis'),                              # reference default use_vis=True (dist_decoder.py:54-58)
    ({'fine_dist_decoder_cfg': {'use_vis': True}}, 'use_vis'),
    ({'dist_decoder_cfg': {'use_vis': False, 'bias_val': 0.1}}, 'bias_val'),
    ({'disable_view_dir': True}, 'disable_view_dir'),
    ({'fine_depth_use_all': True}, 'fine_depth_use_all'),
    ({'alpha_value_ground_state': -10}, 'alpha_value_ground_state'),
    ({'agg_net_type': 'default'}, 'agg_net_type'),
])
def test_unsupported_configuration_is_refused(patch, match):
    from graspnerf_b200.network import name2network
    cfg = {**CFG, **patch}
    with pytest.raises(NotImplementedError, match=match):
        name2network[cfg['network']](cfg)


def test_shipped_configuration_is_accepted_and_seed0_weights_match_the_reference():
    from graspnerf_b200.weights import seed0_weights
    sd = seed0_weights()
    g = load_golden('weights_seed0.npz')
    assert set(sd) == set(g)
    assert all(np.array_equal(sd[k].numpy(), g[k]) for k in g)
