"""Synthetic cases behind the committed golden fixtures (kwargs of graspnerf_b200.synth)."""

# name -> synth kwargs.  cfg1 = BASELINE.json configs[0]/[1]; the others stress masks.
VOLUME_CASES = {
    'cfg1': dict(seed=0, num_views=6, h=288, w=512),
    'small_v4': dict(seed=3, num_views=4, h=96, w=160, radius=0.45),
    'close_v3': dict(seed=5, num_views=3, h=64, w=96, radius=0.30, theta=1.2),
}
RENDER_CASES = {
    'rays64': dict(scene=dict(seed=0, num_views=6, h=288, w=512), num_rays=64, qseed=0),
    'rays48_small': dict(scene=dict(seed=3, num_views=4, h=96, w=160, radius=0.45), num_rays=48, qseed=7),
}
