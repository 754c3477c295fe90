"""graspnerf_b200: Blackwell-native (sm_100a) implementation of GraspNeRF's volumetric TSDF hot path."""
__version__ = '0.2.0'


def install(verbose=False):
    """Drop-in switch for an UNMODIFIED reference checkout: with the reference's `src/nr` (and `src`) on sys.path, replaces
    the entry of its model registry - `network.renderer.name2network['grasp_nerf']` (renderer.py:333-335), the only place
    Trainer (trainer.py:61), ValidationEvaluator and GraspNeRFPlanner (main.py:152) look the model up - by this package's
    mirror class, and the module-level class names with it.  Same ctor cfg, forward(data), output keys and state_dict keys,
    so train.sh / scripts/sim_grasp.py run with zero source edits:

        python -m graspnerf_b200 src/nr/run_training.py --cfg src/nr/configs/nrvgn_sdf.yaml

    Returns the patched registry."""
    import importlib
    ref = importlib.import_module('network.renderer')            # the reference's module (needs src/nr on sys.path)
    from .network import renderer as mirror
    if getattr(ref, '__graspnerf_b200__', False):
        return ref.name2network
    ref.name2network['grasp_nerf'] = mirror.GraspNeRF
    ref.GraspNeRF, ref.NeuralRayRenderer = mirror.GraspNeRF, mirror.NeuralRayRenderer
    ref.__graspnerf_b200__ = True
    if verbose:
        print('[graspnerf_b200] network.renderer.name2network["grasp_nerf"] -> graspnerf_b200.network.GraspNeRF')
    return ref.name2network
