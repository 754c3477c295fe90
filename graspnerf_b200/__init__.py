"""graspnerf_b200: Blackwell-native (sm_100a) implementation of GraspNeRF's
generalizable-NeRF volumetric TSDF hot path (reference: src/nr/network)."""
__version__ = "0.1.0"
