from nerf_slam_b200.nerf_fusion import NerfFusion  # noqa: F401
