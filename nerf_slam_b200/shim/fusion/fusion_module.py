from nerf_slam_b200.pipeline import FusionModule  # noqa: F401
