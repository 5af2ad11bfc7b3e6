from nerf_slam_b200.pipeline import DataModule  # noqa: F401
