from nerf_slam_b200.pipeline import SLAM  # noqa: F401
