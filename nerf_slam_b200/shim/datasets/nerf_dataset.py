from nerf_slam_b200.datasets import NeRFDataset  # noqa: F401
