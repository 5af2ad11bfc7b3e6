from nerf_slam_b200.pipeline import SlamModule  # noqa: F401
