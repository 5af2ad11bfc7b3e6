from nerf_slam_b200.pipeline import MIMOPipelineModule  # noqa: F401
