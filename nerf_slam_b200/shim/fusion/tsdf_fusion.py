from nerf_slam_b200.tsdf_fusion import TsdfFusion  # noqa: F401
