from nerf_slam_b200.pipeline import VioSLAM, WORLD_T_IMU_T0  # noqa: F401
