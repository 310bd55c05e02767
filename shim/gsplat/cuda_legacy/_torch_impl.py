from bilateral_driving_b200.render import quat_to_rotmat  # noqa: F401
