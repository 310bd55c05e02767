from bilateral_driving_b200.render import num_sh_bases  # noqa: F401
