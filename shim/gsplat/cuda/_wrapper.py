from bilateral_driving_b200.render import spherical_harmonics  # noqa: F401
