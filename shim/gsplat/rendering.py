from bilateral_driving_b200.render import rasterization  # noqa: F401
