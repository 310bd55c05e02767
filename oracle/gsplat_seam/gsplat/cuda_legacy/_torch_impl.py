from oracle.raster_ref import quat_to_rotmat  # noqa: F401
