"""A ``gsplat`` package for the TESTS (test infrastructure only, never on the product path): put
``<repo>/oracle/gsplat_seam`` ahead on ``sys.path`` and the reference's ``models/gaussians/basics.py:12-15`` imports
resolve here.  Calls with CPU tensors run the CPU oracle (``oracle/raster_ref.py`` / ``oracle/sh_ref.py``); calls
with CUDA tensors are handed to the product (``bilateral_driving_b200.render``) exactly as ``<repo>/shim`` would.
That lets ONE process hold the reference's unmodified trainer on the CPU (the checker) next to the drop-in trainer
on the GPU (the thing checked): tests/test_trainer_reference.py."""
__version__ = "1.3.0+bds_test_seam"
