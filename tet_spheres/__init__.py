"""``from tet_spheres import tet_spheres_ext`` -- the import the reference's
``energies/smooth_barrier.py:6`` and ``tssplat_ext/test_ext.py:3`` perform, served by the
B200-native implementation (``tssplat_b200.tet_spheres_ext``)."""
from tssplat_b200 import tet_spheres_ext  # noqa: F401

__all__ = ["tet_spheres_ext"]
