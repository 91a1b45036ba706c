"""tssplat_b200 -- B200-native geometry-energy hot path of TetSphere Splatting (gmh14/tssplat).

Modules (nothing is imported eagerly; the CUDA library is loaded by ``_capi`` on first use and there
is no CPU fallback):

* ``tet_spheres_ext`` -- drop-in for the reference's pybind11 module (also importable as
  ``from tet_spheres import tet_spheres_ext``)
* ``energies``        -- ``SmoothnessBarrierFunc`` / ``SmoothnessBarrierEnergy`` mirror
* ``optimizer``       -- ``AdamUniform`` drop-in over the C ABI
* ``sharding``        -- sphere-per-rank partition + async scalar all-reduce
* ``mesh``            -- ``.veg`` IO and seeded synthetic tet-sphere packs
* ``build``           -- in-tree nvcc build of ``libtssplat_b200.so`` (sm_100a)
"""
__version__ = "0.1.0"
