"""scone_b200: B200-native event-based Monte Carlo transport engine behind SCONE's interfaces.

The product is the CUDA library scone_b200/libscone_b200.so (C ABI in include/scone_b200.h).
This package is the thin Python binding used by bench.py and the tests; there is no CPU fallback:
creating an engine without the built library or without a CUDA device raises.
"""
from .lib import load_library, library_path, EngineError  # noqa: F401
from .physics_package import EigenPhysicsPackage, FixedSourcePhysicsPackage, GeometryHandle, CycleResult  # noqa: F401
from . import distributed  # noqa: F401,E402
