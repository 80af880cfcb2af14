"""`nvdiffrast.torch` -> fmhr_b200.dr (B200-native rasterize / interpolate / antialias)."""
from fmhr_b200.dr import (RasterizeCudaContext, RasterizeGLContext, antialias, get_antialias_topology_hash,  # noqa: F401
                          interpolate, rasterize)

antialias_construct_topology_hash = get_antialias_topology_hash
