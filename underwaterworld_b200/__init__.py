"""underwaterworld_b200 -- B200-native chunk builder for UnderwaterWorld's hot path.

Perlin density sampling -> blank early-out -> marching-cubes classify -> deduplicated vertex
and index emission, as hand-written sm_100a CUDA kernels behind a C ABI (include/uwcuda.h).
This package is the thin host-side mirror of the reference's `Chunk` interface.
"""
from .chunk import (Batch, Chunk, ChunkBuilder, ChunkMesh, Perlin, build_chunks, CHUNK_SIZE, INTERNAL_SIZE,
                    ISO_LEVEL, PERLIN_OCTAVES)
from ._ffi import UwError, load_library
from . import region
from . import gather
from .gather import MultiBuilder, RegionGather

__all__ = ["Batch", "Chunk", "ChunkBuilder", "ChunkMesh", "Perlin", "build_chunks", "UwError", "load_library",
           "region", "gather", "MultiBuilder", "RegionGather", "CHUNK_SIZE", "INTERNAL_SIZE", "ISO_LEVEL", "PERLIN_OCTAVES"]
