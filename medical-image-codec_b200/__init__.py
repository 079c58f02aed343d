"""micgpu -- B200-native (sm_100a) codec path for MIC (pappuks/medical-image-codec).

The product is libmicgpu.so (hand-written CUDA kernels behind a C ABI, see
include/micgpu.h).  This package is the thin host-side mirror of the
reference's Go API for that path, used by the tests and bench.py; it never
touches oracle/ and has no CPU fallback: importing `api` fails loudly when the
native library is missing.
"""
from . import api  # noqa: F401
from .api import (  # noqa: F401
    CompressMultiFrame,
    CompressParallelStrips,
    CompressRGB,
    CompressWSI,
    WaveletV2CompressBatch,
    WaveletV2RLEFSECompressU16,
    WaveletV2SIMDRLEFSECompressU16,
    CompressParallelStripsBatch,
    CompressParallelStripsAdaptive,
    CompressSingleFrameGrad,
    DecompressParallelStripsAdaptive,
    DecompressSingleFrameGrad,
    AdaptiveStripBoundaries,
    CompressSingleFrame,
    CompressSingleFrame4State,
    CompressSingleFrame8State,
    DeltaRleCompress,
    RleCompress,
    Decoder,
    DecompressFrame,
    DecompressMultiFrame,
    DecompressParallelStrips,
    DecompressParallelStripsBatch,
    DecompressRGB,
    DecompressWSIRegion,
    DecompressWSITile,
    DecompressWSITiles,
    DecompressWSITileRange,
    WsiPlan,
    WsiSlide,
    Init,
    WriteMIC1,
    WriteMICR,
    DecodeMicFile,
    ReadWSIHeader,
    WaveletV2DecompressBatch,
    WaveletV2RLEFSEDecompressU16,
    WaveletV2SIMDRLEFSEDecompressU16,
    DecompressSingleFrame,
    MicGpuError,
    temporal_add_carry,
    temporal_add_carry_peers,
    device_alloc,
    device_free,
    ipc_export,
    ipc_open,
    ipc_close,
    lib,
)
