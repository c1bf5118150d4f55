"""B200-native hot path of mapr-demos/monte-carlo-portfolios.

Everything that computes lives in libmcp.so (CUDA, sm_100a, C ABI in include/mcp.h);
this package is the thin host-side mirror of the reference's Java interfaces for that
path.  Import as `import mcp_b200` (see /mcp_b200.py).
"""
from . import _native as native
from ._native import (CODEC_BP32_VB, CODEC_FASTPFOR128_VB, CODEC_FASTPFOR256_VB, CODEC_IBP32_IVB, CODEC_SK_BP32_VB,
                      CODEC_SK_FASTPFOR128_VB, CODEC_SK_FASTPFOR256_VB, CODEC_SK_IBP32_IVB, CODEC_VB, CODEC_XOR128_IVB, CODEC_OPTPFD_VB, FLAG_DELTA, FRAME_APP, FRAME_WORDS, CapacityError, McpError, build)
from .codec import (BinaryPacking, Composition, FastPFOR, FastPFOR128, IntCompressor, IntegerCODEC, IntegratedBinaryPacking,
                    IntegratedComposition, IntegratedIntCompressor, IntegratedVariableByte, IntegratedVariableByteCodec, IntWrapper,
                    OptPFD, SkippableComposition, XorBinaryPacking,
                    VariableByte,
                    VariableByteCodec, default_context)
from .engine import Context, SimulatePipeline
from .portfolio import LoadPortfolios, SimulatePortfolios, simulation_documents
from . import sharding

__all__ = [n for n in dir() if not n.startswith("_")]
