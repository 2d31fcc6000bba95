"""metacherchant_b200 — B200-native (sm_100a) drop-in for MetaCherchant's environment-finder hot path.

The product is the C-ABI library ``csrc/libmcgpu.so`` (include/mc_b200.h).  This package is the thin
Python host layer over it: a ctypes binding (`_ffi`), mirrors of the reference's host-side entry points
(`io_utils.load_reads` ~ IOUtils.loadReads, `KmerTable.get` ~ BigLong2ShortHashMap.get, ...) and the
one-process-per-GPU orchestration over torch.distributed (`dist`).  There is no CPU fallback: every
compute call fails loudly when the CUDA library or a B200 is missing.
"""
from ._ffi import McError, lib_path, load_library  # noqa: F401
from .api import Context, KmerTable, FMT_FASTA, FMT_FASTQ, FMT_FASTQ33, FMT_FASTQ64  # noqa: F401

__version__ = "0.1.0"
