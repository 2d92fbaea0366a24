"""exon_duckdb_b200 -- B200-native (sm_100a) FASTA/FASTQ scan engine behind the
exon DuckDB extension's read_fasta / read_fastq / sequence-function path.

Layers:
  csrc/            hand-written CUDA kernels + the C ABI (include/exon_b200.h)
  _lib.py          ctypes binding of libexon_b200.so (fails loudly if it is not built)
  device.py        device-resident pipeline on torch-owned HBM buffers
  functions.py     host mirror of the reference's table / scalar functions
  dist.py          byte-range sharding across ranks (torch.distributed)
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
