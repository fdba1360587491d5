"""genometools_b200 -- B200-native enhanced-suffix-array construction behind
`gt suffixerator` (-suf -lcp -bck).

Only what the hot path needs lives here:

  csrc/            hand-written sm_100a CUDA kernels + the C-ABI (include/gtb200.h)
  _lib.py          ctypes binding of libgtb200.so (fails loudly when missing)
  encseq.py        host mirror of what GtEncseq exports for this path
  suffixerator.py  host mirror of the `gt suffixerator` interface + file writers
  sharding.py      bucket-code ranges for -parts / multi-GPU
"""
from .suffixerator import Suffixerator, SuffixeratorOptions, EsaResult, suffixerator_main  # noqa: F401
from .encseq import EncodedSequence, encode_fasta, encode_symbols  # noqa: F401

__all__ = ["Suffixerator", "SuffixeratorOptions", "EsaResult", "suffixerator_main",
           "EncodedSequence", "encode_fasta", "encode_symbols"]
