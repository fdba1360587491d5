"""ctypes binding of libgtb200.so (C-ABI declared in include/gtb200.h).

There is no fallback: if the shared library is missing or no CUDA device is
usable every entry point raises.  The library is built in-tree by
`__graft_entry__.build()` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgtb200.so")


class GtbRange(C.Structure):
    _fields_ = [("start", C.c_uint64), ("end", C.c_uint64)]


class GtbStats(C.Structure):
    _fields_ = [
        ("totallength", C.c_uint64), ("specialcharacters", C.c_uint64),
        ("nonspecials", C.c_uint64), ("sa_offset", C.c_uint64),
        ("longest", C.c_uint64), ("numoflargelcpvalues", C.c_uint64),
        ("maxbranchdepth", C.c_uint64), ("lcptabsum", C.c_double),
        ("prefixlength", C.c_uint32), ("numofchars", C.c_uint32),
        ("unresolved_after_first_sort", C.c_uint64),
        ("doubling_rounds", C.c_uint32), ("radix_passes", C.c_uint32),
        ("radix_pairs_moved", C.c_uint64), ("kernel_launches", C.c_uint32),
        ("ms_total", C.c_float), ("ms_upload", C.c_float), ("ms_count", C.c_float),
        ("ms_hist", C.c_float), ("ms_radix", C.c_float), ("ms_analyze", C.c_float),
        ("ms_doubling", C.c_float), ("ms_lcp", C.c_float), ("ms_tail", C.c_float),
        ("ms_radix_first", C.c_float), ("radix_passes_first", C.c_uint32), ("radix_pairs_first", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class GtbFastaRequest(C.Structure):
    _fields_ = [
        ("filenames", C.POINTER(C.c_char_p)), ("numoffiles", C.c_uint64), ("indexname", C.c_char_p),
        ("symbolmap", C.POINTER(C.c_uint8)), ("decode", C.c_char_p), ("numofchars", C.c_uint),
        ("alphatype", C.c_uint), ("bits_per_symbol", C.c_uint),
        ("out_des", C.c_int), ("out_sds", C.c_int), ("out_ssp", C.c_int), ("out_md5", C.c_int),
        ("clip_desc", C.c_int), ("sat", C.c_char_p), ("threads", C.c_int),
    ]


class GtbFastaSummary(C.Structure):
    _fields_ = [
        ("totallength", C.c_uint64), ("numofsequences", C.c_uint64), ("numoffiles", C.c_uint64),
        ("specialcharacters", C.c_uint64), ("specialranges", C.c_uint64), ("realspecialranges", C.c_uint64),
        ("wildcards", C.c_uint64), ("wildcardranges", C.c_uint64), ("realwildcardranges", C.c_uint64),
        ("sat", C.c_uint64), ("satsep", C.c_uint64), ("characterdistribution", C.c_uint64 * 32),
        ("input_bytes", C.c_uint64), ("satname", C.c_char * 16), ("threads", C.c_uint),
        ("seconds_count", C.c_double), ("seconds_emit", C.c_double), ("seconds_lists", C.c_double),
        ("seconds_pack", C.c_double), ("seconds_md5", C.c_double), ("seconds_write", C.c_double),
        ("seconds_total", C.c_double),
    ]

    def as_dict(self):
        d = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            d[k] = list(v) if k == "characterdistribution" else (v.decode() if isinstance(v, bytes) else v)
        return d


GTB_FASTA_OK, GTB_FASTA_UNSUPPORTED, GTB_FASTA_ERROR = 0, 1, -1

# every symbol include/gtb200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_U64 = C.c_uint64
SYMBOLS = {
    "gtb_abi_version": (C.c_int, []),
    "gtb_fasta_encode": (C.c_int, [C.POINTER(GtbFastaRequest), C.POINTER(GtbFastaSummary), C.c_char_p, C.c_size_t]),
    "gtb_device_count": (C.c_int, []),
    "gtb_release_devices": (C.c_int, []),
    "gtb_esa_new": (_P, [C.c_int, C.c_char_p, C.c_size_t]),
    "gtb_esa_delete": (None, [_P]),
    "gtb_esa_error": (C.c_char_p, [_P]),
    "gtb_esa_set_input_2bit": (C.c_int, [_P, _P, _U64, _U64, _P, _U64]),
    "gtb_esa_set_input_bytes": (C.c_int, [_P, _P, _U64, C.c_uint]),
    "gtb_esa_set_readmode": (C.c_int, [_P, C.c_uint]),
    "gtb_esa_share_input": (C.c_int, [_P, _P]),
    "gtb_esa_set_separators": (C.c_int, [_P, _P, _U64]),
    "gtb_esa_copy_bwttab": (C.c_int, [_P, _P, _U64, _U64]),
    "gtb_esa_set_code_range": (C.c_int, [_P, _U64, _U64, _U64, C.c_int]),
    "gtb_esa_run": (C.c_int, [_P, C.c_uint, C.c_uint]),
    "gtb_esa_count": (C.c_int, [_P, C.c_uint]),
    "gtb_esa_count_partial": (C.c_int, [_P, C.c_uint, _U64, _U64]),
    "gtb_esa_count_finish": (C.c_int, [_P]),
    "gtb_esa_dev_bcktab": (C.c_int, [_P, _P, _P, _P]),
    "gtb_esa_split_ranges": (C.c_int, [_P, C.c_uint, _P, _P]),
    "gtb_esa_coarse_partial": (C.c_int, [_P, C.c_uint, _U64, _U64, _P, _P]),
    "gtb_esa_coarse_split": (C.c_int, [_P, C.c_uint, _P, _P]),
    "gtb_esa_set_code_range_known": (C.c_int, [_P, _U64, _U64, _U64, _U64, C.c_int]),
    "gtb_esa_slice_partition": (C.c_int, [_P, C.c_uint, _U64, _U64, _P, C.c_int, _P, _P, _U64, _P]),
    "gtb_esa_sort_begin_positions": (C.c_int, [_P, C.c_uint, C.c_uint, _P, _U64]),
    "gtb_esa_sort_begin_pairs": (C.c_int, [_P, C.c_uint, C.c_uint, _P, _P, _U64]),
    "gtb_esa_sort_begin": (C.c_int, [_P, C.c_uint, C.c_uint]),
    "gtb_esa_unresolved": (_U64, [_P]),
    "gtb_esa_ensure_ranks": (C.c_int, [_P]),
    "gtb_esa_round_local": (C.c_int, [_P]),
    "gtb_esa_round_prepare": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _U64, _P]),
    "gtb_esa_rank_lookup": (C.c_int, [_P, _P, _U64, _P]),
    "gtb_esa_round_finish": (C.c_int, [_P, _P]),
    "gtb_esa_sort_end": (C.c_int, [_P]),
    "gtb_code_first_key": (_U64, [C.c_uint, C.c_uint, _U64]),
    "gtb_esa_get_stats": (C.c_int, [_P, C.POINTER(GtbStats)]),
    "gtb_esa_boundary_keys": (C.c_int, [_P, C.POINTER(_U64), C.POINTER(_U64)]),
    "gtb_esa_fix_seam": (C.c_int, [_P, _U64]),
    "gtb_esa_num_entries": (_U64, [_P]),
    "gtb_esa_num_llv": (_U64, [_P]),
    "gtb_esa_copy_suftab_u64": (C.c_int, [_P, _P, _U64, _U64]),
    "gtb_esa_copy_suftab_u32": (C.c_int, [_P, _P, _U64, _U64]),
    "gtb_esa_copy_lcptab": (C.c_int, [_P, _P, _U64, _U64]),
    "gtb_esa_copy_tables": (C.c_int, [_P, _P, _P, _U64, _U64]),
    "gtb_esa_copy_results": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "gtb_esa_run_to_host": (C.c_int, [_P, C.c_uint, C.c_uint, _P, _P, _P, _U64, C.POINTER(_U64), _P, _P, _P]),
    "gtb_esa_copy_llv": (C.c_int, [_P, _P]),
    "gtb_esa_copy_bcktab": (C.c_int, [_P, _P, _P, _P]),
    "gtb_bck_sizes": (None, [C.c_uint, C.c_uint, C.POINTER(_U64), C.POINTER(_U64), C.POINTER(_U64)]),
    "gtb_esa_hash_results": (C.c_int, [_P, _U64, C.POINTER(_U64)]),
    "gtb_esa_hash_bcktab": (C.c_int, [_P, C.POINTER(_U64)]),
    "gtb_esa_run_sharded": (C.c_int, [_P, C.c_uint, C.c_uint, C.c_int, C.c_int, _P, _P, C.c_int]),
    "gtb_esa_llv_before": (_U64, [_P]),
    "gtb_group_new": (_P, [C.POINTER(C.c_int), C.c_int, C.c_char_p, C.c_size_t]),
    "gtb_group_delete": (None, [_P]),
    "gtb_group_error": (C.c_char_p, [_P]),
    "gtb_group_size": (C.c_int, [_P]),
    "gtb_group_range": (_P, [_P, C.c_int]),
    "gtb_group_set_readmode": (C.c_int, [_P, C.c_uint]),
    "gtb_group_set_input_2bit": (C.c_int, [_P, _P, _U64, _U64, _P, _U64]),
    "gtb_group_set_input_bytes": (C.c_int, [_P, _P, _U64, C.c_uint]),
    "gtb_group_set_separators": (C.c_int, [_P, _P, _U64]),
    "gtb_group_run": (C.c_int, [_P, C.c_uint, C.c_uint]),
    "gtb_group_get_stats": (C.c_int, [_P, C.POINTER(GtbStats)]),
    "gtb_group_num_entries": (_U64, [_P]),
    "gtb_group_num_llv": (_U64, [_P]),
    "gtb_group_copy_results": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "gtb_group_copy_bwttab": (C.c_int, [_P, _P]),
    "gtb_group_hash_results": (C.c_int, [_P, C.POINTER(_U64)]),
    "gtb_esa_stream": (_P, [_P]),
    "gtb_esa_dev_suftab": (_P, [_P]),
    "gtb_esa_dev_lcptab": (_P, [_P]),
    "gtb_esa_dev_leftborder": (_P, [_P]),
    "gtb_esa_build_2bit": (C.c_int, [C.c_int, _P, _U64, _U64, _P, _U64, C.c_uint, _P, _P, _P, _U64,
                                     C.POINTER(_U64), _P, _P, _P, C.POINTER(GtbStats), C.c_char_p, C.c_size_t]),
    "gtb_esa_build_bytes": (C.c_int, [C.c_int, _P, _U64, C.c_uint, C.c_uint, _P, _P, _P, _U64,
                                      C.POINTER(_U64), _P, _P, _P, C.POINTER(GtbStats), C.c_char_p, C.c_size_t]),
    "gtb_radixsort_pairs_u64_u32": (C.c_int, [C.c_int, _P, _P, _U64, C.c_uint, C.c_uint, C.c_char_p, C.c_size_t]),
    "gtb_radixsort_u64": (C.c_int, [C.c_int, _P, _U64, C.c_char_p, C.c_size_t]),
    "gtb_radixsort_u64pair": (C.c_int, [C.c_int, _P, _U64, C.c_char_p, C.c_size_t]),
    "gtb_radixsort_u64keypair": (C.c_int, [C.c_int, _P, _U64, C.c_char_p, C.c_size_t]),
}

GTB_WANT_SUF, GTB_WANT_LCP, GTB_WANT_BCK, GTB_REUSE_COUNTS = 1, 2, 4, 8
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)

_lib = None


class GtbError(RuntimeError):
    pass


def load():
    """Load libgtb200.so; raises GtbError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GtbError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). genometools_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """raw address of a C-contiguous numpy array (or None)"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data
