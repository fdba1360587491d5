"""Host mirror of what GtEncseq hands to the sorter for this path.

In the drop-in the unchanged GenomeTools encoder produces the
GtEncseq and our replacement `gt_suffixerator` only *exports* from it
(gt_encseq_twobitencoding_export, /root/reference/src/core/encseq.c:6687;
gt_specialrangeiterator_*, src/core/encseq.h:127-133;
gt_encseq_extract_encoded for non-2-bit alphabets).  This module produces the
same three things from FASTA / symbol arrays so that the library can be used,
tested and benchmarked without a GenomeTools build.
"""
import os
from dataclasses import dataclass, field
import numpy as np

WILDCARD = 254     # src/core/chardef.h
SEPARATOR = 255

DNA_BASES = "acgt"
DNA_WILDCARDS = "nsywrkvbdhmNSYWRKVBDHM"           # src/core/alphabet.c:84
PROTEIN_AMINOACIDS = "LVIFKREDAGSTNQYWPHMC"        # src/core/alphabet.c:87
PROTEIN_WILDCARDS = "XUBZJO*-"                     # src/core/alphabet.c:90
UNDEF = 253


def _dna_map():
    m = np.full(256, UNDEF, dtype=np.uint8)
    for i, ch in enumerate(DNA_BASES):              # assign_dna_symbolmap, alphabet.c:337-358
        m[ord(ch)] = i
        m[ord(ch.upper())] = i
    m[ord("u")] = m[ord("U")] = 3
    for ch in DNA_WILDCARDS:
        m[ord(ch)] = WILDCARD
    return m


def _protein_map():
    m = np.full(256, UNDEF, dtype=np.uint8)
    for i, ch in enumerate(PROTEIN_AMINOACIDS):     # alphabet.c:496-503 (upper case only)
        m[ord(ch)] = i
    for ch in PROTEIN_WILDCARDS:
        m[ord(ch)] = WILDCARD
    return m


ALPHABETS = {"dna": (4, _dna_map()), "protein": (20, _protein_map())}


@dataclass
class EncodedSequence:
    """symbols: uint8[n], 0..numofchars-1 regular, 254 wildcard, 255 separator."""
    symbols: np.ndarray
    numofchars: int
    numofsequences: int = 1
    _twobit: tuple = field(default=None, repr=False)

    @property
    def totallength(self):
        return int(self.symbols.shape[0])

    @property
    def is_dna(self):
        return self.numofchars == 4

    def special_mask(self):
        return self.symbols >= WILDCARD

    def special_ranges(self):
        """maximal special runs [start, end) ascending = gt_specialrangeiterator (forward)"""
        sp = self.special_mask()
        if sp.size == 0:
            return np.zeros((0, 2), dtype=np.uint64)
        d = np.diff(np.concatenate(([0], sp.view(np.int8), [0])))
        starts = np.flatnonzero(d == 1)
        ends = np.flatnonzero(d == -1)
        return np.stack([starts, ends], axis=1).astype(np.uint64)

    def twobitencoding(self, filler=None):
        """uint64 words as gt_encseq_twobitencoding_export delivers them (intbits.h:78-83):
        base i in word i/32 at bits 62-2*(i%32).  Special positions carry `filler`
        (an arbitrary base; GenomeTools stores a filler there too, encseq.c:2599)."""
        if self._twobit is not None and filler is None:
            return self._twobit
        assert self.numofchars == 4
        n = self.totallength
        nwords = n // 32 + 2
        s = np.zeros(nwords * 32, dtype=np.uint8)
        s[:n] = self.symbols
        sp = s >= WILDCARD
        if filler is None:
            s[sp] = 0
        else:
            f = np.asarray(filler, dtype=np.uint8)
            s[sp] = f if f.ndim == 0 else f[: int(sp.sum())]
        s &= 3
        q = s.reshape(-1, 4)
        packed = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
        words = np.ascontiguousarray(packed.reshape(-1, 8)).view(">u8").astype(np.uint64).reshape(-1)
        out = (words, self.special_ranges())
        if filler is None:
            self._twobit = out
        return out

    def specialcharinfo(self):
        """the encseq-derived numbers of the .prj file (sfx-outprj.c:52-63). `specialranges`
        / `wildcardranges` equal the real counts whenever no run exceeds the access type's
        range-length limit (GtEncseq splits longer runs, encseq.c:5215-5255)."""
        sp = self.special_mask()
        wc = self.symbols == WILDCARD

        def runs(m):
            if m.size == 0:
                return 0
            d = np.diff(np.concatenate(([0], m.view(np.int8))))
            return int((d == 1).sum())

        def prefix_len(m):
            nz = np.flatnonzero(~m)
            return int(m.size if nz.size == 0 else nz[0])

        return {
            "specialcharacters": int(sp.sum()), "specialranges": runs(sp), "realspecialranges": runs(sp),
            "lengthofspecialprefix": prefix_len(sp), "lengthofspecialsuffix": prefix_len(sp[::-1]),
            "wildcards": int(wc.sum()), "wildcardranges": runs(wc), "realwildcardranges": runs(wc),
            "lengthofwildcardprefix": prefix_len(wc), "lengthofwildcardsuffix": prefix_len(wc[::-1]),
        }


def encode_symbols(symbols, numofchars, numofsequences=None):
    s = np.ascontiguousarray(symbols, dtype=np.uint8)
    bad = (s >= numofchars) & (s < WILDCARD)
    if bad.any():
        raise ValueError(f"symbol {int(s[bad][0])} out of range for alphabet size {numofchars}")
    if numofsequences is None:
        numofsequences = int((s == SEPARATOR).sum()) + 1
    return EncodedSequence(s, int(numofchars), int(numofsequences))


def parse_fasta_bytes(raw, symbolmap):
    """FASTA bytes -> (symbol array with one SEPARATOR between consecutive records, #records)."""
    data = np.frombuffer(raw, dtype=np.uint8)
    if data.size == 0:
        raise ValueError("empty sequence file")
    if data[0] != ord(">"):
        raise ValueError("first character of fasta file has to be '>'")
    nl = data == ord("\n")
    line_start = np.concatenate(([True], nl[:-1]))
    idx = np.arange(data.size, dtype=np.int64)
    last_ls = np.maximum.accumulate(np.where(line_start, idx, 0))
    in_hdr = (data == ord(">"))[last_ls]            # byte belongs to a line starting with '>'
    hdr_start = line_start & in_hdr
    nrec = int(hdr_start.sum())
    keep = ~in_hdr & ~nl & (data != ord("\r")) & (data != ord(" ")) & (data != ord("\t"))
    sepmark = hdr_start.copy()
    sepmark[0] = False                              # records 2.. are preceded by a separator
    out = np.where(sepmark, np.uint8(SEPARATOR), symbolmap[data])
    sel = keep | sepmark
    out = out[sel]
    if (out == UNDEF).any():
        badpos = np.flatnonzero(sel)[np.flatnonzero(out == UNDEF)[0]]
        raise ValueError(f"illegal character '{chr(int(data[badpos]))}' in sequence file")
    return np.ascontiguousarray(out), nrec


def encode_fasta(paths, alphabet="dna"):
    """FASTA file(s) -> EncodedSequence (files are concatenated with a separator, as
    gt_encseq_encoder_encode does for several -db arguments)."""
    if isinstance(paths, (str, bytes)):
        paths = [paths]
    numofchars, symbolmap = ALPHABETS[alphabet]
    parts, nseq = [], 0
    for p in paths:
        with open(p, "rb") as fh:
            raw = fh.read()
        s, r = parse_fasta_bytes(raw, symbolmap)
        if parts:
            parts.append(np.array([SEPARATOR], dtype=np.uint8))
        parts.append(s)
        nseq += r
    sym = np.concatenate(parts) if len(parts) > 1 else parts[0]
    return EncodedSequence(np.ascontiguousarray(sym), numofchars, nseq)


class FastaUnsupported(Exception):
    """the input is outside what gtb_fasta_encode covers (nothing was written): use the reference's encoder"""


def decode_table(alphabet="dna"):
    """gt_alphabet_decode (src/core/alphabet.c:84-92,466-548): DNA a c g t with the wildcard shown as n,
    protein L V I ... C with the wildcard shown as X"""
    t = bytearray(b"\0" * 256)
    chars, wildcard = (DNA_BASES, "n") if alphabet == "dna" else (PROTEIN_AMINOACIDS, "X")
    for i, ch in enumerate(chars):
        t[i] = ord(ch)
    t[WILDCARD] = ord(wildcard)
    return bytes(t)


def write_index_files(paths, indexname, des=True, sds=True, ssp=True, md5=True, clip_desc=False, threads=0,
                      alphabet="dna", sat=None):
    """FASTA file(s) -> <indexname>.esq/.ssp/.des/.sds/.md5 as `gt encseq encode -dna|-protein` /
    `gt suffixerator -dna|-protein -tis` write them (gtb_fasta_encode, include/gtb200.h; host code of libgtb200.so).
    Returns the summary dict; raises FastaUnsupported when the library declines the input."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    if isinstance(paths, (str, bytes)):
        paths = [paths]
    names = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    numofchars, symbolmap = ALPHABETS[alphabet]
    symbolmap = np.ascontiguousarray(symbolmap, dtype=np.uint8)
    rq = _lib.GtbFastaRequest()
    rq.filenames = names
    rq.numoffiles = len(paths)
    rq.indexname = os.fsencode(indexname)
    rq.symbolmap = symbolmap.ctypes.data_as(C.POINTER(C.c_uint8))
    rq.decode = decode_table(alphabet)
    rq.numofchars = numofchars
    rq.alphatype = 0 if alphabet == "dna" else 1
    rq.bits_per_symbol = 3 if alphabet == "dna" else 5       # src/core/alphabet.c:476,543
    rq.out_des, rq.out_sds, rq.out_ssp, rq.out_md5 = int(des), int(sds), int(ssp), int(md5)
    rq.clip_desc = int(clip_desc)
    rq.sat = sat.encode() if sat else None
    rq.threads = int(threads)
    summary = _lib.GtbFastaSummary()
    msg = C.create_string_buffer(1024)
    rc = lib.gtb_fasta_encode(C.byref(rq), C.byref(summary), msg, 1024)
    if rc == _lib.GTB_FASTA_UNSUPPORTED:
        raise FastaUnsupported(msg.value.decode(errors="replace"))
    if rc != _lib.GTB_FASTA_OK:
        raise _lib.GtbError(msg.value.decode(errors="replace"))
    return summary.as_dict()
