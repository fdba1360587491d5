"""Host mirror of the `gt suffixerator` interface for the accelerated path.

  gt suffixerator -db F.. (-dna|-protein) [-suf] [-lcp] [-bck] [-bwt] [-pl [k]] [-parts p]
                  [-dir fwd|rev|cpl|rcl] -indexname I                (/root/reference/src/match/sfx-opt.c:34-122,
                                                index_options.c:274-521, encseq_options.c:181-326)

Same option names, same meaning, same error behaviour for the options this path
covers; everything the path does not cover fails loudly instead of silently
falling back.  The sort core is libgtb200.so; there is no
CPU path in this module.

Outputs (byte-identical to the reference):
  I.suf  uint64[n+1]                 gt_suffixsortspace_to_file, sfx-suffixgetset.c:462-477
  I.lcp  uint8[n+1]                  outlcpvalues / tail zeros, sfx-lcpvalues.c:371-471
  I.llv  {uint64 idx, uint64 lcp}[]  Largelcpvalue, lcpoverflow.h:25-29
  I.bck  three uint32 tables, each padded to 8 bytes: gt_bcktab_flush_to_file,
         bcktab.c:519-577 + gt_mapspec_write, src/core/mapspec.c:350-466
  I.prj  text, sfx-outprj.c:38-81
"""
from dataclasses import dataclass, field
import ctypes as C
import os
import sys
import numpy as np

from . import _lib
from ._lib import GtbError, GtbStats, ptr, GTB_WANT_SUF, GTB_WANT_LCP, GTB_WANT_BCK, GTB_REUSE_COUNTS
from .encseq import EncodedSequence, FastaUnsupported, encode_fasta, write_index_files
from .sharding import suftab_parts

GT_RECOMMENDED_MULTIPLIER_DEFAULT = 0.25       # sfx-apfxlen.h:23
GT_MAXMULTIPLIEROFTOTALLENGTH = 4.0            # sfx-apfxlen.c:44
LCPOVERFLOW = 255                              # lcpoverflow.h:23
UINT32_MAX = 0xFFFFFFFF
READMODES = ("fwd", "rev", "cpl", "rcl")       # GtReadmode, src/core/readmode.c:25-30


# ---------------------------------------------------------------- a14: prefix length policy
def maxbasepower(numofchars):
    """gt_maxbasepower, src/match/initbasepower.c:23-34 (GtCodetype is 32 bit)"""
    minfailure = UINT32_MAX // numofchars
    thepower, i = 1, 0
    while thepower < minfailure:
        thepower *= numofchars
        i += 1
    return i


def bcktab_numofdistpfxidx(numofchars, prefixlength):
    return sum(numofchars ** i for i in range(1, prefixlength - 1))


def bcktab_sizeoftable(numofchars, prefixlength, maxvalue, withspecialsuffixes=True):
    """gt_bcktab_sizeoftable, src/match/bcktab.c:289-324"""
    base = 8 if maxvalue > UINT32_MAX else 4
    size = base * (numofchars ** prefixlength + 1)
    if withspecialsuffixes:
        size += base * numofchars ** (prefixlength - 1) if prefixlength >= 1 else base
        size += base * bcktab_numofdistpfxidx(numofchars, prefixlength)
    return size


def _prefixlengthwithmaxspace(numofchars, maxbytes, factor, maxvalue):
    pl = 1
    while True:                                   # sfx-apfxlen.c:49-80
        if bcktab_sizeoftable(numofchars, pl, maxvalue) / factor > maxbytes:
            return pl - 1
        pl += 1


def recommendedprefixlength(numofchars, totallength):
    """gt_recommendedprefixlength, src/match/sfx-apfxlen.c:82-107"""
    pl = _prefixlengthwithmaxspace(numofchars, totallength, GT_RECOMMENDED_MULTIPLIER_DEFAULT,
                                   totallength + 1)
    if pl == 0:
        return 1
    mbp = maxbasepower(numofchars)
    return min(mbp, pl) if mbp >= 1 else pl


def whatisthemaximalprefixlength(numofchars, totallength):
    """gt_whatisthemaximalprefixlength with prefixlenbits == 0, sfx-apfxlen.c:109-147"""
    m = _prefixlengthwithmaxspace(numofchars, totallength, GT_MAXMULTIPLIEROFTOTALLENGTH, totallength + 1)
    m = min(maxbasepower(numofchars), m)
    return 1 if m == 0 else m


# ---------------------------------------------------------------- options
@dataclass
class SuffixeratorOptions:
    db: list = field(default_factory=list)
    indexname: str = None
    dna: bool = False
    protein: bool = False
    suf: bool = False
    lcp: bool = False
    bck: bool = False
    bwt: bool = False
    pl: int = None            # None: option absent; 0: "-pl" without argument (automatic)
    parts: int = 1
    dir: str = "fwd"
    device: int = 0
    verbose: bool = False
    tis: bool = False         # -tis: write <indexname>.esq (+ .ssp .des .sds .md5 unless switched off)
    des: bool = True
    sds: bool = True
    ssp: bool = True
    md5: bool = True

    UNSUPPORTED = ("-mirrored", "-dc", "-spmopt", "-sortmaxdepth", "-suftabuint",
                   "-compressedoutput", "-genomediff", "-lcpdist", "-memlimit", "-algbds",
                   "-cmpcharbychar", "-maxdepth", "-ii", "-smap", "-sat", "-kys", "-dccheck",
                   "-samplewithprefixlengthnull", "-storespecialcodes", "-showprogress")

    @classmethod
    def parse(cls, argv):
        o = cls()
        i = 0
        flags = {"-dna": "dna", "-protein": "protein", "-suf": "suf", "-lcp": "lcp", "-bck": "bck",
                 "-bwt": "bwt", "-v": "verbose"}
        tables = ("-tis", "-des", "-sds", "-ssp", "-md5")   # the encoder's files (gtb_fasta_encode, DNA only)
        while i < len(argv):
            a = argv[i]
            if a in flags:
                setattr(o, flags[a], True)
            elif a == "-db":
                i += 1
                while i < len(argv) and not argv[i].startswith("-"):
                    o.db.append(argv[i]); i += 1
                if not o.db:
                    raise GtbError("missing argument to option \"-db\"")
                continue
            elif a == "-indexname":
                i += 1
                if i >= len(argv):
                    raise GtbError("missing argument to option \"-indexname\"")
                o.indexname = argv[i]
            elif a == "-pl":
                if i + 1 < len(argv) and argv[i + 1].isdigit():
                    i += 1; o.pl = int(argv[i])
                    if o.pl < 1:
                        raise GtbError("argument to option \"-pl\" must be an integer >= 1")
                else:
                    o.pl = 0
            elif a == "-parts":
                i += 1
                if i >= len(argv) or not argv[i].isdigit() or int(argv[i]) < 1:
                    raise GtbError("argument to option \"-parts\" must be a positive integer")
                o.parts = int(argv[i])
            elif a == "-dir":
                i += 1
                if i >= len(argv):
                    raise GtbError("missing argument to option \"-dir\"")
                if argv[i] not in READMODES:
                    raise GtbError("unknown readmode, must be fwd or rev or cpl or rcl")   # readmode.c:44
                o.dir = argv[i]
            elif a == "-device":
                i += 1; o.device = int(argv[i])
            elif a in tables:
                value = True
                if i + 1 < len(argv) and argv[i + 1] in ("yes", "no"):
                    i += 1
                    value = argv[i] == "yes"
                setattr(o, a[1:], value)
            elif a in cls.UNSUPPORTED:
                raise GtbError(f"option \"{a}\" is not supported by the B200 suffixerator path "
                               "(no silent fallback); use the CPU `gt suffixerator` for it")
            else:
                raise GtbError(f"unknown option: {a} (try option -help)")
            i += 1
        if not o.db:
            raise GtbError("option \"-db\" is mandatory")       # suffixerator needs -db or -ii
        if o.dna and o.protein:
            raise GtbError("option \"-dna\" and option \"-protein\" exclude each other")
        if o.indexname is None:
            if len(o.db) > 1:
                raise GtbError("if more than one input file is given, then option -indexname is mandatory")
            import os
            o.indexname = os.path.basename(o.db[0])
        if o.bck and o.pl is None:
            o.pl = 0
        if o.dir in ("cpl", "rcl") and o.protein:          # sfx-run.c:541-549
            raise GtbError(f"option -{o.dir} only can be used for DNA alphabets")
        if o.dir != "fwd" and not (o.suf or o.lcp or o.bwt):   # sfx-run.c:586-593
            raise GtbError(f"option '-dir {o.dir}' only makes sense in combination with at least one of "
                           "the options -suf, -lcp, or -bwt")
        return o


@dataclass
class EsaResult:
    totallength: int
    numofchars: int
    prefixlength: int
    suftab: np.ndarray = None        # uint64[n+1]
    lcptab: np.ndarray = None        # uint8[n+1]
    llvtab: np.ndarray = None        # uint64[k,2]
    bwttab: np.ndarray = None        # uint8[n+1] (option -bwt)
    leftborder: np.ndarray = None    # uint32
    countspecialcodes: np.ndarray = None
    distpfxidx: np.ndarray = None
    longest: int = None
    device_hashes: dict = None       # checksums of the tables in HBM (gtb_group_hash_results)
    job_stats: dict = None           # a sharded job: the numbers of the whole job
    numoflargelcpvalues: int = 0
    maxbranchdepth: int = 0
    lcptabsum: float = 0.0
    readmode: int = 0
    stats: list = field(default_factory=list)   # one gtb_stats dict per part

    @property
    def averagelcp(self):
        return self.lcptabsum / (self.totallength + 1)

    # ---- file images ----
    def suf_bytes(self):
        return self.suftab.astype("<u8", copy=False).tobytes()

    def lcp_bytes(self):
        return self.lcptab.tobytes()

    def bwt_bytes(self):
        return self.bwttab.tobytes()

    def llv_bytes(self):
        return self.llvtab.astype("<u8", copy=False).tobytes()

    def bck_bytes(self):
        out = bytearray()
        for t in (self.leftborder, self.countspecialcodes, self.distpfxidx):
            b = t.astype("<u4", copy=False).tobytes()
            out += b
            out += b"\0" * (-len(b) % 8)                # mapspec pads every table to 8 bytes
        return bytes(out)

    def prj_text(self, specialcharinfo, numofsequences, with_lcp=True):
        """sfx-outprj.c:38-81; the encseq-derived lines come from `specialcharinfo`"""
        L = [f"totallength={self.totallength}"]
        for k in ("specialcharacters", "specialranges", "realspecialranges", "lengthofspecialprefix",
                  "lengthofspecialsuffix", "wildcards", "wildcardranges", "realwildcardranges",
                  "lengthofwildcardprefix", "lengthofwildcardsuffix"):
            L.append(f"{k}={specialcharinfo[k]}")
        L += [f"numofsequences={numofsequences}", f"numofdbsequences={numofsequences}",
              "numofquerysequences=0", f"numberofallsortedsuffixes={self.totallength + 1}"]
        if self.longest is not None:
            L.append(f"longest={self.longest}")
        L.append(f"prefixlength={self.prefixlength}")
        L.append(f"largelcpvalues={self.numoflargelcpvalues if with_lcp else 0}")
        L.append("averagelcp=%.2f" % (self.averagelcp if with_lcp else 0.0))
        L.append(f"maxbranchdepth={self.maxbranchdepth if with_lcp else 0}")
        L += ["integersize=64", "littleendian=1", f"readmode={self.readmode}", "mirrored=0"]
        return "\n".join(L) + "\n"


class Suffixerator:
    """Sorter object bound to one CUDA device (stands in for Sfxiterator,
    /root/reference/src/match/sfx-suffixer.h:33-72)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        buf = C.create_string_buffer(512)
        self.device = device
        self.h = self.lib.gtb_esa_new(device, buf, 512)
        if not self.h:
            raise GtbError(buf.value.decode() or "gtb_esa_new failed")
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.gtb_esa_delete(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise GtbError(self.lib.gtb_esa_error(self.h).decode())

    # ---- input ----
    def set_sequence(self, enc: EncodedSequence, filler=None, readmode="fwd"):
        """readmode: -dir fwd|rev|cpl|rcl -- the library rewrites the sequence in read direction"""
        if readmode not in READMODES:
            raise GtbError("unknown readmode, must be fwd or rev or cpl or rcl")
        self.enc = enc
        self.readmode = READMODES.index(readmode)
        self._ck(self.lib.gtb_esa_set_readmode(self.h, self.readmode))
        if enc.is_dna:
            words, ranges = enc.twobitencoding(filler)
            ranges = np.ascontiguousarray(ranges, dtype=np.uint64)
            self._keep = (words, ranges)
            self._ck(self.lib.gtb_esa_set_input_2bit(self.h, ptr(words), words.shape[0], enc.totallength,
                                                     ptr(ranges) if ranges.shape[0] else None, ranges.shape[0]))
            sep = np.ascontiguousarray(np.flatnonzero(enc.symbols == 255), dtype=np.uint64)
            self._ck(self.lib.gtb_esa_set_separators(self.h, ptr(sep) if sep.shape[0] else None, sep.shape[0]))
        else:
            self._keep = enc.symbols
            self._ck(self.lib.gtb_esa_set_input_bytes(self.h, ptr(enc.symbols), enc.totallength, enc.numofchars))

    def stats(self):
        st = GtbStats()
        self._ck(self.lib.gtb_esa_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def bucket_table(self, prefixlength):
        self._ck(self.lib.gtb_esa_count(self.h, prefixlength))
        return self._copy_bck(prefixlength)

    def _copy_bck(self, prefixlength):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.gtb_bck_sizes(self.enc.numofchars, prefixlength, C.byref(a), C.byref(b), C.byref(c))
        lb = np.empty(a.value + 1, dtype=np.uint32)
        csc = np.empty(b.value, dtype=np.uint32)
        dist = np.empty(c.value, dtype=np.uint32)
        self._ck(self.lib.gtb_esa_copy_bcktab(self.h, ptr(lb), ptr(csc), ptr(dist) if c.value else None))
        return lb, csc, dist

    # ---- the sort ----
    def _collect(self, h, res_lists, want_suf, want_lcp, want_bwt=False):
        suf_parts, lcp_parts, llv_parts = res_lists[:3]
        e = self.lib.gtb_esa_num_entries(h)
        if want_bwt:
            a = np.empty(e, dtype=np.uint8)
            self._ckh(h, self.lib.gtb_esa_copy_bwttab(h, ptr(a), 0, e))
            res_lists[3].append(a)
        if want_suf:
            a = np.empty(e, dtype=np.uint64)
            self._ckh(h, self.lib.gtb_esa_copy_suftab_u64(h, ptr(a), 0, e))
            suf_parts.append(a)
        if want_lcp:
            a = np.empty(e, dtype=np.uint8)
            self._ckh(h, self.lib.gtb_esa_copy_lcptab(h, ptr(a), 0, e))
            lcp_parts.append(a)
            k = self.lib.gtb_esa_num_llv(h)
            a = np.empty((k, 2), dtype=np.uint64)
            if k:
                self._ckh(h, self.lib.gtb_esa_copy_llv(h, ptr(a)))
            llv_parts.append(a)

    def _ckh(self, h, rc):
        if rc != 0:
            raise GtbError(self.lib.gtb_esa_error(h).decode())

    def _stats_of(self, h):
        st = GtbStats()
        self._ckh(h, self.lib.gtb_esa_get_stats(h, C.byref(st)))
        return st.as_dict()

    def run(self, prefixlength, want_suf=True, want_lcp=True, want_bck=True, parts=1, copy=True, want_bwt=False):
        """gt_Sfxiterator_next over all parts.  parts > 1 (option -parts) cuts the bucket
        codes into ranges (gt_suftabparts_new) that are sorted as independent problems --
        the same code path that puts one range on each GPU -- and concatenated."""
        enc = self.enc
        n = enc.totallength
        flags = (GTB_WANT_SUF if want_suf else 0) | (GTB_WANT_LCP if want_lcp else 0) | \
                (GTB_WANT_BCK if want_bck else 0)
        res = EsaResult(n, enc.numofchars, prefixlength, readmode=getattr(self, "readmode", 0))
        plist = []
        if parts > 1 and prefixlength >= 1:
            lb, _, _ = self.bucket_table(prefixlength)
            plist = suftab_parts(lb, parts)
        lists = ([], [], [], [])
        if len(plist) <= 1:
            self._ck(self.lib.gtb_esa_run(self.h, prefixlength, flags))
            handles = [self.h]
            extra = []
        else:
            from .multirange import GpuRangeWorker, run_ranges_local, range_first_keys
            buf = C.create_string_buffer(512)
            extra, workers = [], []
            try:
                for pi, (mincode, maxcode, off, _w) in enumerate(plist):
                    h = self.lib.gtb_esa_new(self.device, buf, 512)
                    if not h:
                        raise GtbError(buf.value.decode())
                    extra.append(h)
                    self._ckh(h, self.lib.gtb_esa_share_input(h, self.h))
                    self._ckh(h, self.lib.gtb_esa_set_code_range(h, mincode, maxcode, off,
                                                                 1 if pi == len(plist) - 1 else 0))
                    workers.append(GpuRangeWorker(h, prefixlength, flags | GTB_WANT_BCK, self.device))
                run_ranges_local(workers, range_first_keys(enc.numofchars, prefixlength, plist), want_lcp)
            except Exception:
                for h in extra:
                    self.lib.gtb_esa_delete(h)
                raise
            handles = extra
        try:
            lcpsum, maxbd, nlarge, longest = 0.0, 0, 0, None
            for h in handles:
                st = self._stats_of(h)
                res.stats.append(st)
                lcpsum += st["lcptabsum"]; maxbd = max(maxbd, st["maxbranchdepth"])
                nlarge += st["numoflargelcpvalues"]
                if st["longest"] != 0xFFFFFFFFFFFFFFFF:
                    longest = st["longest"]
                if copy:
                    self._collect(h, lists, want_suf, want_lcp, want_bwt)
            if want_bck and prefixlength >= 1:
                hb = handles[0]
                a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
                self.lib.gtb_bck_sizes(enc.numofchars, prefixlength, C.byref(a), C.byref(b), C.byref(c))
                lb = np.empty(a.value + 1, dtype=np.uint32); csc = np.empty(b.value, dtype=np.uint32)
                dist = np.empty(c.value, dtype=np.uint32)
                self._ckh(hb, self.lib.gtb_esa_copy_bcktab(hb, ptr(lb), ptr(csc), ptr(dist) if c.value else None))
                res.leftborder, res.countspecialcodes, res.distpfxidx = lb, csc, dist
        finally:
            for h in extra:
                self.lib.gtb_esa_delete(h)
        suf_parts, lcp_parts, llv_parts, bwt_parts = lists
        if copy and want_bwt:
            res.bwttab = np.concatenate(bwt_parts) if len(bwt_parts) > 1 else bwt_parts[0]
        if copy:
            if want_suf:
                res.suftab = np.concatenate(suf_parts) if len(suf_parts) > 1 else suf_parts[0]
            if want_lcp:
                res.lcptab = np.concatenate(lcp_parts) if len(lcp_parts) > 1 else lcp_parts[0]
                res.llvtab = np.concatenate(llv_parts) if len(llv_parts) > 1 else llv_parts[0]
        res.longest = longest
        res.lcptabsum, res.maxbranchdepth, res.numoflargelcpvalues = lcpsum, maxbd, nlarge
        return res


def build_esa_group(enc: EncodedSequence, prefixlength, devices, want_suf=True, want_lcp=True, want_bck=True,
                    filler=None, want_bwt=False, readmode="fwd"):
    """The job sharded over len(devices) code ranges inside this process (gtb_group, include/gtb200.h):
    range i runs on CUDA device devices[i] -- one range per GPU (`gt -j N`), or several ranges on one
    GPU (-parts).  The ranges reach each other's HBM directly (peer access); the results are gathered
    into one table by every range copying its shard to its offset."""
    lib = _lib.load()
    if readmode not in READMODES:
        raise GtbError("unknown readmode, must be fwd or rev or cpl or rcl")
    buf = C.create_string_buffer(512)
    devs = (C.c_int * len(devices))(*devices)
    g = lib.gtb_group_new(devs, len(devices), buf, 512)
    if not g:
        raise GtbError(buf.value.decode() or "gtb_group_new failed")

    def ck(rc):
        if rc != 0:
            raise GtbError(lib.gtb_group_error(g).decode())
    try:
        n = enc.totallength
        ck(lib.gtb_group_set_readmode(g, READMODES.index(readmode)))
        if enc.is_dna:
            words, ranges = enc.twobitencoding(filler)
            ranges = np.ascontiguousarray(ranges, dtype=np.uint64)
            ck(lib.gtb_group_set_input_2bit(g, ptr(words), words.shape[0], n,
                                            ptr(ranges) if ranges.shape[0] else None, ranges.shape[0]))
            sep = np.ascontiguousarray(np.flatnonzero(enc.symbols == 255), dtype=np.uint64)
            ck(lib.gtb_group_set_separators(g, ptr(sep) if sep.shape[0] else None, sep.shape[0]))
        else:
            ck(lib.gtb_group_set_input_bytes(g, ptr(enc.symbols), n, enc.numofchars))
        flags = (GTB_WANT_SUF if want_suf else 0) | (GTB_WANT_LCP if want_lcp else 0) | (GTB_WANT_BCK if want_bck else 0)
        ck(lib.gtb_group_run(g, prefixlength, flags))
        res = EsaResult(n, enc.numofchars, prefixlength, readmode=READMODES.index(readmode))
        st = GtbStats()
        ck(lib.gtb_group_get_stats(g, C.byref(st)))
        res.job_stats = st.as_dict()          # sums / maxima over the ranges (gtb_group_get_stats)
        for i in range(lib.gtb_group_size(g)):
            sti = GtbStats()
            lib.gtb_esa_get_stats(lib.gtb_group_range(g, i), C.byref(sti))
            res.stats.append(sti.as_dict())
        e = int(lib.gtb_group_num_entries(g))
        if e != n + 1:
            raise GtbError(f"the code ranges hold {e} entries, expected {n + 1}")
        k = int(lib.gtb_group_num_llv(g))
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.gtb_bck_sizes(enc.numofchars, prefixlength, C.byref(a), C.byref(b), C.byref(c))
        suf = np.empty(e, dtype=np.uint64) if want_suf else None
        lcp = np.empty(e, dtype=np.uint8) if want_lcp else None
        llv = np.empty((k, 2), dtype=np.uint64) if want_lcp else None
        lb = np.empty(a.value + 1, dtype=np.uint32) if want_bck else None
        csc = np.empty(b.value, dtype=np.uint32) if want_bck else None
        dist = np.empty(c.value, dtype=np.uint32) if want_bck else None
        ck(lib.gtb_group_copy_results(g, ptr(suf), ptr(lcp), ptr(llv) if (want_lcp and k) else None, ptr(lb),
                                      ptr(csc) if (want_bck and b.value) else None,
                                      ptr(dist) if (want_bck and c.value) else None))
        if want_bwt:
            res.bwttab = np.empty(e, dtype=np.uint8)
            ck(lib.gtb_group_copy_bwttab(g, ptr(res.bwttab)))
        res.suftab, res.lcptab, res.llvtab = suf, lcp, llv
        res.leftborder, res.countspecialcodes, res.distpfxidx = lb, csc, dist
        res.longest = None if st.longest == 0xFFFFFFFFFFFFFFFF else st.longest
        res.lcptabsum, res.maxbranchdepth, res.numoflargelcpvalues = st.lcptabsum, st.maxbranchdepth, st.numoflargelcpvalues
        h4 = (C.c_uint64 * 4)()
        if want_suf and want_lcp and want_bck:
            ck(lib.gtb_group_hash_results(g, h4))
            res.device_hashes = {"suf": h4[0], "lcp": h4[1], "llv": h4[2], "bck": h4[3]}
        return res
    finally:
        lib.gtb_group_delete(g)


def build_esa(enc: EncodedSequence, prefixlength=None, device=0, parts=1, want_suf=True, want_lcp=True,
              want_bck=True, filler=None, want_bwt=False, readmode="fwd", devices=None, protocol=None):
    """One call: encoded sequence -> EsaResult (the public entry the benchmarks time end to end).
    devices: a list of CUDA devices, one code range on each (times `parts`).  protocol="exchange"
    keeps the request/answer rank exchange of genometools_b200/multirange.py for -parts on one GPU
    (the single-GPU twin of the NCCL protocol); the default for several ranges is the gtb_group path."""
    if prefixlength is None or prefixlength == 0:
        prefixlength = recommendedprefixlength(enc.numofchars, enc.totallength)
    else:
        maxpl = whatisthemaximalprefixlength(enc.numofchars, enc.totallength)
        if prefixlength > maxpl:
            raise GtbError(f"prefix length {prefixlength} is too large, maximal prefix length for this input "
                           f"size and alphabet size is {maxpl}")          # gt_checkprefixlength, sfx-apfxlen.c:149
    if protocol is None:
        protocol = os.environ.get("GTB200_PARTS_PROTOCOL", "group")
    if devices is not None or (parts > 1 and prefixlength >= 1 and protocol != "exchange"):
        devs = [d for d in (devices if devices is not None else [device]) for _ in range(max(parts, 1))]
        if len(devs) > 1:
            return build_esa_group(enc, prefixlength, devs, want_suf, want_lcp, want_bck, filler, want_bwt, readmode)
        device = devs[0]
    with Suffixerator(device) as sfx:
        sfx.set_sequence(enc, filler, readmode)
        return sfx.run(prefixlength, want_suf, want_lcp, want_bck, parts, want_bwt=want_bwt)


def suffixerator_main(argv, out=sys.stdout):
    """`gt suffixerator` for the accelerated option set. Returns the exit code."""
    try:
        o = SuffixeratorOptions.parse(list(argv))
        alphabet = "protein" if o.protein else "dna"
        if not o.dna and not o.protein:
            raise GtbError("one of the options -dna or -protein is required (alphabet guessing is not on this path)")
        enc = encode_fasta(o.db, alphabet)
        if o.tis:
            try:
                write_index_files(o.db, o.indexname, des=o.des, sds=o.sds and o.des, ssp=o.ssp, md5=o.md5,
                                  alphabet=alphabet)
            except FastaUnsupported as e:
                raise GtbError(f"the index files of this input need the reference's encoder: {e}")
        pl = o.pl
        res = build_esa(enc, pl if pl else None, o.device, o.parts, o.suf, o.lcp, o.bck or pl is not None,
                        want_bwt=o.bwt, readmode=o.dir)
        if o.bwt:
            with open(o.indexname + ".bwt", "wb") as fh:
                fh.write(res.bwt_bytes())
        if o.suf:
            with open(o.indexname + ".suf", "wb") as fh:
                fh.write(res.suf_bytes())
        if o.lcp:
            with open(o.indexname + ".lcp", "wb") as fh:
                fh.write(res.lcp_bytes())
            with open(o.indexname + ".llv", "wb") as fh:
                fh.write(res.llv_bytes())
        if o.bck:
            with open(o.indexname + ".bck", "wb") as fh:
                fh.write(res.bck_bytes())
        with open(o.indexname + ".prj", "w") as fh:
            fh.write(res.prj_text(enc.specialcharinfo(), enc.numofsequences, with_lcp=o.lcp))
        if o.verbose:
            for st in res.stats:
                print("# " + " ".join(f"{k}={v}" for k, v in st.items()), file=out)
        return 0
    except (GtbError, ValueError, OSError) as e:
        print(f"gt suffixerator: error: {e}", file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(suffixerator_main(sys.argv[1:]))
