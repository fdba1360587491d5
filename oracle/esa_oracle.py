"""oracle/esa_oracle.py -- TEST INFRASTRUCTURE ONLY.

Python face of the CPU restatement (oracle/esa_oracle.c) plus an independent
FASTA reader and independent writers of the .suf/.lcp/.llv/.bck file images, so
that the parity tests never compare the product with itself.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Parity status: pinned against reference outputs, see
tests/golden/README.md.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle_esa.so")
GTREF = os.path.join(HERE, "_ref", "gtref")
WILDCARD, SEPARATOR = 254, 255

DNA = {**{c: i for i, c in enumerate("acgt")}, **{c: i for i, c in enumerate("ACGT")}, "u": 3, "U": 3}
for _c in "nsywrkvbdhmNSYWRKVBDHM":      # /root/reference/src/core/alphabet.c:84
    DNA[_c] = WILDCARD
PROTEIN = {c: i for i, c in enumerate("LVIFKREDAGSTNQYWPHMC")}   # alphabet.c:87
for _c in "XUBZJO*-":                                            # alphabet.c:90
    PROTEIN[_c] = WILDCARD


class Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("totallength", "specialcharacters", "numofallcodes",
                                          "numofspecialcodes", "numofdistpfxidx", "longest",
                                          "numoflargelcpvalues", "maxbranchdepth")] + [("lcptabsum", C.c_double)]


def build_lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "esa_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_lib())
        _lib.esa_oracle_build.restype = C.c_int
        _lib.esa_oracle_build.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_uint] + [C.c_void_p] * 5 + [C.POINTER(Stats)]
        _lib.esa_oracle_bck_sizes.argtypes = [C.c_uint, C.c_uint] + [C.POINTER(C.c_uint64)] * 3
    return _lib


def read_fasta(paths, alphabet):
    """plain-Python FASTA reader: records joined by one SEPARATOR (also across files)"""
    table = DNA if alphabet == "dna" else PROTEIN
    if isinstance(paths, str):
        paths = [paths]
    out, nseq = bytearray(), 0
    for p in paths:
        with open(p) as fh:
            for line in fh:
                if line.startswith(">"):
                    if nseq > 0:
                        out.append(SEPARATOR)
                    nseq += 1
                else:
                    for ch in line.strip():
                        if ch in " \t\r":
                            continue
                        out.append(table[ch])
    return np.frombuffer(bytes(out), dtype=np.uint8).copy(), nseq


def esa(symbols, numofchars, prefixlength):
    """returns dict(suf, lcp (exact, uint64), leftborder, countspecialcodes, distpfxidx, stats...)"""
    L = lib()
    s = np.ascontiguousarray(symbols, dtype=np.uint8)
    n = s.shape[0]
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    L.esa_oracle_bck_sizes(numofchars, prefixlength, C.byref(a), C.byref(b), C.byref(c))
    suf = np.zeros(n + 1, dtype=np.uint64)
    lcp = np.zeros(n + 1, dtype=np.uint64)
    lb = np.zeros(a.value + 1, dtype=np.uint64)
    csc = np.zeros(max(b.value, 1), dtype=np.uint64)
    dist = np.zeros(max(c.value, 1), dtype=np.uint64)
    st = Stats()
    rc = L.esa_oracle_build(s.ctypes.data, n, numofchars, prefixlength, suf.ctypes.data, lcp.ctypes.data,
                            lb.ctypes.data, csc.ctypes.data, dist.ctypes.data, C.byref(st))
    if rc != 0:
        raise RuntimeError("esa_oracle_build failed")
    return {"suf": suf, "lcp": lcp, "leftborder": lb, "countspecialcodes": csc[: b.value],
            "distpfxidx": dist[: c.value], "longest": st.longest,
            "numoflargelcpvalues": st.numoflargelcpvalues, "maxbranchdepth": st.maxbranchdepth,
            "lcptabsum": st.lcptabsum, "specialcharacters": st.specialcharacters, "totallength": n,
            "prefixlength": prefixlength, "numofchars": numofchars}


def apply_readmode(symbols, mode):
    """the sequence as gt_encseq_get_encoded_char(encseq, i, readmode) reads it
    (/root/reference/src/core/encseq.c:6094-6140): rev/rcl mirror the positions, cpl/rcl
    complement the regular DNA symbols (a<->t, c<->g = 3 - c); specials are kept.  Verified
    against the reference: its -dir output equals its forward output on this sequence
    (tests/golden/readmode_vectors.npz)."""
    s = np.array(symbols, dtype=np.uint8, copy=True)
    if mode in ("rev", "rcl"):
        s = s[::-1].copy()
    if mode in ("cpl", "rcl"):
        reg = s < 4
        s[reg] = 3 - s[reg]
    elif mode not in ("fwd", "rev"):
        raise ValueError("unknown readmode, must be fwd or rev or cpl or rcl")
    return s


def file_images(o):
    """the byte images of .suf .lcp .llv .bck implied by an oracle result
    (formats: SURVEY.md appendix A; u32 bucket table since n+1 <= UINT_MAX)"""
    lcp = o["lcp"]
    small = np.minimum(lcp, 255).astype(np.uint8)
    big = np.flatnonzero(lcp >= 255)
    llv = np.stack([big.astype(np.uint64), lcp[big]], axis=1) if big.size else np.zeros((0, 2), np.uint64)
    bck = b""
    for t in (o["leftborder"], o["countspecialcodes"], o["distpfxidx"]):
        raw = t.astype("<u4").tobytes()
        bck += raw + b"\0" * (-len(raw) % 8)
    return {"suf": o["suf"].astype("<u8").tobytes(), "lcp": small.tobytes(),
            "llv": llv.astype("<u8").tobytes(), "bck": bck}


def bwt_image(o, symbols):
    """the byte image of .bwt (bwttab2file, /root/reference/src/match/sfx-run.c:173-210): the
    encoded symbol before each suffix, UNDEFBWTCHAR (254) for the suffix starting at 0"""
    suf = o["suf"].astype(np.int64)
    sym = np.asarray(symbols, dtype=np.uint8)
    out = np.full(suf.shape[0], 254, dtype=np.uint8)
    nz = suf > 0
    out[nz] = sym[suf[nz] - 1]
    return out.tobytes()


def prj_sorter_lines(o):
    """the lines of the .prj file the sorter is responsible for (sfx-outprj.c:66-77)"""
    n = o["totallength"]
    return [f"numberofallsortedsuffixes={n + 1}", f"longest={o['longest']}",
            f"prefixlength={o['prefixlength']}", f"largelcpvalues={o['numoflargelcpvalues']}",
            "averagelcp=%.2f" % (o["lcptabsum"] / (n + 1)), f"maxbranchdepth={o['maxbranchdepth']}"]


def have_reference():
    return os.path.exists(GTREF) and os.access(GTREF, os.X_OK)


def run_reference(fasta_paths, workdir, alphabet="dna", pl=None, parts=1, indexname="ref", extra=()):
    """run the unmodified reference (oracle/_ref/gtref) and return its file images"""
    if isinstance(fasta_paths, str):
        fasta_paths = [fasta_paths]
    cmd = [GTREF, "suffixerator", "-" + alphabet, "-suf", "-lcp", "-bck"]
    cmd += ["-pl"] + ([str(pl)] if pl else [])
    if parts > 1:
        cmd += ["-parts", str(parts)]
    cmd += list(extra) + ["-indexname", os.path.join(workdir, indexname), "-db"] + list(fasta_paths)
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    out = {}
    for ext in ("suf", "lcp", "llv", "bck", "prj"):
        with open(os.path.join(workdir, indexname + "." + ext), "rb") as fh:
            out[ext] = fh.read()
    if "-bwt" in extra:
        with open(os.path.join(workdir, indexname + ".bwt"), "rb") as fh:
            out["bwt"] = fh.read()
    return out
