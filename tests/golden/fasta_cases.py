"""FASTA inputs for the encoder in front of the sorter (gtb_fasta_encode, include/gtb200.h).

Every case is generated from its seed: make_golden_fasta.py runs the unmodified reference on it
(`gtref suffixerator -tis ...`) and stores the md5 of each index file in fasta_index_md5.json;
tests/test_fasta_encseq.py generates the same bytes again and compares what the library writes.
The cases are chosen to reach every representation the reference picks for DNA (eqlen, bit, uchar,
ushort, uint32: src/core/encseq_access_type.c:96-129) and the one it picks for protein (bytecompress),
wildcard runs longer than a table page,
totals around the page borders, several files, and the odd FASTA the reference's reader accepts
('>' in the middle of a line, blank lines, white space inside lines, CR LF, empty descriptions).
"""
import random

import numpy as np

WILD = "NnRYkmSWBDHVrywsKM"


def _wrap(seq, width, eol):
    return eol.join(seq[i:i + width] for i in range(0, len(seq), width)) + eol


def small_case(seed):
    """returns (list of file contents as bytes, options dict)"""
    rng = random.Random(seed)
    kind = rng.choice(["plain", "eq", "wild", "wild", "manywild", "lower"])
    nfiles = rng.choice([1, 1, 1, 2, 3])
    files = []
    for _ in range(nfiles):
        nseq = rng.choice([1, 1, 2, 3, 7, 40])
        base = rng.choice([10, 33, 64, 100, 300, 1000, 5000])
        out = []
        for s in range(nseq):
            n = base if kind == "eq" else max(1, int(rng.expovariate(1 / base)))
            q = [rng.choice("ACGT") for _ in range(n)]
            if kind == "wild":
                for _ in range(rng.randint(0, 5)):
                    a = rng.randrange(n)
                    b = min(n, a + rng.choice([1, 1, 2, 5, 30, 300, 700]))
                    for i in range(a, b):
                        q[i] = rng.choice(WILD)
            if kind == "manywild":
                q = ["N" if rng.random() < 0.4 else c for c in q]
            if kind == "lower":
                q = [c.lower() if rng.random() < 0.5 else c for c in q]
            q = "".join(q)
            hdr = ">s%d desc\twith  spaces %d" % (s, rng.randrange(1000))
            if rng.random() < 0.1:
                hdr = ">"
            width = rng.choice([60, 70, 7, 1000000])
            eol = rng.choice(["\n", "\n", "\r\n"])
            lines = [q[i:i + width] for i in range(0, n, width)]
            if rng.random() < 0.2:
                lines = [ln[:len(ln) // 2] + " " + ln[len(ln) // 2:] for ln in lines]
            if rng.random() < 0.2:
                lines.insert(rng.randrange(len(lines) + 1), "")
            out.append(hdr + eol + eol.join(lines) + eol)
        files.append("".join(out).encode())
    opts = {k: rng.random() < 0.8 for k in ("des", "sds", "ssp", "md5")}
    if not opts["des"]:
        opts["sds"] = False
    opts["clip_desc"] = rng.random() < 0.2
    return files, opts


def big_file(seed, nseq, total, nruns, runlens, width=70):
    rng = np.random.default_rng(seed)
    parts = []
    for s in range(nseq):
        n = total // nseq + (int(rng.integers(0, 1000)) if nseq > 1 else 0)
        a = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
        for _ in range(nruns):
            ln = int(rng.choice(runlens))
            st = int(rng.integers(0, max(1, n - ln)))
            a[st:st + ln] = ord("N")
        nlines = (n + width - 1) // width
        text = np.full(n + nlines, 10, dtype=np.uint8)
        idx = np.arange(n)
        text[idx + idx // width] = a
        parts.append(b">seq%d of case %d\n" % (s, seed) + text.tobytes())
    return b"".join(parts)


AMINO = "LVIFKREDAGSTNQYWPHMC"
AMINO_WILD = "XUBZJO*-"


def protein_case(seed):
    """protein: the byte-compressed representation (5 bits per residue, fillViabytecompress, encseq.c:2324-2440)"""
    rng = random.Random(1000 + seed)
    files = []
    for _ in range(rng.choice([1, 1, 2])):
        nseq = rng.choice([1, 1, 2, 3, 7, 40])
        base = rng.choice([1, 7, 8, 9, 33, 100, 1000, 5000])
        equal = rng.random() < 0.2
        out = []
        for s in range(nseq):
            n = base if equal else max(1, int(rng.expovariate(1 / base)))
            q = [rng.choice(AMINO) for _ in range(n)]
            for _ in range(rng.randint(0, 4)):
                a = rng.randrange(n)
                b = min(n, a + rng.choice([1, 1, 2, 30, 300]))
                for i in range(a, b):
                    q[i] = rng.choice(AMINO_WILD)
            q = "".join(q)
            width = rng.choice([60, 7, 100000])
            eol = rng.choice(["\n", "\r\n"])
            out.append(">p%d some protein\n" % s + eol.join(q[i:i + width] for i in range(0, n, width)) + eol)
        files.append("".join(out).encode())
    opts = {k: rng.random() < 0.8 for k in ("des", "sds", "ssp", "md5")}
    if not opts["des"]:
        opts["sds"] = False
    opts["clip_desc"] = False
    opts["alphabet"] = "protein"
    return files, opts


ALL_ON = {"des": True, "sds": True, "ssp": True, "md5": True, "clip_desc": False}

ODD = {
    "header_in_midline": b">a\nACGT>b desc\nACNNT\n\n  AC GT\n>c\r\nA\r\nC\r\n",
    "no_final_newline": b">a\nACGT\n>b\nACG",
    "one_symbol": b">x\nA\n",
    "only_wildcards": b">x\nNNNNNNNNNN\n>y\nnnnn\n",
    "wildcard_at_both_ends": b">x\nNNACGTACGTNN\n>y\nNACGTN\n",
    "separator_between_wildcards": b">x\nACGTNN\n>y\nNNACGT\n",
    "31_32_33": b">x\n" + b"ACGT" * 8 + b"\n>y\n" + b"ACGT" * 7 + b"ACG\n",
    "exactly_32": b">x\n" + b"ACGT" * 8 + b"\n",
    "exactly_64_two_equal": b">x\n" + b"ACGT" * 8 + b"\n>y\n" + b"TTGA" * 8 + b"\n",
    "iupac_and_u": b">rna\nACGUacguRYKMSWBDHVN\n",
    "tabs_and_formfeed": b">a\tb\nAC\tGT\x0cAC\x0bGT\n",
    "long_description": b">" + b"d" * 5000 + b"\nACGT\n>short\nAC\n",
    "gt_inside_description": b">a >b >c\nACGT\n",
}


def big_cases():
    """name -> (list of file contents, options)"""
    c = {}
    c["uint32_single"] = ([big_file(1, 1, 3_000_000, 3, [5, 100])], ALL_ON)
    c["long_runs"] = ([big_file(2, 1, 3_000_000, 3, [70000, 300000])], ALL_ON)
    c["ushort_multi"] = ([big_file(3, 5, 3_000_000, 400, [1, 3, 600])], ALL_ON)
    c["many_sequences"] = ([big_file(4, 300, 2_000_000, 2, [1, 300, 66000])], ALL_ON)
    c["three_files"] = ([big_file(5, 2, 1_000_000, 0, [1]), big_file(6, 1, 500_000, 3, [5, 100]),
                         big_file(7, 5, 700_000, 40, [1, 3, 600])], ALL_ON)
    c["page_edge_65535x3"] = ([big_file(8, 1, 65535 * 3, 1, [65536])], ALL_ON)
    c["page_edge_65536x2"] = ([big_file(9, 1, 65536 * 2, 1, [65537])], ALL_ON)
    c["run_of_256"] = ([big_file(10, 1, 2000, 1, [256])], ALL_ON)
    c["run_of_257"] = ([big_file(11, 1, 2000, 1, [257])], ALL_ON)
    c["one_line_3M"] = ([big_file(13, 1, 3_000_000, 5, [1, 40, 70000], width=1 << 30)], ALL_ON)
    c["one_line_per_sequence"] = ([big_file(14, 7, 2_000_000, 3, [1, 300], width=1 << 30)], ALL_ON)
    c["equal_length_reads"] = ([big_file(12, 1, 100, 0, [1]).replace(b"seq0", b"r") * 2000], ALL_ON)
    return c


def file_names_and_bytes(files, opts):
    """the names the files are written under and what is written: f<i>.fa, or f<i>.fa.gz with the gzip of the text"""
    import bz2
    import gzip
    out = []
    for i, raw in enumerate(files):
        if opts.get("gz") and opts["gz"][i] == "bz2":
            out.append(("f%d.fa.bz2" % i, bz2.compress(raw, 1)))
        elif opts.get("gz") and opts["gz"][i]:
            out.append(("f%d.fa.gz" % i, gzip.compress(raw, compresslevel=1, mtime=0)))
        else:
            out.append(("f%d.fa" % i, raw))
    return out


def all_cases(nsmall=120):
    cases = {}
    for seed in range(nsmall):
        cases["small_%03d" % seed] = small_case(seed)
    for name, raw in ODD.items():
        cases["odd_" + name] = ([raw], ALL_ON)
    cases["odd_clip_leading_space"] = ([b">  lead\nACGT\n>x y\nAC\n"], dict(ALL_ON, clip_desc=True))
    cases["odd_second_file_continues"] = ([b">a\nACGT\n", b">b\nAC\n>c\nGGG\n"], ALL_ON)
    cases.update(big_cases())
    for seed in range(40):
        cases["protein_%02d" % seed] = protein_case(seed)
    # -sat: every representation forced on inputs that would get another one (or the same)
    for i, sat in enumerate(("uchar", "ushort", "uint32", "bit", "direct")):
        for seed in (3, 17, 40 + i):
            files, opts = small_case(seed)
            cases["sat_%s_%03d" % (sat, seed)] = (files, dict(opts, sat=sat))
    cases["sat_eqlen_reads"] = (cases["equal_length_reads"][0], dict(ALL_ON, sat="eqlen"))
    cases["sat_uint32_long_runs"] = (cases["long_runs"][0], dict(ALL_ON, sat="uint32"))
    cases["sat_uchar_long_runs"] = (cases["long_runs"][0], dict(ALL_ON, sat="uchar"))
    for sat in ("bytecompress", "direct"):
        for seed in (2, 9):
            files, opts = protein_case(seed)
            cases["sat_%s_protein_%02d" % (sat, seed)] = (files, dict(opts, sat=sat))
    # .gz / .bz2 files (inflated by zlib / libbz2 in the reference as well as here); "gz": which of the files are
    # compressed (True: gzip, "bz2": bzip2)
    cases["gz_small_010"] = (small_case(10)[0], dict(small_case(10)[1], gz=[True] * len(small_case(10)[0])))
    cases["gz_mixed_three_files"] = (cases["three_files"][0], dict(ALL_ON, gz=[True, False, "bz2"]))
    cases["bz2_small_011"] = (small_case(11)[0], dict(small_case(11)[1], gz=["bz2"] * len(small_case(11)[0])))
    cases["gz_protein_04"] = (protein_case(4)[0], dict(protein_case(4)[1], gz=[True] * len(protein_case(4)[0])))
    rng = np.random.default_rng(77)
    residues = np.frombuffer(AMINO.encode(), dtype=np.uint8)[rng.integers(0, 20, 1_000_003)]
    cases["protein_1M"] = ([b">one protein of a million residues\n" + residues.tobytes() + b"\n>second\nMKV\n"],
                           dict(ALL_ON, alphabet="protein"))
    return cases
