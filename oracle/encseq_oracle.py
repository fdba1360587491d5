"""TEST INFRASTRUCTURE -- a plain sequential restatement of the reference's FASTA -> encseq encoder, the way the
reference does it: one character at a time, two passes.  Only tests/ may import it; the product
(genometools_b200/csrc/gtb_fasta.cpp) is organised differently (parallel passes over chunks, tables built from run
lists) and shares no code with it.  Pinned: tests/test_fasta_encseq.py checks it against the md5 sums of the files the
unmodified reference wrote (tests/golden/fasta_index_md5.json) before it is used to check the library on random inputs.

What it follows (paths under /root/reference/src/core):
  reader        gt_sequence_buffer_fasta_advance, sequence_buffer_fasta.c:41-165; process_char,
                sequence_buffer_inline.h:26-58
  pass 1        gt_inputfiles2sequencekeyvalues, encseq.c:5421-5673 with encseq_charproc.gen
  sizes         gt_encseq_determine_size :5149-5213, gt_encseq_sizeofSWtable :923-950, doupdatesumranges :5215-5256,
                determinesmallestrep / gt_encseq_access_type_determine, encseq_access_type.c:96-251,
                determineoptimalsssptablerep, encseq.c:1714-1736
  pass 2        fillSWtable_*, accspecialrange.gen:29-262; fillViaequallength, encseq.c:2521-2640; fillViabitaccess
                :2738-2860; fillViabytecompress :2324-2440; ssptaboutinfo_*, :1841-1910
  files         gt_encseq_assign_header_mapspec :1288-1307, gt_encseq_assign_sequence_mapspec :1346-1402,
                addswtabletomapspectable :829-897, gt_mapspec_write, mapspec.c:366-466
"""
import hashlib
import struct

WILDCARD, SEPARATOR, UNDEF = 254, 255, 253
DNA = ("acgt", "nsywrkvbdhmNSYWRKVBDHM", "n")
PROTEIN = ("LVIFKREDAGSTNQYWPHMC", "XUBZJO*-", "X")
SAT = {"direct": 0, "bytecompress": 1, "eqlen": 2, "bit": 3, "uchar": 4, "ushort": 5, "uint32": 6}
MAXRANGE = {4: 0xff, 5: 0xffff, 6: 0xffffffff}
WIDTH = {4: 1, 5: 2, 6: 4}
SPACE = b" \t\n\v\f\r"


class Declined(Exception):
    """the reference reports an error for this input (or the case is outside both implementations)"""


def symbolmap(alphabet):
    chars, wild, _ = DNA if alphabet == "dna" else PROTEIN
    m = [UNDEF] * 256
    for i, ch in enumerate(chars):
        m[ord(ch)] = i
        if alphabet == "dna":
            m[ord(ch.upper())] = i
    if alphabet == "dna":
        m[ord("u")] = m[ord("U")] = 3
    for ch in wild:
        m[ord(ch)] = WILDCARD
    return m


def read_fasta(files, smap):
    """the reader: symbols (codes, 254, 255), original characters, descriptions, (bytes, effective length) per file"""
    codes, orig, descs, flv = [], [], [], []
    first_overall = True
    for raw in files:
        if not raw or raw[:1] != b">":
            raise Declined("a file that does not begin with '>'")
        indesc, first_in_file, added, desc = False, True, 0, None
        for c in raw:
            if indesc:
                if c == 10:
                    indesc = False
                    descs.append(bytes(desc))
                elif c != 13:
                    desc.append(c)
            elif c in SPACE:
                pass
            elif c == ord(">"):
                if first_overall:
                    first_overall = first_in_file = False
                else:
                    if first_in_file:
                        first_in_file = False
                    else:
                        added += 1
                    codes.append(SEPARATOR)
                    orig.append(0)
                indesc, desc = True, bytearray()
            else:
                if smap[c] == UNDEF or c >= 128:
                    raise Declined("illegal character")
                codes.append(smap[c])
                orig.append(c)
                added += 1
        if indesc:
            raise Declined("the file ends inside a description")
        flv.append((len(raw), added))
    return codes, orig, descs, flv


def ranges_tab(lengths):
    """currentspecialrangevalue, encseq.c:5061-5074: table entries for runs of these lengths, per entry width"""
    tab = []
    for maxv in (0xff, 0xffff):
        tab.append(sum(1 if ln <= maxv + 1 else -(-ln // (maxv + 1)) for ln in lengths))
    tab.append(len(lengths))
    return tab


def size_swtable(sat, withlen, n, items):
    if items == 0:
        return 0
    return (2 if withlen else 1) * WIDTH[sat] * items + 8 * (n // MAXRANGE[sat] + 1)


def units_twobit(n):
    return 2 if n < 32 else 2 + (n - 1) // 32


def pad8(b):
    return b + b"\0" * (-len(b) % 8)


def field(fmt, values):
    if not values:
        return b""
    return pad8(struct.pack("<%d%s" % (len(values), fmt), *values))


def swtable(starts_and_lengths, sat, n, withlen):
    """fillSWtable: an entry per run, a run longer than a page's worth continues in a new entry; endidxinpage[p] =
    entries that began at or before the last position of page p"""
    maxv = MAXRANGE[sat]
    positions, lengths = [], []
    npages = n // maxv + 1
    endidx = [0] * npages
    for start, ln in starts_and_lengths:
        while ln > 0:
            take = min(ln, maxv + 1) if withlen else 1
            positions.append(start & maxv)
            lengths.append(take - 1)
            endidx[start // (maxv + 1)] += 1
            start += take
            ln -= take
    for p in range(1, npages):
        endidx[p] += endidx[p - 1]
    fmt = {1: "B", 2: "H", 4: "I"}[WIDTH[sat]]
    if not positions:
        return b""
    out = field(fmt, positions)
    if withlen:
        out += field(fmt, lengths)
    return out + field("Q", endidx)


def encode(files, names, alphabet="dna", des=True, sds=True, ssp=True, md5=True, clip_desc=False, sat=None):
    """-> {suffix: bytes} for the files the reference writes"""
    chars, _, wildshow = DNA if alphabet == "dna" else PROTEIN
    K = len(chars)
    forced = sat
    smap = symbolmap(alphabet)
    codes, orig, descs, flv = read_fasta(files, smap)
    n = len(codes)
    if n == 0:
        raise Declined("no symbols")
    # ---- pass 1 (encseq_charproc.gen), one symbol at a time
    sci = dict.fromkeys(("specialcharacters", "specialranges", "realspecialranges", "lengthofspecialprefix",
                         "lengthofspecialsuffix", "wildcards", "wildcardranges", "realwildcardranges",
                         "lengthofwildcardprefix", "lengthofwildcardsuffix", "lengthoflongestnonspecial"), 0)
    special_runs, wild_runs, seqlens, seppos = [], [], [], []
    lastspecial = lastwild = lastnonspecial = curlen = 0
    specialprefix = wildprefix = True
    digests, h = [], hashlib.md5()
    dist = [0] * K
    for pos, cc in enumerate(codes):
        if cc < WILDCARD:
            curlen += 1
            specialprefix = wildprefix = False
            if lastspecial:
                special_runs.append(lastspecial); lastspecial = 0
            if lastwild:
                wild_runs.append(lastwild); lastwild = 0
            lastnonspecial += 1
            dist[cc] += 1
            h.update(chars[cc].upper().encode())
        else:
            if lastnonspecial:
                sci["lengthoflongestnonspecial"] = max(sci["lengthoflongestnonspecial"], lastnonspecial)
                lastnonspecial = 0
            if cc == WILDCARD:
                if wildprefix:
                    sci["lengthofwildcardprefix"] += 1
                lastwild += 1
                sci["wildcards"] += 1
                h.update(wildshow.upper().encode())
                curlen += 1
            else:
                wildprefix = False
                if lastwild:
                    wild_runs.append(lastwild); lastwild = 0
                digests.append(h.hexdigest()); h = hashlib.md5()
                if curlen == 0:
                    raise Declined("an empty sequence")
                seqlens.append(curlen); curlen = 0
                seppos.append(pos)
            if specialprefix:
                sci["lengthofspecialprefix"] += 1
            sci["specialcharacters"] += 1
            lastspecial += 1
    if curlen == 0:
        raise Declined("an empty last sequence")
    seqlens.append(curlen)
    digests.append(h.hexdigest())
    if lastspecial:
        special_runs.append(lastspecial)
    if lastnonspecial:
        sci["lengthoflongestnonspecial"] = max(sci["lengthoflongestnonspecial"], lastnonspecial)
    if lastwild:
        wild_runs.append(lastwild)
    sci["lengthofspecialsuffix"], sci["lengthofwildcardsuffix"] = lastspecial, lastwild
    numseq = len(seqlens)
    equallength = len(set(seqlens)) == 1 and sci["wildcards"] == 0
    # the distinct original characters per code (determine_original_subdist, encseq.c:5270-5359)
    perclass = {}
    for c in set(o for o, cc in zip(orig, codes) if cc != SEPARATOR):
        perclass.setdefault(smap[c], set()).add(c)
    numofallchars = sum(len(v) for v in perclass.values())
    maxsub = max(len(v) for v in perclass.values())
    # ---- representation
    specialtab, wildtab = ranges_tab(special_runs), ranges_tab(wild_runs)
    sci["realspecialranges"], sci["realwildcardranges"] = len(special_runs), len(wild_runs)
    twobit = units_twobit(n) * 8
    smallest = None
    for k, sat in enumerate((4, 5, 6)):
        size = twobit + size_swtable(sat, True, n, wildtab[k])
        if smallest is None or size < smallest:
            smallest = size
            sci["specialranges"], sci["wildcardranges"] = specialtab[k], wildtab[k]
    if forced:                                # -sat (getsatforcevalue, encseq.c:797-814; encseq_access_type.c:164-247)
        if forced == "direct" or (alphabet != "dna" and forced == "bytecompress") or \
                (alphabet == "dna" and (forced == "bit" or (forced == "eqlen" and equallength))):
            sat, items = SAT[forced], wildtab[0]
        elif alphabet == "dna" and forced in ("uchar", "ushort", "uint32"):
            sat = SAT[forced]
            items = wildtab[sat - 4]
            sci["specialranges"], sci["wildcardranges"] = specialtab[sat - 4], wildtab[sat - 4]
        else:
            raise Declined("-sat %s is an error of the reference for this input" % forced)
    elif alphabet != "dna":
        sat, items = SAT["bytecompress"], 0
    elif equallength:
        sat, items = SAT["eqlen"], 0
    else:
        sat, items = SAT["bit"], wildtab[0]
        nbits = n + 64
        cmin = twobit + ((8 * (1 if (nbits >> 6) == 0 else 1 + ((nbits - 1) >> 6)))
                         if (wildtab[0] > 0 or numseq > 1) else 0)
        for k, cand in enumerate((4, 5, 6)):
            size = twobit + size_swtable(cand, True, n, wildtab[k])
            if size < cmin:
                cmin, sat, items = size, cand, wildtab[k]
    satsep = None
    if numseq > 1 and sat != SAT["eqlen"] and (ssp or sat >= 4):
        best = None
        for cand in (4, 5, 6):
            size = size_swtable(cand, False, n, numseq - 1)
            if best is None or size < best:
                best, satsep = size, cand
    # ---- pass 2
    lpc = dist.index(min(dist))
    wild_ranges, start = [], None
    for pos, cc in enumerate(codes + [0]):
        if cc == WILDCARD and pos < n:
            if start is None:
                start = pos
        elif start is not None:
            wild_ranges.append((start, pos - start)); start = None
    if sat == SAT["direct"]:
        body = pad8(bytes(codes))
    elif alphabet == "dna":
        words = [0] * units_twobit(n)
        for pos, cc in enumerate(codes):
            v = cc if cc < 4 else (lpc if sat != SAT["bit"] else (0 if cc == WILDCARD else 1))
            words[pos // 32] |= v << (2 * (31 - pos % 32))
        body = field("Q", words)
        if sat == SAT["bit"] and (wildtab[0] > 0 or numseq > 1):
            nbits = n + 64
            bits = [0] * (1 if (nbits >> 6) == 0 else 1 + ((nbits - 1) >> 6))
            for pos in [p for p, cc in enumerate(codes) if cc >= WILDCARD] + list(range(n, n + 64)):
                bits[pos >> 6] |= 1 << (63 - (pos & 63))
            body += field("Q", bits)
        elif sat >= 4 and items:
            body += swtable(wild_ranges, sat, n, True)
    else:
        acc = 0
        for cc in codes:
            acc = (acc << 5) | (cc if cc < K else (K if cc == WILDCARD else K + 1))
        nbytes = (5 * n + 7) // 8
        body = pad8((acc << (8 * nbytes - 5 * n)).to_bytes(nbytes, "big"))
    # ---- files
    namebytes = b"".join(nm.encode() + b"\0" for nm in names)
    order = ("specialcharacters", "specialranges", "realspecialranges", "lengthofspecialprefix", "lengthofspecialsuffix",
             "wildcards", "wildcardranges", "realwildcardranges", "lengthofwildcardprefix", "lengthofwildcardsuffix",
             "lengthoflongestnonspecial")
    esq = (field("B", [1]) + field("Q", [3, sat, n, numseq, len(files), len(namebytes)]) +
           field("Q", [sci[k] for k in order] + [0, 0, 0]) + field("Q", [min(seqlens), max(seqlens)]) +
           field("Q", [0 if alphabet == "dna" else 1, 0]) + pad8(namebytes) + field("B", [maxsub]) +
           field("Q", [numofallchars]) + field("Q", [x for pair in flv for x in pair]) + field("Q", dist) + body)
    out = {"esq": esq}
    if satsep is not None:
        out["ssp"] = swtable([(p, 1) for p in seppos], satsep, n, False)
    if des:
        text, ends, longest = b"", [], 0
        for i, d in enumerate(descs):
            if clip_desc:
                for j, c in enumerate(d):
                    if c in SPACE:
                        d = d[:j]
                        break
            longest = max(longest, len(d))
            text += d
            if i + 1 < numseq:
                ends.append(len(text))
            text += b"\n"
        out["des"] = text + struct.pack("<QQ", longest, 2 ** 64 - 1)
        if sds:
            out["sds"] = struct.pack("<%dQ" % len(ends), *ends)
    elif sds:
        out["sds"] = b""
    if md5:
        out["md5"] = b"".join(d.encode() + b"\0" for d in digests)
    return out
