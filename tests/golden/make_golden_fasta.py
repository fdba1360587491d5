"""Golden vectors for the FASTA -> encseq encoder: md5 of every index file the UNMODIFIED reference
(oracle/_ref/gtref, built from /root/reference by oracle/Makefile) writes for the inputs of
fasta_cases.py.  Run here (the reference is not on the GPU box); writes fasta_index_md5.json.

  python tests/golden/make_golden_fasta.py
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import fasta_cases  # noqa: E402

GTREF = os.path.join(ROOT, "oracle", "_ref", "gtref")
SUFFIXES = ("esq", "ssp", "des", "sds", "md5")


def reference_files(files, opts, workdir):
    names = []
    for name, data in fasta_cases.file_names_and_bytes(files, opts):
        names.append(name)
        with open(os.path.join(workdir, name), "wb") as fh:
            fh.write(data)
    cmd = [GTREF, "suffixerator", "-" + opts.get("alphabet", "dna"), "-tis"]
    for k in ("des", "sds", "ssp", "md5"):
        cmd += ["-" + k, "yes" if opts[k] else "no"]
    if opts["clip_desc"]:
        cmd.append("-clipdesc")
    if opts.get("sat"):
        cmd += ["-sat", opts["sat"]]
    cmd += ["-indexname", "ref", "-db"] + names
    r = subprocess.run(cmd, cwd=workdir, capture_output=True, text=True)
    assert r.returncode == 0, (cmd, r.stderr)
    out = {}
    for s in SUFFIXES:
        p = os.path.join(workdir, "ref." + s)
        if os.path.exists(p):
            with open(p, "rb") as fh:
                data = fh.read()
            out[s] = {"bytes": len(data), "md5": hashlib.md5(data).hexdigest()}
    prj = open(os.path.join(workdir, "ref.prj")).read()
    return out, prj


def main():
    golden = {}
    for name, (files, opts) in fasta_cases.all_cases().items():
        with tempfile.TemporaryDirectory() as d:
            out, prj = reference_files(files, opts, d)
        golden[name] = {"files": out, "prj": prj,
                        "input_md5": [hashlib.md5(f).hexdigest() for f in files]}
    with open(os.path.join(HERE, "fasta_index_md5.json"), "w") as fh:
        json.dump(golden, fh, indent=0, sort_keys=True)
    print(len(golden), "cases")


if __name__ == "__main__":
    main()
