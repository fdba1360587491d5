import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Golden:
    """tests/golden/reference_vectors.npz: outputs of the unmodified reference"""

    def __init__(self):
        self.z = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"), allow_pickle=False)
        self.cases = [str(c) for c in self.z["__cases__"]]

    def get(self, case, key):
        return self.z[f"{case}/{key}"]

    def has(self, case, key):
        return f"{case}/{key}" in self.z.files

    def symbols(self, case):
        if self.has(case, "symbols"):
            return self.get(case, "symbols")
        import synth
        return synth.SYNTH_CASES[case.split("/")[1]][0]()

    def meta(self, case):
        return {"numofchars": int(self.get(case, "numofchars")), "prefixlength": int(self.get(case, "prefixlength")),
                "numofsequences": int(self.get(case, "numofsequences")), "alphabet": str(self.get(case, "alphabet"))}

    def prj(self, case):
        text = bytes(self.get(case, "prj")).decode()
        return dict(line.split("=", 1) for line in text.strip().split("\n")), text


_G = None


def golden_obj():
    global _G
    if _G is None:
        _G = Golden()
    return _G


@pytest.fixture(scope="session")
def golden():
    return golden_obj()


def golden_cases(small_only=False):
    g = golden_obj()
    if small_only:
        return [c for c in g.cases if g.has(c, "suf")]
    return list(g.cases)
