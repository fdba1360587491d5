"""TEST DOUBLE (not product code): a numpy stand-in for one code range that speaks the
RangeWorker protocol of genometools_b200/multirange.py, so that the lock-step driver
(all-to-all of positions / ranks, termination, seam fix-up) can be exercised with
torch.distributed + gloo on CPU tensors, world_size > 1, without a GPU.  It follows the
same key layout as the CUDA path (29 DNA symbols + tail) but is written independently."""
import numpy as np
import torch

from genometools_b200.multirange import RangeWorker

M_SYM = 29


def filled_key(sym, run, p):
    """(29 symbols T-filled) << 6 | (29 - u); None for special positions"""
    u = min(int(run[p]), M_SYM)
    if u == 0:
        return None
    k = 0
    for i in range(M_SYM):
        k = (k << 2) | (int(sym[p + i]) if i < u else 3)
    return (k << 6) | (M_SYM - u)


class CpuRangeWorker(RangeWorker):
    def __init__(self, sym, first_key, next_first_key, sa_offset):
        n = sym.shape[0]
        self.sym, self.n = sym, n
        run = np.zeros(n + 1, dtype=np.int64)
        for i in range(n - 1, -1, -1):
            run[i] = 0 if sym[i] >= 254 else run[i + 1] + 1
        self.run = run
        self.keys = {}
        for p in range(n):
            k = filled_key(sym, run, p)
            if k is not None:
                self.keys[p] = k
        self.mine = [p for p, k in self.keys.items() if first_key <= k and (next_first_key is None or k < next_first_key)]
        self.sa_offset = sa_offset
        self.lcp0 = None

    def sort_begin(self):
        order = sorted(self.mine, key=lambda p: (self.keys[p], p))
        self.sa = order
        self.head = [0] * len(order)
        for j in range(len(order)):
            tied = j > 0 and self.keys[order[j]] == self.keys[order[j - 1]] and (self.keys[order[j]] & 63) == 0
            self.head[j] = self.head[j - 1] if tied else j
        self.round = 0
        self._refresh()

    def _refresh(self):
        from collections import Counter
        size = Counter(self.head)
        self.unres = [j for j in range(len(self.sa)) if size[self.head[j]] > 1]

    def unresolved(self):
        return len(self.unres)

    def ensure_ranks(self):
        S = int((self.sym >= 254).sum())
        self.isa = {self.n: self.n}
        for idx, p in enumerate(np.flatnonzero(self.sym >= 254)):
            self.isa[int(p)] = self.n - S + idx
        for j, p in enumerate(self.sa):
            self.isa[p] = self.sa_offset + self.head[j]

    def round_prepare(self, first_keys, my_range):
        h = M_SYM << self.round
        fk = [int(x) for x in first_keys]
        fk[0] = 0
        self.local, buckets = {}, [[] for _ in fk]
        for j in self.unres:
            q = self.sa[j] + h
            k = self.keys.get(q)
            owner = my_range if k is None else max(r for r in range(len(fk)) if fk[r] <= k)
            if owner == my_range:
                self.local[j] = self.isa[q]
            else:
                buckets[owner].append((q, j))
        self.order = [j for b in buckets for (_q, j) in b]
        send = torch.tensor([q for b in buckets for (q, _j) in b], dtype=torch.int32)
        return send, [len(b) for b in buckets]

    def rank_lookup(self, positions):
        return torch.tensor([self.isa[int(q)] for q in positions.tolist()], dtype=torch.int32)

    def round_finish(self, answers):
        rank = dict(self.local)
        for j, a in zip(self.order, answers.tolist()):
            rank[j] = int(a)
        groups = {}
        for j in self.unres:
            groups.setdefault(self.head[j], []).append(j)
        for g, members in groups.items():
            items = sorted(((rank[j], self.sa[j]) for j in members), key=lambda t: t[0])   # stable
            for off, (r, p) in enumerate(items):
                j = g + off
                self.sa[j] = p
                self.head[j] = self.head[j - 1] if off > 0 and r == items[off - 1][0] else j
                rank_j = self.head[j]
                self.isa[p] = self.sa_offset + rank_j
        self.round += 1
        self._refresh()

    def sort_end(self):
        pass

    def boundary_keys(self):
        if not self.sa:
            return False, 0, 0
        return True, self.keys[self.sa[0]], self.keys[self.sa[-1]]

    def fix_seam(self, prev_last_key):
        a, b = prev_last_key, self.keys[self.sa[0]]
        x = (a ^ b) >> 6
        l = M_SYM if x == 0 else (58 - x.bit_length()) // 2
        self.lcp0 = min(l, M_SYM - (a & 63), M_SYM - (b & 63))
