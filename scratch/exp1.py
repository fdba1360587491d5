import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden')
import synth
from genometools_b200 import encode_symbols
from genometools_b200.suffixerator import Suffixerator
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
pl = int(sys.argv[2]) if len(sys.argv) > 2 else 11
t=time.time(); sym = synth.random_dna(n, 42); enc = encode_symbols(sym, 4, 1); enc.twobitencoding(); print("gen+pack", time.time()-t)
with Suffixerator(0) as s:
    t=time.time(); s.set_sequence(enc); print("set_sequence", time.time()-t)
    for it in range(4):
        t=time.time(); r = s.run(pl, copy=False); dt=time.time()-t
        st = r.stats[0]
        print(f"run {it}: wall {dt*1e3:.1f} ms  dev {st['ms_total']:.2f} ms  count {st['ms_count']:.2f} hist {st['ms_hist']:.2f} radix {st['ms_radix']:.2f} analyze {st['ms_analyze']:.2f} dbl {st['ms_doubling']:.2f} lcp {st['ms_lcp']:.2f} tail {st['ms_tail']:.2f} passes {st['radix_passes']} launches {st['kernel_launches']} unres {st['unresolved_after_first_sort']} maxlcp {st['maxbranchdepth']} longest {st['longest']}")
    N = st['nonspecials']
    print("radix GB/s:", 24*st['radix_pairs_moved']/st['ms_radix']/1e6, " Msuf/s:", (n+1)/st['ms_total']/1e3)
    t=time.time(); r = s.run(pl, copy=True); print("run+copy wall", time.time()-t)
