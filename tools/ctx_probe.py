"""Where the start-up time of a fresh process that uses libgtb200.so goes: driver initialisation (cuInit),
primary context, loading the library, creating a handle (stream + first allocations), first sort (module load).
Each variant runs in a fresh python process without torch.  Run on the GPU box: python tools/ctx_probe.py"""
import subprocess, sys, os, time
CHILD = r'''
import ctypes as C, time, os, sys
t0 = time.perf_counter()
which = sys.argv[1]
out = []
def stamp(what):
    out.append("%s %.3f" % (what, time.perf_counter() - t0))
if which == "driver":
    cu = C.CDLL("libcuda.so.1")
    rc = cu.cuInit(0); stamp("cuInit rc=%d" % rc)
    dev = C.c_int(); cu.cuDeviceGet(C.byref(dev), 0)
    ctx = C.c_void_p(); rc = cu.cuDevicePrimaryCtxRetain(C.byref(ctx), dev); stamp("primaryCtxRetain rc=%d" % rc)
else:
    sys.path.insert(0, os.getcwd())
    import numpy as np
    stamp("import numpy")
    from genometools_b200 import _lib
    lib = _lib.load(); stamp("dlopen libgtb200")
    buf = C.create_string_buffer(512)
    h = lib.gtb_esa_new(0, buf, 512); stamp("gtb_esa_new")
    n = 1 << 20
    words = np.random.default_rng(1).integers(0, 1 << 63, n // 32 + 2, dtype=np.uint64)
    lib.gtb_esa_set_input_2bit(h, words.ctypes.data_as(C.c_void_p), words.shape[0], n, None, 0); stamp("set_input 1 Mbp")
    lib.gtb_esa_run(h, 8, 7); stamp("first run (module load)")
    lib.gtb_esa_run(h, 8, 7); stamp("second run")
    lib.gtb_esa_delete(h); stamp("delete")
print(which, "; ".join(out), flush=True)
'''
for which in ("driver", "library", "driver", "library"):
    t = time.perf_counter()
    r = subprocess.run([sys.executable, "-c", CHILD, which], capture_output=True, text=True)
    print(r.stdout.strip(), "| process wall %.3f" % (time.perf_counter() - t), r.stderr[-300:])
