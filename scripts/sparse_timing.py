"""Cycle breakdown of the sparse matrix kernel's warp roles (needs lib/libphybin_rf_timing.so, built with -DPRF_TIMING)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from phybin_b200 import _native, build
build.RF_SO = os.path.join(build.LIBDIR, "libphybin_rf_timing.so")
from phybin_b200.host import Forest
from phybin_b200.rfdistance import RFContext
T = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 200
editdist = int(sys.argv[4]) if len(sys.argv) > 4 else -1
f = Forest.synthetic(T, n, 0x5EED0003, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bips, off = f.all_bips()
ctx = RFContext(0)
L = ctx._L
ctx.hashrf_load(bips, off)
for it in range(3):
    st = ctx.hashrf_compute(0, ctx.P, editdist=editdist)
grid = C.c_uint32(0)
buf = np.zeros(1024 * 64, dtype=np.uint64)
L.prf_debug_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
L.prf_debug_timing(ctx._h, buf.ctypes.data, buf.size, C.byref(grid))
g = grid.value
t = buf[: g * 64].reshape(g, 64).astype(np.float64)
print("matrix ms", st["ms_matrix"], "grid", g)
def show(name, w):
    wait, work = t[:, 2 * w], t[:, 2 * w + 1]
    tot = wait + work
    print(f"{name:12s} total {tot.mean()/1e6:7.3f} Mcyc  wait {100*wait.mean()/tot.mean():5.1f}%  work {100*work.mean()/tot.mean():5.1f}%  (min/max total {tot.min()/1e6:.3f}/{tot.max()/1e6:.3f})")
show("producer", 0)
print(f"emitter      wait_done {t[:,2].mean()/1e6:.3f}  emit {t[:,3].mean()/1e6:.3f}  wait_read {t[:,63].mean()/1e6:.3f} Mcyc")
nw = (t[0, 4:62:2] + t[0, 5:62:2] > 0).sum()
for w in range(2, 2 + nw):
    show(f"walker {w-2}", w)
