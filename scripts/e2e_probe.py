"""Where does the host-buffer call prf_hashrf spend its wall time?  (run with PRF_TRACE=1)"""
import ctypes as C, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from phybin_b200.host import Forest
from phybin_b200.rfdistance import RFContext
from phybin_b200._native import prf_stats
T, n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, 200
f = Forest.synthetic(T, n, 0x5EED0003)
bips, off = f.all_bips()
P = T * (T - 1) // 2
ctx = RFContext(0)
pin_in = torch.from_numpy(bips.view(np.int64)).pin_memory()
pin_out = torch.empty(P, dtype=torch.int16).pin_memory()
st = prf_stats()
for i in range(4):
    sys.stderr.write(f"---- call {i}\n")
    t0 = time.perf_counter()
    rc = ctx._L.prf_hashrf(ctx._h, pin_in.data_ptr(), off.ctypes.data, T, bips.shape[1], -1, pin_out.data_ptr(), None, C.byref(st))
    torch.cuda.synchronize()
    sys.stderr.write(f"call {i}: rc={rc} wall {1e3*(time.perf_counter()-t0):.1f} ms  total_ev {st.ms_total:.1f} h2d {st.ms_h2d:.1f} dedup {st.ms_dedup:.2f} inc {st.ms_incidence:.2f} mat {st.ms_matrix:.2f} d2h {st.ms_d2h:.1f}\n")
