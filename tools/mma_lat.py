"""tcgen05 MMA + commit latency table (see b200_debug_i8_mma_latency)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpptraj_b200 as b
b.init(1)
L = b.lib()
L.b200_debug_i8_mma_latency.argtypes = [C.c_int, C.c_int]; L.b200_debug_i8_mma_latency.restype = C.c_double
for ctas in (1, 148):
    for n in (0, 1, 2, 4, 8, 16, 32):
        print("ctas %3d  nMma %2d  cycles %8.0f" % (ctas, n, L.b200_debug_i8_mma_latency(n, ctas)))
