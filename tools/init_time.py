"""Fixed cost of a fresh process: b200_init (cuInit + primary context + streams) and b200_warmup (device chunk buffers, kernel
loading), timed without torch in the process.  usage: python tools/init_time.py"""
import ctypes as C, os, sys, time
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = C.CDLL(os.environ.get("B200_RMSD_LIB") or os.path.join(here, "cpptraj_b200", "libb200rmsd.so"))
used = C.c_int(0)
t0 = time.perf_counter(); rc = L.b200_init(1, C.byref(used)); t1 = time.perf_counter()
rc2 = L.b200_warmup(); t2 = time.perf_counter()
print("b200_init %.3f s (rc %d, %d device), b200_warmup %.3f s (rc %d)" % (t1 - t0, rc, used.value, t2 - t1, rc2))
