"""Only the cpptraj-binary legs of bench.py (cpptraj.B200 vs the unmodified cpptraj.OMP on the same binpos file)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print(json.dumps(bench.cpptraj_leg(bench.CONFIGS["cfg2"], 3.0), indent=1))
