#!/usr/bin/env python
"""Golden outputs of tests/cpptraj_decks.py from the UNMODIFIED reference: every deck is run by the plain OpenMP build of
cpptraj (recipe of SURVEY.md 8c; built by tools/make_golden_cluster_sieve.sh into /tmp/cpptraj_plain_build) on the CPU;
the listed output files go to tests/golden/cpptraj/<deck>/.  Only possible where /root/reference exists.
usage: [GOLDEN_THREADS=n] python tools/make_golden_cpptraj.py [plain cpptraj binary] [reference test dir] [deck ...]"""
import os, shutil, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cpptraj_decks import DECKS
BIN = sys.argv[1] if len(sys.argv) > 1 else "/tmp/cpptraj_plain_build/bin/cpptraj.OMP"
D = sys.argv[2] if len(sys.argv) > 2 else "/root/reference/test"
env = dict(os.environ, OMP_NUM_THREADS=os.environ.get("GOLDEN_THREADS", "4"))
only = sys.argv[3:]
for name, (text, outs) in DECKS.items():
    if only and name not in only:
        continue
    w = tempfile.mkdtemp()
    open(os.path.join(w, "in"), "w").write(text.replace("{D}", D))
    r = subprocess.run([BIN, "-i", "in"], cwd=w, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 or "Error" in r.stdout:
        print(r.stdout[-3000:])
        raise SystemExit("deck %s failed" % name)
    dst = os.path.join(ROOT, "tests", "golden", "cpptraj", name)
    os.makedirs(dst, exist_ok=True)
    for f, _ in outs:
        shutil.copy(os.path.join(w, f), os.path.join(dst, f))
    print(name, [f for f, _ in outs])
    shutil.rmtree(w)
