"""CPU-only: the C-ABI library loads without a GPU and exports every declared symbol."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


HEADERS = ("b200_rmsd.h", "b200_rmsd_debug.h")   # the drop-in boundary; test hooks and probes


def declared_symbols(header=None):
    syms = set()
    for h in ([header] if header else HEADERS):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        syms |= set(re.findall(r"\b(b200_[a-z0-9_A-Z]+)\s*\(", src))
    return sorted(syms)


def test_product_header_has_no_debug_exports():
    prod = declared_symbols("b200_rmsd.h")
    assert not [s for s in prod if "debug" in s or "measure" in s], prod
    assert "b200_rmsd_1vN_set_ref" in prod and "b200_set_fixed_point_bits" in prod


def test_header_symbols_exported(built):
    import cpptraj_b200 as b
    L = ctypes.CDLL(b.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "missing export %s" % s


def test_no_torch_types_in_header():
    for h in HEADERS:
        src = open(os.path.join(ROOT, "include", h)).read()
        code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)     # declarations only, comments stripped
        assert "torch" not in code.lower() and "at::" not in code and "std::" not in code and "&" not in code


def test_library_does_not_link_oracle(built):
    """The product must not route through the CPU checker."""
    import subprocess
    import cpptraj_b200 as b
    out = subprocess.run(["ldd", b.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in out and "cpptraj_ref" not in out
    blob = open(b.LIB_PATH, "rb").read()
    assert b"orc_rms2d" not in blob and b"ref_rms2d" not in blob


def test_shard_rows_partition_and_balance(built):
    import cpptraj_b200 as b
    for n in (0, 1, 2, 31, 32, 33, 1000, 10000, 100000):
        for cnt in (1, 2, 4, 8):
            prev = 0
            pairs = []
            for r in range(cnt):
                a, e = b.shard_rows(n, r, cnt)
                assert a == prev and e >= a and (a % 32 == 0 or a == n) and (e % 32 == 0 or e == n)
                prev = e
                pairs.append(((n - 1 - a) + (n - e)) * (e - a) / 2 if e > a else 0)
            assert prev == n
            if n >= 10000:
                assert max(pairs) / (sum(pairs) / cnt) < 1.12, (n, cnt, pairs)


def test_shard_query_without_device(built):
    import cpptraj_b200 as b
    crd = np.zeros((100, 30), np.float32)
    sel = np.arange(10, dtype=np.int32)
    tot = 0
    first_expected = 0
    for r in range(4):
        _, first, n = b.rms2d_tri_shard(crd, sel, r, 4, query_only=True)
        assert first == first_expected
        first_expected += n
        tot += n
    assert tot == b.tri_size(100)


def test_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cpptraj_b200 as b
    with pytest.raises(b.B200Error) as ei:
        b.init(1)
    assert ei.value.code == 1
    with pytest.raises(b.B200Error):
        b.rms2d_tri(np.zeros((4, 9), np.float32), np.arange(3, dtype=np.int32))


def test_cpptraj_host_glue_compiles_against_reference():
    """cpptraj_host/ (C++ glue + reference.patch) applies to the reference tree and compiles against its headers
    (-fsyntax-only, -DCUDA_B200).  Needs /root/reference: skipped on the GPU box."""
    import subprocess
    ref = os.environ.get("CPPTRAJ_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "src")):
        pytest.skip("reference tree not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["bash", os.path.join(root, "tools", "build_cpptraj_b200.sh"), "--check", ref],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("syntax ok") == 18   # the glue, sixteen patched reference sources, configure


def test_every_cpptraj_deck_has_its_goldens():
    """tests/cpptraj_decks.py and tests/golden/cpptraj/ (outputs of the UNMODIFIED reference, tools/make_golden_cpptraj.py)
    stay in step: every output a deck lists has a non-empty golden, and no golden directory is orphaned."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    from cpptraj_decks import DECKS
    gold = os.path.join(here, "golden", "cpptraj")
    for name, (text, outs) in DECKS.items():
        assert "{D}" in text and outs, name
        for fname, kind in outs:
            p = os.path.join(gold, name, fname)
            assert os.path.isfile(p) and os.path.getsize(p) > 0, "missing golden %s/%s" % (name, fname)
            assert kind in ("table", "crd", "text", "cmatrix"), (name, fname, kind)
    assert sorted(os.listdir(gold)) == sorted(DECKS), "golden directories and decks differ"
