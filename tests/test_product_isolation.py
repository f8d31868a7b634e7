"""The product package must not reach the CPU oracle or the reference tree (it would void parity claims)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_package_never_imports_oracle_or_reference():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dbgsom_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "/root/reference" in text or "_refshim" in text:
                    bad.append(f)
    assert not bad, bad


def test_gpu_side_entry_points_do_not_read_the_reference_tree():
    for name in ("bench.py", "__graft_entry__.py"):
        text = open(os.path.join(ROOT, name), encoding="utf-8").read()
        assert "/root/reference" not in text and "_refshim" not in text, name
