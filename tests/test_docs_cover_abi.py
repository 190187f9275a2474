"""INTEGRATION.md must name every entry point include/grace_b200.h declares (CPU test)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_md_lists_every_c_entry_point():
    header = open(os.path.join(ROOT, "include", "grace_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = sorted(set(re.findall(r"\b(grace_b200_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 40
    # the table abbreviates families: `grace_b200_create/destroy/...`, `_keys63_f4`, `_xor32/64`
    words = set(re.findall(r"[a-z0-9_]+", doc))
    missing = []
    for n in names:
        tail = n[len("grace_b200_"):]
        parts = tail.split("_")
        ok = n in doc or ("_" + tail) in doc or tail in words or any(
            "_".join(parts[k:]) in words or ("_" + "_".join(parts[k:])) in doc for k in range(1, len(parts)))
        # numeric alternatives such as u32/u64, xor32/64
        ok = ok or re.sub(r"\d+$", "", tail) in doc
        if not ok:
            missing.append(n)
    assert not missing, missing
