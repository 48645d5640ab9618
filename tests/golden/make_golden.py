"""Regenerates tests/golden/planes.json: SHA-256 of the colour / depth / stencil planes that the
UNMODIFIED reference (strict IEEE build, oracle/_ref/libref_strict.so) produces for every case in
tests/cases.py, plus a few scalar known-answer facts.  Run in the build container (needs /root/reference
to have been compiled by `make ref`); the GPU box only consumes the committed JSON.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from cases import CASES, LARGE_CASES, case_id  # noqa: E402
from oracle_loader import load_reference  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = load_reference("strict")
    out = {}
    for c in CASES + LARGE_CASES:
        col, dep, sten, err = ref.render(*c)
        out[case_id(c)] = {
            "color": digest(col), "depth": digest(dep), "stencil": digest(sten), "gl_error": int(err),
            "nonbackground": int((col != col[0, 0]).sum()), "stencil_sum": int(sten.astype(np.int64).sum()),
        }
    (ROOT / "tests" / "golden" / "planes.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
