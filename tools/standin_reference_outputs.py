"""Writes the ORACLE's values in the format of tests/golden/reference_outputs.json to the given path, to exercise the
plumbing of tests/test_reference_golden.py (GPAR_REFERENCE_GOLDEN=<path> pytest ...) where Julia is unavailable.
A stand-in: it pins nothing and must never be committed as tests/golden/reference_outputs.json."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_reference_golden as T

def conv(v):
    if isinstance(v, dict):
        return {k: conv(x) for k, x in v.items()}
    if isinstance(v, np.ndarray):
        return v.tolist()
    return v.item() if hasattr(v, "item") else v

out = {k: conv(v) for k, v in T.oracle_entries().items()}
out["versions"] = {"note": "STAND-IN written from the oracle by tools/standin_reference_outputs.py — not reference output"}
out["selfcheck_A_maxabsdiff"] = 0.0
json.dump(out, open(sys.argv[1], "w"))
print("wrote", sys.argv[1], len(out), "entries")
