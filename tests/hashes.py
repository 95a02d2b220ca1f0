"""Hash record of a witness (test infrastructure): the form in which the CPU oracle's full-size results are committed
(tests/golden/lhs_*_hashes.json, written by tools/golden_full_size.py) and in which the CUDA path is compared with them."""
import hashlib

import numpy as np


def witness_hashes(digits, carries, fa, fb):
    """digits (n,d) u8 MSD first, carries (d,8) u64, canonical a_k / b_k as (len,4) u64 Montgomery limbs (little endian)"""
    rec = {"digits": hashlib.sha256(np.ascontiguousarray(digits).tobytes()).hexdigest(),
           "carries": hashlib.sha256(np.ascontiguousarray(carries).tobytes()).hexdigest(), "functions": []}
    for a, b in zip(fa, fb):
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(a).tobytes())
        h.update(np.ascontiguousarray(b).tobytes())
        rec["functions"].append({"la": int(len(a)), "lb": int(len(b)), "sha256": h.hexdigest()})
    return rec
