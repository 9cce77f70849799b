"""debug: decode one rand100 block with the spec decoder and list the words that differ"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dfdb_b200 import _capi
from oracle import oracle as O
import test_gpu_parity as T
_capi.init(0)
L = _capi.lib()
rng = np.random.default_rng(7)
body = rng.integers(1, 101, 65536).astype(np.int64).tobytes()
comp = O.compress_block(body)
for ctas in (4, 5, 6):
    L.dfdb_set_option(b"spec_ctas", ctas)
    T._set_variant("spec")
    got, status = T._gpu_decode([comp], [len(body)])
    T._set_variant(None)
    g = np.frombuffer(got[0], dtype=np.int64); e = np.frombuffer(body, dtype=np.int64)
    bad = np.nonzero(g != e)[0]
    print("ctas", ctas, "status", status, "bad words", len(bad), bad[:40])
    for w in bad[:12]:
        where = np.nonzero(e[:w] == g[w])[0]
        print("  word", w, "got", g[w], "expected", e[w], "got value last seen at words", where[-3:], "expected value at", np.nonzero(e[:w] == e[w])[0][-3:])
