import sys, os, tempfile, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfdb_b200 as D
from dfdb_b200 import _capi
from oracle import oracle as O
O.build()
L = _capi.lib()
tmp = tempfile.mkdtemp()
p = os.path.join(tmp, "f")
O.gen_table(p, "a:Int64:iuniform:1:100;b:Float64:funiform;c:Int64:iseq;d:Int64:iuniform:-5:5", 3_000_017, 65536, 0xDFDB0F5E, 4)
qs = {"sum b": lambda t: t[(t.a > 25) & (t.a <= 75), ["b"]].b, "sum c": lambda t: t[t.a > 50, ["c"]].c, "sum d": lambda t: t[(t.a != 7) & (t.a != 93), ["d"]].d}
for name, mk in qs.items():
    for fused in (1, 0):
        L.dfdb_set_option(b"no_decode_fused", 1 - fused)
        t = D.open_table(p)
        L.dfdb_profile_enable(1); L.dfdb_profile_reset()
        n0 = L.dfdb_kernel_launches()
        r = D.aggregate(mk(t))
        n1 = L.dfdb_kernel_launches()
        ph = {}
        for k in (b"decode", b"consume", b"h2d", b"unpack"):
            ms, ln, by = C.c_double(), C.c_int64(), C.c_int64()
            L.dfdb_profile_get(k, C.byref(ms), C.byref(ln), C.byref(by))
            ph[k.decode()] = (round(ms.value, 3), ln.value)
        print(name, "fused" if fused else "plain", n1 - n0, ph, r.count, r.sum_i64, r.sum_f64)
        t.close()
