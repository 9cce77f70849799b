"""Two executors for the same plans: the CPU oracle and the CUDA library (through its C ABI)."""
import numpy as np

import dfdb_b200 as D


def _norm_col(c):
    if hasattr(c, "tolist") and not isinstance(c, (np.ndarray, np.ma.MaskedArray)):
        return c.tolist()                       # FlatStrings / FlatStringsVector
    if isinstance(c, np.ma.MaskedArray):
        return [None if m else v for v, m in zip(c.data.tolist(), np.ma.getmaskarray(c).tolist())]
    if isinstance(c, tuple):                    # oracle nullable: (values, missing)
        return [None if m else v for v, m in zip(c[0].tolist(), c[1].tolist())]
    return c.tolist()


class OracleEngine:
    name = "oracle"

    def __init__(self, oracle_module, path):
        self.O = oracle_module
        self.ot = oracle_module.OracleTable(path)

    def nrow(self, v):
        return self.ot.count(D.plan_bytes(v))

    def materialize(self, v):
        if isinstance(v, D.DFColumn):
            v = v.view
        try:
            cols = self.ot.materialize(D.plan_bytes(v))
        except self.O.OracleError as e:
            if e.code in (4, 5):
                raise D.ArgumentError(e.msg) from None
            raise
        return {n: _norm_col(c) for n, c in zip(v.projection.keys(), cols)}

    def column(self, col):
        return list(self.materialize(col.view).values())[0]

    def mask(self, v):
        return self.ot.mask(D.plan_bytes(v))

    def agg(self, col):
        a = self.ot.aggregate(D.plan_bytes(col), 0)
        a.sum_f64 = a.sum_kahan
        return a

    def close(self):
        self.ot.close()


class GpuEngine:
    name = "gpu"

    def __init__(self, path=None):
        pass

    def nrow(self, v):
        return D.nrow(v)

    def materialize(self, v):
        return D.materialize(v).to_dict()

    def column(self, col):
        return _norm_col(D.materialize(col))

    def mask(self, v):
        return D.selection_mask(v)

    def agg(self, col):
        return D.aggregate(col)

    def close(self):
        pass
