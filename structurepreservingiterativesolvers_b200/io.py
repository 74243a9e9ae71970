"""CSR system export/import (.npz) so that systems assembled elsewhere (e.g. by a real Firedrake
installation: `M.handle.getValuesCSR()`, lkdv/lkdv.py:109-111) can be replayed bit-exactly here.

save_system(path, dic) stores every scipy-sparse entry of a `linforms`-style dictionary as its three
CSR arrays, every ndarray as is, and every scalar as a 0-d array; load_system(path) rebuilds it.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps

_SEP = "::"


def save_system(path, dic):
    out = {}
    for key, val in dic.items():
        if sps.issparse(val):
            m = val.tocsr()
            out[f"{key}{_SEP}indptr"] = m.indptr
            out[f"{key}{_SEP}indices"] = m.indices
            out[f"{key}{_SEP}data"] = m.data
            out[f"{key}{_SEP}shape"] = np.asarray(m.shape, dtype=np.int64)
        else:
            out[key] = np.asarray(val)
    np.savez_compressed(path, **out)


def load_system(path):
    dic = {}
    with np.load(path) as data:
        mats = {}
        for name in data.files:
            if _SEP in name:
                key, part = name.split(_SEP)
                mats.setdefault(key, {})[part] = data[name]
            else:
                arr = data[name]
                dic[name] = arr.item() if arr.ndim == 0 else arr
        for key, p in mats.items():
            dic[key] = sps.csr_matrix((p["data"], p["indices"], p["indptr"]), shape=tuple(p["shape"]))
    return dic
