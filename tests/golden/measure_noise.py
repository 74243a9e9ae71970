"""Measure the reference algorithm's self-noise (oracle on a symmetrically permuted system vs golden)."""
import sys, os; sys.path[:0]=[os.path.dirname(os.path.abspath(__file__)), os.path.join(os.path.dirname(os.path.abspath(__file__)),'..'), os.path.join(os.path.dirname(os.path.abspath(__file__)),'..','..')]
import numpy as np, warnings, scipy.sparse as sps
import helpers, cases
from oracle import cgmres_oracle as orc
from structurepreservingiterativesolvers_b200 import wrappers
warnings.simplefilter('ignore')
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)),'reference_outputs.npz'))
class Q:  # permuted class-form constraint
    def __init__(s, M, v, c): s.M, s.v, s.c = M, v, c
for name in cases.CASES:
    spec, dic, prob, x0, pre = cases.instantiate(name)
    if spec['exp']=='lkdvRK' or spec.get('pre')=='ilu': continue
    n = dic['b'].size
    rng = np.random.default_rng(1); p = rng.permutation(n)
    P = sps.csr_matrix((np.ones(n), (np.arange(n), p)), shape=(n,n))   # (Px)_i = x_{p_i}
    A = (P@dic['A']@P.T).tocsr(); b = P@dic['b']; x0p = P@x0
    wrap = getattr(wrappers, spec['exp'])
    prep = None if pre is None else (P@pre@P.T)
    if spec['kind']=='gmres':
        x, info = orc.fgmres(A,b,x0p,spec['k'],tol=spec['tol'],pre=prep)
    else:
        cl = [Q((P@c.M@P.T).tocsr(), P@np.asarray(c.v).reshape(-1), c.c) for c in wrap.conlist(dic, x0)]
        proto = (spec["tol"] <= 1e-20) if spec["exp"] in ("lkdv",) else (spec["tol"] < 1e-20)
        if proto: x, info = orc.cgmres_prototype(A,b,x0p,spec['k'],conlist=cl,pre=prep)
        else:
            kw = {'contol':spec['contol']} if 'contol' in spec else {}
            x, info = orc.cgmres(A,b,x0p,spec['k'],tol=spec['tol'],conlist=cl,pre=prep,**kw)
    xu = P.T@x
    print(f"{name:24s} oracle self-noise under symmetric permutation: {helpers.rel_diff(xu, g[name+'/x_last']):.1e}  steps {info.get('steps')} vs {int(g[name+'/steps'])}")
