"""Fit of the exp2-of-polynomial GELU used by gdfn_math.cuh (gelu_gate2e): Q(a) = -log2(erfc(a / sqrt 2)) on a = |x| in [0, 6] as a
polynomial without constant term, weighted by the GELU error it causes; prints the float32 max-abs error of gelu per degree."""
import numpy as np
from scipy.special import erfc, erf
from numpy.polynomial import chebyshev as C, polynomial as P
XMAX=6.0
def target(a):  # Q(a) = -log2(erfc(a/sqrt2)), a=|x|
    return -np.log2(erfc(a/np.sqrt(2.0)))
a=np.linspace(0,XMAX,20001)
q=target(a)
# weight: gelu error = 0.5*a*e*ln2*dQ (x<0 and x>0 same magnitude)
w=0.5*np.maximum(a,1e-3)*erfc(a/np.sqrt(2))*np.log(2)
best={}
for deg in range(3,9):   # the library uses degree 6 (IRB_GELU_DEG in csrc/gdfn_math.cuh); 4 and 5 are selectable, measured no faster
    # iteratively reweighted LS to approach minimax of weighted error
    ww=w.copy()
    for it in range(200):
        V=np.vander(a,deg+1,increasing=True)[:,1:]   # no constant term: Q(0)=0
        coef,*_=np.linalg.lstsq(V*ww[:,None],q*ww,rcond=None)
        err=(V@coef-q)*w
        ww=ww*(1+4*np.abs(err)/np.abs(err).max())
        ww/=ww.max()/w.max()
    coef=np.concatenate([[0.0],coef])
    # float32 evaluation
    x=np.linspace(-8,8,400001).astype(np.float32)
    ax=np.minimum(np.abs(x),np.float32(XMAX))
    c32=coef.astype(np.float32)
    Q=np.zeros_like(ax)
    for k in range(deg,0,-1):
        Q=(Q+c32[k])*ax if k==deg else (Q+c32[k])*ax
    # Horner: Q = ax*(c1 + ax*(c2+...))
    Q=np.zeros_like(ax)
    for k in range(deg,0,-1):
        Q=np.float32(Q*ax+c32[k]) if k<deg else np.full_like(ax,c32[k])
    Q=(Q*ax).astype(np.float32)
    e=np.exp2(-Q.astype(np.float64)).astype(np.float32)
    hx=(np.float32(0.5)*x)
    r=(hx*e).astype(np.float32)
    out=np.where(x>0,(x-r).astype(np.float32),r)
    ref=0.5*x.astype(np.float64)*(1+erf(x.astype(np.float64)/np.sqrt(2)))
    print(deg, "max abs gelu err", np.abs(out-ref).max(), "at", x[np.abs(out-ref).argmax()])
    best[deg]=coef
np.save("/tmp/gelu_coef.npy", best[6])
np.set_printoptions(precision=12); print(repr(best[6]))
