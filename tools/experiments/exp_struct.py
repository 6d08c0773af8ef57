import sys, time; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import numpy as np, scipy.linalg as sla
from oracle import gp_oracle as o
def run(nugget, N=900, Nb=124, steps=4, seed=0):
    np.random.seed(seed)
    Xd,Xb=o.sampled_pts_rdm(N,Nb,np.array([[0,1],[0,1.]]))
    init=np.random.normal(0,1,N)
    p=o.Nonlinear_elliptic2d(alpha=1.0,m=3)
    p.set_points(Xd,Xb,o.elliptic_f(Xd[:,0],Xd[:,1]),o.elliptic_u(Xb[:,0],Xb[:,1]))
    p.Gram_matrix('Gaussian',0.2,nugget,'adaptive')
    truth=o.elliptic_u(Xd[:,0],Xd[:,1])
    res={}
    p.Gram_Cholesky('lu'); p.GN_method(steps,1,init); 
    err=np.abs(truth-p.sol_sampled_pts); res['lu']=(np.sqrt(np.mean(err**2)),err.max(),p.loss_hist[-1])
    L=p.L; M=L.shape[0]
    def loss(z): s=sla.solve_triangular(L,p.F(z),lower=True); return s@s
    def gn(variant, hsolve='chol'):
        z=init.copy()
        if variant in ('A','W'):
            W=sla.solve_triangular(L,np.eye(M),lower=True)
            if variant=='A':
                A=W.T@W
                A11=A[:N,:N];A12=A[:N,N:2*N];A22=A[N:2*N,N:2*N]
        for it in range(steps):
            F=p.F(z); D=3*z*z
            s=sla.solve_triangular(L,F,lower=True)
            t=sla.solve_triangular(L,s,lower=True,trans='T')
            g=2*(D*t[:N]+t[N:2*N])
            if variant=='dense':
                J=p.jacobian(z); S=sla.solve_triangular(L,J,lower=True); H=2*S.T@S
            elif variant=='W':
                S=W[:,:N]*D[None,:]+W[:,N:2*N]; H=2*S.T@S
            else:
                H=2*(D[:,None]*A11*D[None,:]+D[:,None]*A12+A12.T*D[None,:]+A22)
            if hsolve=='chol':
                try:
                    c=sla.cho_factor(H,lower=True); d=sla.cho_solve(c,g)
                except Exception as e:
                    d=np.linalg.solve(H,g); print('   chol(H) failed, LU used')
            else: d=np.linalg.solve(H,g)
            z=z-d
        err=np.abs(truth-z); return (np.sqrt(np.mean(err**2)),err.max(),loss(z))
    for v in ['dense','W','A']:
        res[v]=gn(v)
    res['dense_luH']=gn('dense','lu')
    print(f'nugget {nugget:g}')
    for k,v in res.items(): print(f'  {k:10s} L2 {v[0]:.6e} max {v[1]:.6e} loss {v[2]:.12e}  relL2 vs lu {abs(v[0]-res["lu"][0])/res["lu"][0]:.2e}')
for ng in [1e-13,1e-10,1e-8,1e-5]:
    run(ng)
