import sys, time; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import numpy as np, scipy.linalg as sla
from oracle import gp_oracle as o
def kroute_elliptic(p, init, steps):
    N, Nb = p.N_domain, p.N_boundary; Th = p.Theta; L = p.L
    T11=Th[:N,:N]; T12=Th[:N,N:2*N]; T13=Th[:N,2*N:]; T22=Th[N:2*N,N:2*N]; T23=Th[N:2*N,2*N:]; T33=Th[2*N:,2*N:]
    z = init.copy(); hist=[]
    def loss(z): s=sla.solve_triangular(L,p.F(z),lower=True); return s@s
    hist.append(loss(z))
    for it in range(steps):
        D = p.alpha*p.m*o.int_pow(z,p.m-1)
        c = (p.alpha*o.int_pow(z,p.m)-p.rhs_f) - D*z
        K11 = T11 - T12*D[None,:] - D[:,None]*T12.T + D[:,None]*T22*D[None,:]
        K1b = T13 - D[:,None]*T23
        K = np.block([[K11,K1b],[K1b.T,T33]])
        y = np.concatenate([c,p.bdy_g])
        cf = sla.cho_factor(K, lower=True); lam = sla.cho_solve(cf, y)
        l1, lb = lam[:N], lam[N:]
        # z+ = Theta[delta_int rows,:] @ C^T lam, C^T lam = [l1; -D l1; lb]
        z = Th[N:2*N,:N]@l1 - T22@(D*l1) + T23@lb
        hist.append(loss(z))
    return z, hist
for nug in [1e-4,1e-8,1e-10,1e-13]:
    np.random.seed(0)
    N,Nb=900,124
    Xd,Xb=o.sampled_pts_rdm(N,Nb,np.array([[0,1],[0,1.]])); init=np.random.normal(0,1,N)
    p=o.Nonlinear_elliptic2d(alpha=1.0,m=3); p.set_points(Xd,Xb,o.elliptic_f(Xd[:,0],Xd[:,1]),o.elliptic_u(Xb[:,0],Xb[:,1]))
    p.Gram_matrix('Gaussian',0.2,nug,'adaptive'); p.Gram_Cholesky('lu'); p.GN_method(4,1,init)
    truth=o.elliptic_u(Xd[:,0],Xd[:,1]); e_ref=np.sqrt(np.mean((truth-p.sol_sampled_pts)**2))
    try:
        z,hist=kroute_elliptic(p,init,4)
        e_k=np.sqrt(np.mean((truth-z)**2))
        print(f'nugget {nug:g}: ref L2 {e_ref:.6e}  K-route L2 {e_k:.6e} rel diff {abs(e_k-e_ref)/e_ref:.2e}; loss hist rel diff', np.max(np.abs(np.array(hist)-np.array(p.loss_hist))/np.abs(p.loss_hist)))
    except Exception as ex:
        print(f'nugget {nug:g}: K-route failed: {ex}')
