import sys, time; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import numpy as np
from oracle import gp_oracle as o
np.random.seed(10)
Xd,Xb=o.notebook_sample_points(900,124)
init=np.random.normal(0,1,900)
p=o.Nonlinear_elliptic2d(alpha=1,m=3)
p.set_points(Xd,Xb,o.elliptic_f(Xd[:,0],Xd[:,1]),o.elliptic_u(Xb[:,0],Xb[:,1]))
for mode in ['lu','tri']:
    t=time.time()
    p.Gram_matrix('Gaussian',0.2,1e-4,'adaptive'); p.Gram_Cholesky(mode)
    print('ratio',repr(p.ratio))
    p.GN_method(5,1,init)
    print(mode,[repr(h) for h in p.loss_hist], time.time()-t)
    err=np.abs(o.elliptic_u(Xd[:,0],Xd[:,1])-p.sol_sampled_pts)
    print('L2',repr(np.sqrt(np.sum(err**2)/900)),'max',repr(err.max()))
    xx=np.linspace(0,1,100); XX,YY=np.meshgrid(xx,xx); Xt=np.stack([XX.ravel(),YY.ravel()],1)
    p.extend_sol(Xt); e=np.abs(p.extended_sol-o.elliptic_u(Xt[:,0],Xt[:,1]))
    print('test L2',repr(np.linalg.norm(e)/100),'max',repr(e.max()))
