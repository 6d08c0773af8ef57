import sys, time; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import numpy as np
from oracle import gp_oracle as o
from scipy.interpolate import griddata
# ---- G2 Eikonal
np.random.seed(20)
Xd,Xb=o.notebook_sample_points(400,84)
p=o.Eikonal(eps=1e-2); p.set_points(Xd,Xb,np.ones(400),np.zeros(84))
p.Gram_matrix('Gaussian',0.2,1e-6,'adaptive'); print('ratio',p.ratio)
p.Gram_Cholesky('lu'); t=time.time(); p.GN_method(10,1.0,'zero'); print(time.time()-t)
for h in p.loss_hist: print(repr(h))
XX,YY,truth=o.solve_Eikonal(100,1e-2)
Xt=np.stack([XX.ravel(),YY.ravel()],1)
p.extend_sol(Xt); e=np.abs(p.extended_sol.reshape(100,100)-truth)
print('L2',np.linalg.norm(e)/100,'max',e.max(), 'gold 0.02506445909677251 0.06384742102752328')
# ---- G3 Darcy
np.random.seed(10)
ut=o.FD_Darcy_flow_2d(100)
xx=np.linspace(0,1,102); XX,YY=np.meshgrid(xx,xx); XXv=XX.flatten(); YYv=YY.flatten()
Xd,Xb=o.notebook_sample_points(400,100)
init=np.random.normal(0,1.0,2400)
data_u=griddata((XXv,YYv),ut.reshape(-1),(Xd[:40,0],Xd[:40,1]),method='linear')
d=o.Darcy_flow2d(); d.set_points(Xd,Xb,40,np.ones(400),np.zeros(100))
d.get_observation(data_u,1e-3)
d.Gram_matrix('Gaussian',0.2,1e-5,'adaptive'); print(d.ratio_u,d.ratio_a)
d.Gram_Cholesky('lu'); t=time.time(); d.GN_method(8,1,init); print(time.time()-t)
for h in d.loss_hist: print(repr(h))
