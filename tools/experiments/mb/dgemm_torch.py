import torch, time
torch.backends.cuda.matmul.allow_tf32=False
n=8192
a=torch.randn(n,n,dtype=torch.float64,device='cuda'); b=torch.randn(n,n,dtype=torch.float64,device='cuda')
for _ in range(2): c=a@b.T
torch.cuda.synchronize()
best=1e9
for r in range(5):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); c=a@b.T; e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
print(f"cuBLAS DGEMM NT {n}^3 burst: {best:.2f} ms {2*n**3/best/1e9:.2f} TFLOP/s")
t0=time.time(); k=0
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
while time.time()-t0<4: 
    for _ in range(5): c=a@b.T
    torch.cuda.synchronize(); k+=5
e1.record(); torch.cuda.synchronize()
print(f"cuBLAS DGEMM sustained: {2*n**3*k/e0.elapsed_time(e1)/1e9:.2f} TFLOP/s over {k} iters")
# potrf comparator
for m in (8192, 20480):
    x=torch.randn(m,m,dtype=torch.float64,device='cuda'); s=x@x.T+m*torch.eye(m,dtype=torch.float64,device='cuda'); del x
    torch.linalg.cholesky(s); torch.cuda.synchronize()
    e0.record(); L=torch.linalg.cholesky(s); e1.record(); torch.cuda.synchronize()
    print(f"cusolver potrf {m}: {e0.elapsed_time(e1):.2f} ms {m**3/3/e0.elapsed_time(e1)/1e9:.2f} TFLOP/s")
    del s,L
