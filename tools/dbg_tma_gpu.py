import sys, os, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import oracle, darsia_b200 as da
from golden.generate_golden import phantom
def run(shape, dtype, tile, nit=2):
    img=phantom(shape,13,dtype)
    kw=dict(mu=0.1,omega=1.3,ell=0.9,dim=3,max_num_iter=nit,isotropic=False)
    out=da.split_bregman_tvd(img,solver=da.Jacobi(maxiter=1),tile=tile,**kw)
    ref=oracle.split_bregman_tvd(img,solver=oracle.Jacobi(maxiter=1),**kw)
    err=np.abs(out.astype(np.float64)-ref.astype(np.float64))
    bad=np.argwhere(err>1e-4)
    print(shape,dtype.__name__,tile,nit,'max',err.max(),'nbad',len(bad))
    if len(bad):
        for ax,n in zip(range(3),'zyx'):
            vals=sorted(set(bad[:,ax].tolist())); print('  ',n,len(vals),vals[:24],'...',vals[-6:])
for nit in (1,2):
    run((130,61,257),np.float32,(0,0,0),nit)
run((130,61,257),np.float32,(16,8,0))
run((130,61,257),np.float32,(8,16,0))
run((130,61,257),np.float64,(0,0,0))
run((130,61,257),np.float64,(32,8,0))
run((130,61,257),np.float64,(16,8,0))
run((64,61,257),np.float32,(0,0,0))
run((130,61,120),np.float32,(0,0,0))
run((130,30,257),np.float32,(0,0,0))
