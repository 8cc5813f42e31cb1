import sys; sys.path.insert(0,'.')
import numpy as np, finetools_jl_b200 as fe
n=int(sys.argv[1]) if len(sys.argv)>1 else 96
fens, fes = fe.H8block(1.0,1.0,1.0,n,n,n)
u = fe.NodalField(np.zeros((fens.count(),3))); fe.numberdofs(u)
lam, mu = 0.3/(1.3*0.4), 1/2.6
C=np.zeros((6,6)); C[:3,:3]=lam; C[np.arange(3),np.arange(3)]+=2*mu; C[3:,3:]=mu*np.eye(3)
a = fe.SysmatAssemblerSparseGPU(0.0)
femm = fe.FEMMBase(fe.IntegDomain(fes, fe.GaussRule(3,2)))
geom = fe.NodalField(fens.xyz)
a.setnomatrixresult(True)  # keep the CSC on the device: kernels only
for i in range(2):
    a.invalidate_patterns()
    fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, fe.DataCache(C), raw=True)
print(a.timings())
