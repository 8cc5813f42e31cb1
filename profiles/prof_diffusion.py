import sys; sys.path.insert(0,'.')
import numpy as np, finetools_jl_b200 as fe
n=int(sys.argv[1]) if len(sys.argv)>1 else 256
fens, fes = fe.H8block(1.0,1.0,1.0,n,n,n)
u = fe.NodalField(np.zeros((fens.count(),1))); fe.numberdofs(u)
kappa = np.array([[1.5, 0.2, 0.1], [0.2, 2.5, 0.3], [0.1, 0.3, 3.5]])
a = fe.SysmatAssemblerSparseGPU(0.0)
femm = fe.FEMMBase(fe.IntegDomain(fes, fe.GaussRule(3,2)))
geom = fe.NodalField(fens.xyz)
a.setnomatrixresult(True)  # keep the CSC on the device: kernels only
for i in range(2):
    a.invalidate_patterns()
    fe.bilform_diffusion(femm, a, geom, u, fe.DataCache(kappa), raw=True)
print(a.timings())
