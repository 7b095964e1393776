import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from helpers import golden_region
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import single_region_case, synthetic_coeffs
from oracle import pyoracle
g=np.load('/root/repo/tests/golden/polymesh_addr.npz')
ctx=ldu.Context(0)
case=single_region_case(synthetic_coeffs(40, np.empty(0, np.int32), np.empty(0, np.int32), symmetric=False))
O=pyoracle.OracleSystem(case); S=ldu.LduSystem(ctx,case.ranks[0])
x0,b=case.concat('psi'),case.concat('source')
try:
    xo,io=O.solve(x0,b,'BiCGStab','DILU',tolerance=1e-12,maxIter=10)
    xg,ig=S.solve(x0,b,ldu.SOLVER_BICGSTAB,ldu.PRECOND_DILU,tolerance=1e-12,maxIter=10)
    print('no_faces oracle',io['nIterations'],io['history'],'gpu',ig['nIterations'],ig['history'], np.abs(xg-xo).max())
except Exception as e: print('no_faces ERR',e)
S.close()
case=single_region_case(golden_region(g,'duineveld1',False))
O=pyoracle.OracleSystem(case); S=ldu.LduSystem(ctx,case.ranks[0])
x0,b=case.concat('psi'),case.concat('source')
xo,io=O.solve(x0,b,'BiCGStab','none',tolerance=1e-10,maxIter=2000)
xg,ig=S.solve(x0,b,ldu.SOLVER_BICGSTAB,ldu.PRECOND_NONE,tolerance=1e-10,maxIter=2000)
O.set_reduction_mode(1); xa,ia=O.solve(x0,b,'BiCGStab','none',tolerance=1e-10,maxIter=2000)
ho,hg,ha=io['history'],ig['history'],ia['history']
k=21
print('none: its',io['nIterations'],ig['nIterations'],ia['nIterations'])
print('rel gpu',np.abs(hg[:k]-ho[:k])/ho[:k]); print('rel alt',np.abs(ha[:k]-ho[:k])/ho[:k])
