import sys, os
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/cu-bens_b200/python')
import numpy as np
import cubens_b200 as cb
from cubens_b200 import meshgen, model as M
from oracle import refbind as ref
from util import skyline_to_dense, csc_to_dense
p = meshgen.plate_model(4, 3, z_bump=0.02, pinned=False)
X = p.x.reshape(-1, 3); shells = p.minc.reshape(-1, 3)
trusses = np.array([[1, 7], [6, 12], [11, 17]])
fixed = [(1, d) for d in range(1, 8)] + [(20, 1), (20, 2), (20, 3), (16, 3)]
m = M.build_model(p.x, trusses=trusses, shells=shells, fixed=fixed, truss_props=(2.1e11, 1e-4, 8050.0, 3.45e8), shell_props=meshgen.SHELL_5C, ANAFLAG=2)
s = ref.RefState(m); s.begin_increment()
a = ref.stiff(m, s, SLVFLAG=0); Ka = skyline_to_dense(m.NEQ, m.maxa, a)
Kd = ref.stiff(m, s, SLVFLAG=2).reshape(m.NEQ,m.NEQ).T
jc = m.jcode.reshape(-1,7)
def who(eq):
    jj,dd = np.argwhere(jc==eq+1)[0]; return (int(jj)+1,int(dd))
for lay,name in ((cb.CB_MAT_SKYLINE,'sky'),(cb.CB_MAT_CSC,'csc')):
    asm = cb.Assembler(m, layout=lay); asm.begin_increment(); asm.stiff()
    if name=='sky':
        Kb = skyline_to_dense(m.NEQ, m.maxa, asm.skyline()); R=Ka
    else:
        Kb = csc_to_dense(m.NEQ, *asm.csc()); R=Kd
    D = np.abs(R-Kb); bad = np.argwhere(D > 1e-9*np.abs(R).max())
    print(name, 'nbad', len(bad))
    seen=set()
    for r,c in bad:
        key=(who(r)[0],who(c)[0])
        if key not in seen: seen.add(key); print('   joints', key, 'ref', R[r,c], 'dev', Kb[r,c])
    asm.close()
