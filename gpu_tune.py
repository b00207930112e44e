import sys, os, time, json; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, bench
from qgdsolver_b200 import api
api.init(0)
n=int(sys.argv[1]) if len(sys.argv)>1 else 256
c=bench.build_case(n)
dm=api.Mesh(c.mesh)
for v in (sys.argv[2].split(',') if len(sys.argv)>2 else ['0','1','2','3','4','5','6']):
    os.environ['QGD_FACE_VARIANT']=v
    s=c.make_solver(api, dm)
    s.step(3); api.synchronize()
    s.profile(True); api.timer_begin(); s.step(10); ms=api.timer_end(); kt=s.kernel_times(); s.profile(False)
    print('variant',v,'ms/step %.3f'%(ms/10),'MCUPS %.0f'%(c.mesh.n_cells/(ms/10*1e-3)/1e6),{k:round(x/10,3) for k,x in kt.items() if k!='steps'}, flush=True)
    s.close()
