import sys, json, numpy as np
sys.path.insert(0,'.')
from glomap_b200 import estimators as E, synthetic as S
wl = sys.argv[1] if len(sys.argv)>1 else "config2"
C,P,ch = (1000,200000,25000) if wl=="config2" else (10000,2000000,50000)
sc = S.make_scene(C,P,10.0,seed=1,pixel_sigma=0.5,chunk=ch); init=S.perturb_scene(sc,chunk=ch); mask=E.first_frame_mask(C)
ctx=E.default_context(); prob=E.BAProblem(ctx,sc,3,mask); prob.set_state(init.intr_params,init.quat,init.trans,init.points); prob.save_state()
for tol in (0.3,0.1,0.05,0.03,0.01,0.003,1e-3,1e-6):
    o=E.BundleAdjusterOptions(optimize_intrinsics=False); o.solver_options.pcg_rel_tolerance=tol; o.solver_options.pcg_max_iterations=2000
    for _ in range(2):
        prob.restore_state(); st=prob.solve(o)
    print(json.dumps(dict(workload=wl,tol=tol,lm=st.iterations,pcg=st.pcg_iterations,ms=round(st.ms_total,2),cost=st.final_cost,term=st.termination)))
# the reference default: optimize_intrinsics = 1 (extended path), shared intrinsics
for tol in (0.1,0.01):
    o=E.BundleAdjusterOptions(optimize_intrinsics=True); o.solver_options.pcg_rel_tolerance=tol; o.solver_options.pcg_max_iterations=2000
    for _ in range(2):
        prob.restore_state(); st=prob.solve(o)
    print(json.dumps(dict(workload=wl,intrinsics=True,tol=tol,lm=st.iterations,pcg=st.pcg_iterations,ms=round(st.ms_total,2),cost=st.final_cost)))
