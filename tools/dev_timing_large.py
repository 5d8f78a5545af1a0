import ctypes as C, importlib.util, os, sys
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg=os.path.join(ROOT,"pde-based-heston-solver-gpu-accelerated_b200")
spec=importlib.util.spec_from_file_location("hadi",os.path.join(pkg,"hadi.py")); hadi=importlib.util.module_from_spec(spec); spec.loader.exec_module(hadi)
hadi.LIB_PATH=os.path.join(pkg,os.environ.get("HADI_LIB","libhadi_timing.so"))
os.environ["HADI_FORCE_VARIANT"]="5"
L=hadi.lib(); L.hadi_batch_phase_cycles.argtypes=[C.c_void_p,C.POINTER(C.c_longlong)]
ctx=hadi.Context(0)
mdl=hadi.make_model(S0=100.0,V0=0.04,r_d=0.025,r_f=0.0,rho=-0.9,sigma=0.3,kappa=1.5,eta=0.04)
names=["setup+div","a1fwd","explicit","a1","a2","project","ringwait","rhs2"]
for (n,N,m1,m2) in [(148,20,400,200),(1,20,400,200)]:
    num=hadi.make_numerics(m1,m2,0.8,0,0,0,None)
    pts,n=hadi.make_points([100.0+k for k in range(n)],1.0,N)
    bt=ctx.batch(mdl,num,pts,n)
    for r in range(2): bt.launch(); bt.fetch()
    ms=bt.elapsed_ms(); cyc=(C.c_longlong*8)(); L.hadi_batch_phase_cycles(bt._h,cyc)
    print(m1,m2,N,"ms",ms,{names[k]:round(cyc[k]/(n*N)) for k in range(8)})
