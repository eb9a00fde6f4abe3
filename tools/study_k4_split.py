"""CPU study behind K4 on the tensor cores: the surrogate energy / score of tests/golden/scat_energy.npz with the hidden
GEMMs emulated as plain bf16 (x1), bf16x3 (hi hi + hi lo + lo hi) and bf16x6 split products, in the units of the test
tolerances (f: 1e-5, E: 2e-4 rel + 1e-2, grad: 2e-4 of its scale).  python tools/study_k4_split.py"""
import sys, numpy as np, torch
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from util import load_golden, surrogate_params
fx = load_golden("scat_energy"); sp = surrogate_params()
x = fx["x"].double().numpy(); y = fx["y"].double().numpy()
Ws=[(W.double().numpy(), b.double().numpy()) for W,b in sp]
def bf(v):  # round to bf16 (RNE) keep as float64
    t = torch.from_numpy(np.asarray(v,dtype=np.float32)).to(torch.bfloat16).to(torch.float32).double().numpy(); return t
def split(v, parts):
    out=[]; r=np.asarray(v,dtype=np.float32).astype(np.float64)
    for _ in range(parts):
        h=bf(r); out.append(h); r=r-h
    return out
def mm(A,B,mode):
    # A [n,k], B [k,m]
    if mode=="f64": return A@B
    if mode=="f32": return (A.astype(np.float32)@B.astype(np.float32)).astype(np.float64)
    if mode=="x1":
        return (bf(A)@bf(B)).astype(np.float32).astype(np.float64)
    if mode=="x3":
        a=split(A,2); b=split(B,2)
        return (a[0]@b[0]+a[0]@b[1]+a[1]@b[0]).astype(np.float32).astype(np.float64)
    if mode=="x6":
        a=split(A,3); b=split(B,3)
        return (a[0]@b[0]+a[0]@b[1]+a[1]@b[0]+a[0]@b[2]+a[2]@b[0]+a[1]@b[1]).astype(np.float32).astype(np.float64)
def run(mode, l0="f64" ):
    f32=lambda v: v if mode=="f64" else v.astype(np.float32).astype(np.float64)
    z=[None]*4; h=x
    hs=[x]
    for l,(W,b) in enumerate(Ws):
        m = mode if l in (1,2,3) else ("f64" if mode=="f64" else "f32")
        zz = f32(mm(h,W.T,m)+b)
        z[l]=zz
        h = np.maximum(zz,0) if l<3 else zz
        hs.append(h)
    f=h; a,b_,lam=0.2,0.01,1000.0
    pre=(a*f)**2+b_**2
    E=0.5*np.log(pre).sum(1)+0.5*((y-f)**2/pre).sum(1)+lam*(np.maximum(x-1,0)+np.maximum(-1-x,0)).sum(1)
    dEdf = a*a*f/pre - (y-f)/pre - (y-f)**2*a*a*f/pre**2
    g=f32(dEdf)
    for l in (3,2,1,0):
        W=Ws[l][0]
        m = mode if l in (1,2,3) else ("f64" if mode=="f64" else "f32")
        g=f32(mm(g,W,m))
        if l>0: g=g*(z[l-1]>0)
    g=g+lam*((x>1).astype(float)-(x<-1).astype(float))
    return f,E,g
f0,E0,g0=run("f64")
print("fixture vs f64: f",np.abs(f0-fx["fx"].numpy()).max(), "E rel",(np.abs(E0-fx["E"].numpy())/(2e-4*np.abs(E0)+1e-2)).max(), "g", np.abs(g0-fx["grad"].numpy()).max()/(2e-4*np.abs(g0).max()))
for mode in ("f32","x6","x3","x1"):
    f,E,g=run(mode)
    print(mode,"e_f(1e-5 units)",np.abs(f-f0).max()/1e-5,"e_E",(np.abs(E-E0)/(2e-4*np.abs(E0)+1e-2)).max(),"e_g",np.abs(g-g0).max()/(2e-4*np.abs(g0).max()), "E max", np.abs(E0).max(), "gmax", np.abs(g0).max())
print("---- mixed: x3 on layers 1,2 (+backward), last layer modes")
def run2(mid, last):
    f32=lambda v: v.astype(np.float32).astype(np.float64)
    z=[None]*4; h=x
    for l,(W,b) in enumerate(Ws):
        m = "f32" if l==0 else (last if l==3 else mid)
        zz=f32(mm(h,W.T,m)+b); z[l]=zz; h=np.maximum(zz,0) if l<3 else zz
    f=h; a,b_,lam=0.2,0.01,1000.0
    pre=(a*f)**2+b_**2
    E=0.5*np.log(pre).sum(1)+0.5*((y-f)**2/pre).sum(1)+lam*(np.maximum(x-1,0)+np.maximum(-1-x,0)).sum(1)
    dEdf = a*a*f/pre - (y-f)/pre - (y-f)**2*a*a*f/pre**2
    g=f32(dEdf)
    for l in (3,2,1,0):
        W=Ws[l][0]; m = "f32" if l==0 else (last if l==3 else mid)
        g=f32(mm(g,W,m))
        if l>0: g=g*(z[l-1]>0)
    g=g+lam*((x>1).astype(float)-(x<-1).astype(float))
    return f,E,g
for mid,last in (("x3","x3"),("x3","x6"),("x3","f32"),("x6","x6")):
    f,E,g=run2(mid,last)
    print(mid,last,"e_f",np.abs(f-f0).max()/1e-5,"e_E",(np.abs(E-E0)/(2e-4*np.abs(E0)+1e-2)).max(),"e_g",np.abs(g-g0).max()/(2e-4*np.abs(g0).max()))
print("f scale", np.abs(f0).max(), "x scale", np.abs(x).max())
