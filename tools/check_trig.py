"""tools/check_trig.py — derivation and CPU check of the f32 sin/cos in hpt_b200/csrc/ops.cuh (sincos_f32_fast).

Fits the minimax polynomials (Lawson-weighted least squares on Chebyshev nodes, relative error), rounds the
coefficients and the three-term split of pi/2 to f32, emulates the device code path with numpy (fmaf = exact
product in f64, one rounding) and prints the max ulp error against f64 sin/cos per input range."""
import numpy as np, mpmath as mp
f32=np.float32; f64=np.float64
def fma(a,b,c): return (np.asarray(a,f32).astype(f64)*np.asarray(b,f32).astype(f64)+np.asarray(c,f32).astype(f64)).astype(f32)
mp.mp.prec=200
PI2=mp.pi/2
c1=f32(float(PI2)); c2=f32(float(PI2-mp.mpf(float(c1)))); c3=f32(float(PI2-mp.mpf(float(c1))-mp.mpf(float(c2))))
print("c", repr(c1),repr(c2),repr(c3))
# fit on r in [-pi/4-eps, pi/4+eps]
R=np.pi/4*1.0005
k=np.arange(4000); r=R*np.cos((k+0.5)*np.pi/4000); r=r[r>1e-6]
s=r*r
def lawson(A,y,w0,it=60):
    w=np.ones_like(y)
    for _ in range(it):
        c,*_=np.linalg.lstsq(A*(w*w0)[:,None],y*w*w0,rcond=None)
        e=np.abs((A@c-y)*w0); w=w*(0.5+e/e.max()); w/=w.mean()
    return c,e.max()
# sin: (sin r - r)/(r s) = S0+S1 s+S2 s^2 ; relative error weight: r*s/sin r
for ns in (3,4):
    A=np.stack([s**i for i in range(ns)],1); y=(np.sin(r)-r)/(r*s); w0=r*s/np.sin(r)
    cs,e=lawson(A,y,w0); print("sin",ns,[repr(f32(x)) for x in cs],"relerr %.3g ulp~%.3f"%(e,e/2**-24))
for nc in (3,4):
    A=np.stack([s**i for i in range(nc)],1); y=(np.cos(r)-1)/s; w0=s/np.cos(r)
    cc,e=lawson(A,y,w0); print("cos",nc,[repr(f32(x)) for x in cc],"relerr %.3g ulp~%.3f"%(e,e/2**-24))

S=[f32(-0.16666655),f32(0.008332158),f32(-0.00019514957)]
C=[f32(-0.5),f32(0.04166662),f32(-0.0013886677),f32(2.4383144e-05)]
def sincos_emul(x, want_cos=False):
    x=np.asarray(x,f32)
    j=fma(x,f32(0.63661977),f32(12582912.0))
    q=j.view(np.int32).copy()
    j=(j-f32(12582912.0)).astype(f32)
    r=fma(j,-c1,x); r=fma(j,-c2,r); r=fma(j,-c3,r)
    if want_cos: q=q+1
    s=(r*r).astype(f32)
    # sin poly
    p=fma(S[2],s,S[1]); p=fma(p,s,S[0]); t=(r*s).astype(f32); ps=fma(p,t,r)
    pc=fma(C[3],s,C[2]); pc=fma(pc,s,C[1]); pc=fma(pc,s,C[0]); pc=fma(pc,s,f32(1.0))
    res=np.where(q&1,pc,ps)
    res=np.where(q&2,-res,res)
    return res.astype(f32)
def ulp_err(got,x,fn):
    ref=fn(x.astype(f64)); 
    ref32=ref.astype(f32)
    u=np.abs(np.spacing(np.abs(ref32))).astype(f64)
    # handle binade boundary: use ulp of ref magnitude
    return np.abs(got.astype(f64)-ref)/u
rng=np.random.default_rng(0)
for name,x in [("randn",rng.standard_normal(4_000_000).astype(f32)),
               ("u[-10,10]",rng.uniform(-10,10,4_000_000).astype(f32)),
               ("u[-1000,1000]",rng.uniform(-1000,1000,4_000_000).astype(f32)),
               ("u[-105615,105615]",rng.uniform(-105615,105615,8_000_000).astype(f32)),
               ("near k*pi/2",(np.arange(1,67000)[:,None]*np.pi/2+np.linspace(-3e-3,3e-3,61)[None,:]).astype(f32).ravel()),
               ("tiny",rng.uniform(-1e-3,1e-3,1_000_000).astype(f32))]:
    for wc,fn in ((False,np.sin),(True,np.cos)):
        e=ulp_err(sincos_emul(x,wc),x,fn)
        print(name,"cos" if wc else "sin","max ulp %.3f"%e.max(), "at", x[e.argmax()])
