#!/usr/bin/env python
"""Numerical check of the solid-harmonic identities oracle/fmm_oracle.c is built on (normalisation of Dehnen 2014):
expansion of 1/|x - y|, addition theorems of the regular / irregular harmonics, derivative ladder, three-term recurrences.
Each block prints the exact value next to the series / recurrence value.  Needs scipy (lpmv); test infrastructure only."""
import numpy as np
from scipy.special import lpmv, factorial
rng=np.random.default_rng(0)
def sph(x):
    r=np.linalg.norm(x); ct=x[2]/r; ph=np.arctan2(x[1],x[0]); return r,ct,ph
def Preg(n,m,x):
    # Dehnen-like regular harmonic: r^n P_n^m(cos t) e^{i m phi}/(n+m)!   (m may be negative)
    r,ct,ph=sph(x)
    if abs(m)>n: return 0j
    am=abs(m)
    v=r**n*lpmv(am,n,ct)*np.exp(1j*am*ph)/factorial(n+am)
    if m<0: v=(-1)**am*np.conj(v)
    return v
def Pirr(n,m,x):
    r,ct,ph=sph(x)
    if abs(m)>n: return 0j
    am=abs(m)
    v=factorial(n-am)*lpmv(am,n,ct)*np.exp(1j*am*ph)/r**(n+1)
    if m<0: v=(-1)**am*np.conj(v)
    return v
x=rng.standard_normal(3)*3; y=rng.standard_normal(3)*0.3
exact=1/np.linalg.norm(x-y)
N=25
for desc,f in [("conj reg", lambda n,m: np.conj(Preg(n,m,y))*Pirr(n,m,x)),
               ("reg(-m)", lambda n,m: Preg(n,-m,y)*Pirr(n,m,x)),
               ("(-1)^m reg(-m)", lambda n,m: (-1)**m*Preg(n,-m,y)*Pirr(n,m,x)),]:
    s=sum(f(n,m) for n in range(N) for m in range(-n,n+1))
    print(desc, s, exact)
print("--- regular addition")
a=rng.standard_normal(3); b=rng.standard_normal(3)
for (n,m) in [(3,1),(4,-2),(2,2),(5,0)]:
    ex=Preg(n,m,a+b)
    s=sum(Preg(k,l,a)*Preg(n-k,m-l,b) for k in range(n+1) for l in range(-k,k+1))
    print(n,m,ex,s)
print("--- irregular addition |a|<|b|")
a=rng.standard_normal(3)*0.3; b=rng.standard_normal(3)*3
K=22
for (n,m) in [(0,0),(2,1),(3,-2)]:
    ex=Pirr(n,m,a+b)
    for desc,f in [("(-1)^k conj(R_k^l(a)) I_{n+k}^{m+l}", lambda k,l: (-1)**k*np.conj(Preg(k,l,a))*Pirr(n+k,m+l,b)),
                   ("(-1)^k R_k^{-l}(a)... (-1)^l", lambda k,l: (-1)**(k+l)*Preg(k,-l,a)*Pirr(n+k,m+l,b)),
                   ("conj(R_k^l(-a)) I", lambda k,l: np.conj(Preg(k,l,-a))*Pirr(n+k,m+l,b))]:
        s=sum(f(k,l) for k in range(K) for l in range(-k,k+1))
        print(n,m,desc,ex,s)
print("--- derivatives of regular harmonics")
x=rng.standard_normal(3); h=1e-6
def grad(f,x):
    g=[]
    for c in range(3):
        e=np.zeros(3); e[c]=h
        g.append((f(x+e)-f(x-e))/(2*h))
    return g
for (n,m) in [(3,1),(4,-2),(3,3),(2,0),(3,-3)]:
    gx,gy,gz=grad(lambda p: Preg(n,m,p), x)
    print(n,m,"dz",gz,Preg(n-1,m,x)," d+",gx+1j*gy,Preg(n-1,m+1,x)," d-",gx-1j*gy,Preg(n-1,m-1,x))
print("--- recurrences")
def tables(x,N):
    X,Y,Z=x; r2=X*X+Y*Y+Z*Z
    R={}; I={}
    R[(0,0)]=1+0j; I[(0,0)]=1/np.sqrt(r2)+0j
    for m in range(0,N+1):
        if m>0:
            R[(m,m)]=-(X+1j*Y)/(2*m)*R[(m-1,m-1)]
            I[(m,m)]=-(2*m-1)*(X+1j*Y)/r2*I[(m-1,m-1)]
        for n in range(m+1,N+1):
            Rm2=R.get((n-2,m),0j); Im2=I.get((n-2,m),0j)
            R[(n,m)]=((2*n-1)*Z*R[(n-1,m)]-r2*Rm2)/((n+m)*(n-m))
            I[(n,m)]=((2*n-1)*Z*I[(n-1,m)]-((n-1)**2-m*m)*Im2)/r2
    return R,I
x=rng.standard_normal(3)
R,I=tables(x,6)
err=max(abs(R[k]-Preg(k[0],k[1],x)) for k in R); err2=max(abs(I[k]-Pirr(k[0],k[1],x))/abs(I[k]) for k in I)
print(err,err2)
