import numpy as np, ctypes, itertools
rng = np.random.default_rng(1)
libm = ctypes.CDLL('libm.so.6'); libm.fma.restype = ctypes.c_double; libm.fma.argtypes=[ctypes.c_double]*3
fma = libm.fma
def rot(rng):
    q = rng.normal(size=4); q/=np.linalg.norm(q)
    w,x,y,z=q
    return np.array([[1-2*(y*y+z*z),2*(x*y-z*w),2*(x*z+y*w)],[2*(x*y+z*w),1-2*(x*x+z*z),2*(y*z-x*w)],[2*(x*z-y*w),2*(y*z+x*w),1-2*(x*x+y*y)]])
# enumerate expression trees: terms t0,t1,t2 = a_i*v_i
cands = {}
for perm in itertools.permutations(range(3)):
    i,j,k = perm
    # ((ti op tj) op tk)
    def mk(i,j,k,m1,m2):
        def f(a,v):
            if m1==0: s = a[i]*v[i] + a[j]*v[j]
            elif m1==1: s = fma(a[j],v[j], a[i]*v[i])
            elif m1==2: s = fma(a[j],v[j], fma(a[i],v[i],0.0))
            if m2==0: return s + a[k]*v[k]
            elif m2==1: return fma(a[k],v[k], s)
            elif m2==2: return s + fma(a[k],v[k],0.0)
        return f
    for m1 in range(3):
        for m2 in range(3):
            cands[f'{i}{j}{k}_m{m1}{m2}'] = mk(i,j,k,m1,m2)
N=1000
hits = {k:0 for k in cands}
tot=0
for _ in range(N):
    R = rot(rng); v = rng.normal(size=3)*10
    A = np.ascontiguousarray(R)
    y = A.dot(v)
    for r in range(3):
        tot+=1
        for k,f in cands.items():
            hits[k] += int(f(A[r],v) == y[r])
best = sorted(hits.items(), key=lambda kv:-kv[1])[:8]
print([(k, round(h/tot,3)) for k,h in best])
# matmul variant and also np.matmul(A, v)
hits = {k:0 for k in cands}; tot=0
for _ in range(N):
    R = rot(rng); v = rng.normal(size=3)*10
    A = np.ascontiguousarray(R)
    y = np.matmul(A, v)
    for r in range(3):
        tot+=1
        for k,f in cands.items():
            hits[k] += int(f(A[r],v) == y[r])
best = sorted(hits.items(), key=lambda kv:-kv[1])[:4]
print('matmul', [(k, round(h/tot,3)) for k,h in best])
