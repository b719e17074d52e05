import numpy as np
a,b,c = np.deg2rad([2.0,1.0,3.0])
Rz=np.array([[1,0,0],[0,np.cos(a),-np.sin(a)],[0,np.sin(a),np.cos(a)]])
Ry=np.array([[np.cos(b),0,np.sin(b)],[0,1,0],[-np.sin(b),0,np.cos(b)]])
Rx=np.array([[np.cos(c),-np.sin(c),0],[np.sin(c),np.cos(c),0],[0,0,1]])
Mg=np.eye(4); Mg[:3,:3]=Rz@Ry@Rx@np.diag([1.03,0.97,1.1]); Mg[:3,3]=[0.4,-1.2,2.3]
M90=np.array([[1.0,0,0,3.5],[0,0,-1.288,2040.0],[0,1.288,0,-20.0],[0,0,0,1]])
M90t=M90.copy(); M90t[0,1:3]=[0.02,-0.015]; M90t[1,0],M90t[2,0]=0.03,-0.02
rng=np.random.default_rng(0)
def wavefronts(M, lane_shape, swap, pitch, trials=400):
    # lane_shape (la, lb): la lanes along the "lane axis", lb along the other; lane axis = o2 (or o1 if swap)
    la, lb = lane_shape
    tot=0
    for _ in range(trials):
        o0,o1,o2 = rng.integers(0,100), rng.integers(0,2000), rng.integers(0,1200)
        l = np.arange(32); i = l % la; j = l // la
        if swap: d1, d2 = i, j
        else: d1, d2 = j, i
        y = M[1,3]+M[1,0]*o0+M[1,1]*(o1+d1)+M[1,2]*(o2+d2)
        x = M[2,3]+M[2,0]*o0+M[2,1]*(o1+d1)+M[2,2]*(o2+d2)
        z = M[0,3]+M[0,0]*o0+M[0,1]*(o1+d1)+M[0,2]*(o2+d2)
        addr = np.floor(y).astype(int)*pitch + np.floor(x).astype(int)   # slot offset multiple of 32: ignore z
        # distinct addresses hitting the same bank conflict; same address broadcast
        u = np.unique(addr + 1000000*np.floor(z).astype(int))
        ua = np.unique(addr)
        banks = ua % 32
        tot += np.bincount(banks, minlength=32).max()
    return tot/trials
for name,M,swap in (("general",Mg,False),("rot90tilt",M90t,True),("rot90",M90,True)):
    print(name, "x per lane:", M[2,1] if swap else M[2,2], "y per lane:", M[1,1] if swap else M[1,2])
    for shape in ((32,1),(16,2),(8,4)):
        res = {p: round(wavefronts(M,shape,swap,p),2) for p in range(64,96,4)}
        print("  lanes",shape, res)
