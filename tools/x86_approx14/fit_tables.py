import numpy as np
def fit(name, nseg_bits, q):
    T=np.fromfile(name+".bin",dtype=np.uint32).astype(np.int64)
    v=(T>>7)
    L=65536>>nseg_bits
    t=np.arange(L,dtype=np.int64)>>q
    res=[]
    for seg in range(1<<nseg_bits):
        vs=v[seg*L:(seg+1)*L]
        s=9
        b=(vs[0]-vs[-1])/t[-1]
        c1c=int(round(b*(1<<s)))
        sols=[]
        for c1 in range(max(0,c1c-80),c1c+81):
            lo=(vs*(1<<s)+c1*t).max(); hi=((vs+1)*(1<<s)+c1*t).min()
            if lo<hi: sols.append((c1,int(lo)))
        assert len(sols)==1
        res.append(sols[0])
    return res
rcp=fit("rcp",6,0); r1=fit("rsq1",5,1); r2=fit("rsq2",5,1)
def emit(name, rows):
    print(f"static const uint32_t {name}[{len(rows)}][2] = {{")
    for i in range(0,len(rows),4):
        print("    "+" ".join(f"{{0x{c0:08X}u, {c1}u}}," for c1,c0 in rows[i:i+4]))
    print("};")
emit("kRcp14", rcp)
emit("kRsqrt14", r1+r2)
