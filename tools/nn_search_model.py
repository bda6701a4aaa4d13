import sys, numpy as np, collections
sys.path.insert(0,'/root/repo')
import bench
from mrg_slam_b200 import synth
from tests import oraclelib as O
def cloud(i):
    pc=bench.oracle_prefilter(O, synth.scan(synth.HDL64, i))
    c,_=O.voxelgrid(pc, bench.LEAF, 1); return c
A=cloud(108)[:,:3].astype(np.float64); B=cloud(112)[:,:3].astype(np.float64)
T=np.linalg.inv(synth.pose(108))@synth.pose(112)
def grid(P, factor=2.5, xf=0.34):
    mn=P.min(0); ext=P.max(0)-mn
    h=factor*np.sqrt(ext[0]*ext[1]/len(P)); hx=xf*h
    cy=np.floor((P[:,1]-mn[1])/h).astype(int); cz=np.floor((P[:,2]-mn[2])/h).astype(int)
    order=np.lexsort((P[:,0],cy,cz))
    return mn,h,hx,cy,cz,order
def run(factor,xf,nwarps=200):
    mn,h,hx,cy,cz,order=grid(A,factor,xf)
    rowpts=collections.defaultdict(list)
    for i in order: rowpts[(cy[i],cz[i])].append(i)
    rowx={k:A[v,0] for k,v in rowpts.items()}; rowP={k:A[v] for k,v in rowpts.items()}
    # source in its own cell order
    _,_,_,_,_,so=grid(B,2.5,0.34)
    Bt=((T[:3,:3]@B[so].T).T+T[:3,3])
    rng=np.random.default_rng(1)
    starts=rng.choice(len(Bt)//32-1,nwarps,replace=False)*32
    tot_cand=0; costA=0; costB=0; costC=0; costD=0; nq=0; rows_tot=0
    for s0 in starts:
        lanes=[]
        for q in Bt[s0:s0+32]:
            qy=(q[1]-mn[1])/h; qz=(q[2]-mn[2])/h
            iy=int(np.floor(qy)); iz=int(np.floor(qz)); fy=qy-iy; fz=qz-iz
            best=[np.inf]; visits=[]  # (slot, ncell, nleft, nright)
            def visit(y,z,g2,slot):
                k=(y,z)
                if k not in rowx: return
                xs=rowx[k]; P=rowP[k]
                cx=np.floor((q[0]-mn[0])/hx); lo=np.searchsorted(xs,mn[0]+cx*hx); hi=np.searchsorted(xs,mn[0]+(cx+1)*hx)
                nc=hi-lo
                if nc: best[0]=min(best[0],((P[lo:hi]-q)**2).sum(1).min())
                nl=0; j=lo-1
                while j>=0 and (q[0]-xs[j])**2+g2<=best[0]:
                    nl+=1; best[0]=min(best[0],((P[j]-q)**2).sum()); j-=1
                nr=0; j=hi
                while j<len(xs) and (q[0]-xs[j])**2+g2<=best[0]:
                    nr+=1; best[0]=min(best[0],((P[j]-q)**2).sum()); j+=1
                visits.append((slot,nc,nl,nr))
            visit(iy,iz,0.0,0)
            r=1
            while True:
                slot=0
                for dz in range(-r,r+1):
                    for dy in range(-r,r+1):
                        if max(abs(dy),abs(dz))!=r: continue
                        slot+=1
                        gy=(dy-fy if dy>0 else (fy-(dy+1) if dy<0 else 0)); gz=(dz-fz if dz>0 else (fz-(dz+1) if dz<0 else 0))
                        g2=(gy*gy+gz*gz)*h*h
                        if g2<best[0]: visit(iy+dy,iz+dz,g2,(r,slot))
                if best[0]<=(r*h)**2 or r>80: break
                r+=1
            lanes.append(visits)
        tots=[sum(v[1]+v[2]+v[3] for v in l) for l in lanes]
        tot_cand+=sum(tots); nq+=32; rows_tot+=sum(len(l) for l in lanes)
        costA+=max(tots)
        # B: per-lane row lists
        kmax=max(len(l) for l in lanes)
        for k in range(kmax):
            costB+=max((l[k][1] if len(l)>k else 0) for l in lanes)+max((l[k][2] if len(l)>k else 0) for l in lanes)+max((l[k][3] if len(l)>k else 0) for l in lanes)
        # C: lock-step by slot
        slots=collections.defaultdict(lambda:[0,0,0])
        for l in lanes:
            for v in l:
                s=slots[v[0]]; s[0]=max(s[0],v[1]); s[1]=max(s[1],v[2]); s[2]=max(s[2],v[3])
        costC+=sum(sum(s) for s in slots.values())
        slotsD=collections.defaultdict(int)
        for l in lanes:
            for v in l: slotsD[v[0]]=max(slotsD[v[0]], v[1]+v[2]+v[3])
        costD+=sum(slotsD.values())
    print(f'factor {factor} xf {xf}: cands/query {tot_cand/nq:.1f} rows/query {rows_tot/nq:.2f}  per-warp candidate-steps: flattened {costA/nwarps:.1f}  per-lane-rows {costB/nwarps:.1f}  lockstep {costC/nwarps:.1f}  lockstep-rows+merged-row-loop {costD/nwarps:.1f}   ideal {tot_cand/nq:.1f}')
for f,xf in [(2.5,0.34),(2.5,0.17)]:
    run(f,xf)
