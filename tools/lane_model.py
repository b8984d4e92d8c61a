"""Lane-utilisation model of the forward blend's pass 2 (dev tool; CPU only, uses the oracle -- test infrastructure).

For a sample of tiles of the f3d-like scene it builds every pixel's contributing (pixel, record) pairs with the
oracle's preprocess + binning and a float64 evaluation of the blend's alpha test and saturation, then simulates the
warp-level trip loop (two survivors per lane and trip) for a lookahead of L chunks: L = 0 is the per-chunk lock-step
of round 1, L = 1 the shipped kernel.

    python tools/lane_model.py [view (-1 = canonical)] [chunk size]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch, cases, oracle_cpu
view = int(sys.argv[1]) if len(sys.argv)>1 else 2
CH = int(sys.argv[2]) if len(sys.argv)>2 else 128
c = cases.f3d_case(0, 256, 256, view if view>=0 else None)
cn = oracle_cpu.case_to_numpy(c)
pre = oracle_cpu.preprocess(cn)
b = oracle_cpu.binning(256,256,pre["means2D"],pre["depths"],pre["radii"],pre["tiles_touched"])
W=H=256; fx = W/(2*cn["tanfovx"]); fy=H/(2*cn["tanfovy"])
v2g = pre["view2gaussian"].astype(np.float64); wop = pre["conic_opacity"][:,3].astype(np.float64)
rng = b["ranges"].astype(np.int64); pl = b["point_list"].astype(np.int64)
Ls=[0,1,2,99]
trips={L:0 for L in Ls}; work_tot=0
def simulate(cnt, L):
    # cnt [32 lanes, nch] survivors per lane per chunk; each trip a lane consumes up to 2 survivors from its current chunk
    # (a trip never spans two chunks for one lane, like the kernel's per-chunk walk would if generalised: allow spanning)
    nl,nch = cnt.shape
    rem = cnt.copy().astype(np.int64)
    cur = np.zeros(nl, dtype=np.int64)
    t=0
    # advance lanes past empty chunks
    while True:
        # lanes move to next chunk with work
        for l in range(nl):
            while cur[l] < nch and rem[l,cur[l]]==0: cur[l]+=1
        if (cur>=nch).all(): break
        base = cur.min()
        # active lanes: those whose current chunk within window
        act = (cur < nch) & (cur <= base+L)
        # one trip: each active lane consumes 2 (may take from the next chunk too if within window: approximate by 2 from the same chunk then spill)
        t+=1
        for l in np.nonzero(act)[0]:
            need=2
            k=cur[l]
            while need>0 and k<nch and k<=base+L:
                d=min(need, rem[l,k]); rem[l,k]-=d; need-=d
                if rem[l,k]==0: k+=1
            # cur updated at loop top
    return t
rs = np.random.RandomState(0)
tiles = rs.choice(256, 48, replace=False)
for tile in tiles:
    ty, tx = divmod(tile,16)
    lo,hi = rng[tile]; n=hi-lo
    if n<=0: continue
    ids = pl[lo:hi]
    q = v2g[ids]; w = wop[ids]
    px = tx*16+np.arange(16); py = ty*16+np.arange(16)
    rx = ((px+0.5-W/2.)/fx)[None,:].repeat(16,0).reshape(-1); ry=((py+0.5-H/2.)/fy)[:,None].repeat(16,1).reshape(-1)
    RX=rx[:,None]; RY=ry[:,None]
    n0 = q[None,:,0]*RX+q[None,:,1]*RY+q[None,:,2]; n1=q[None,:,1]*RX+q[None,:,3]*RY+q[None,:,4]; n2=q[None,:,2]*RX+q[None,:,4]*RY+q[None,:,5]
    AA = n0*RX+n1*RY+n2; BB=2*(q[None,:,6]*RX+q[None,:,7]*RY+q[None,:,8]); CC=q[None,:,9]
    t = -BB/(2*AA); mv = -(BB/AA)*(BB/4)+CC
    alpha = np.minimum(0.99, w[None,:]*np.exp(np.minimum(0,-0.5*mv)))
    contrib = (t>0.2)&(alpha>=1/255.)
    a = np.where(contrib, alpha, 0.0)
    T = np.cumprod(1-a, axis=1)
    stop = (T<1e-4)
    first_stop = np.where(stop.any(1), stop.argmax(1), n)
    live = np.arange(n)[None,:] < first_stop[:,None]
    work = (contrib & live).reshape(16,16,n)
    nch=(n+CH-1)//CH
    pad = nch*CH-n
    wk = np.pad(work, ((0,0),(0,0),(0,pad))).reshape(16,16,nch,CH).sum(3)   # [16,16,nch]
    for wy in range(4):
        for wx in range(2):
            cnt = wk[wy*4:(wy+1)*4, wx*8:(wx+1)*8].reshape(32,nch)
            work_tot += cnt.sum()
            for L in Ls:
                trips[L]+=simulate(cnt,L)
for L in Ls:
    print(f"view {view} chunk {CH} lookahead {L} chunks: active lanes {work_tot/(trips[L]*2*32)*32:.1f}/32")
