"""How much of the Delaunay graph of flame::Flame::update survives from one frame to the next?
Runs the oracle-side pipeline mirror (CPU only) on the C2-shaped synthetic stream (VGA, detection window 8 px,
poseframe every 6th frame) and prints, per frame: vertices / edges, vertices added and removed, edges that
appeared and disappeared.  Result (profiles/r2_topology_churn.md): every steady-state frame changes the vertex
set and ~2 % of the edges, so no topology and no host-built solver plan can be kept -- the reason the
triangulation, the graph sync and the solver plan moved onto the device in round 2."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from flame_ros_b200 import synth, capi
from oracle import oracle as O
from pipeline_mirror import MirrorFlame
O.load(); capi.load_library()
W,H=640,480; K=synth.K_VGA
n=int(sys.argv[1]) if len(sys.argv)>1 else 50
win=int(sys.argv[2]) if len(sys.argv)>2 else 8
nslots=int(sys.argv[3]) if len(sys.argv)>3 else 8
sc=synth.Scene(1, tex_size=1024)
poses=synth.stream_poses(n, step=0.01)
up=capi.default_update_params(); up.detection_win_size=win; up.iters=5
mup=type("UP",(),{})()
for nm,_ in up._fields_: setattr(mup,nm,getattr(up,nm))
p=O.NLTGV2Params()
for nm,_ in p._fields_: setattr(p,nm,getattr(up.rparams,nm))
mup.rparams=p
maxF=8192; maxV=8192
mir=MirrorFlame(O,capi,W,H,K,nslots,maxF,maxV,mup)
prev_vf=None; prev_edges=None
for k in range(n):
    img=sc.render(K,poses[k],W,H)[0]
    t0=time.time()
    ok=mir.update(k/30.0,k,poses[k],img,k%6==0)
    if not ok: print(k,'no update', int(mir.alive.sum())); continue
    vf=list(mir.vert_feat); es=set((vf[a],vf[b]) for a,b in mir.edges)
    if prev_vf is not None:
        a=set(vf); b=set(prev_vf)
        print(k, 'V',len(vf),'E',len(es),'alive',int(mir.alive.sum()),'added',len(a-b),'removed',len(b-a),'edges+',len(es-prev_edges),'edges-',len(prev_edges-es), 'pf' if k%6==0 else '', '%.2fs'%(time.time()-t0))
    prev_vf=vf; prev_edges=es
