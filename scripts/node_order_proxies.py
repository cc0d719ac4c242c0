"""CPU: which static ordering of the BVH nodes puts the most visited ones first?  Surface area, the parent's surface
area and depth against the ideal order (the rays' own visit histogram) on the C2 tree -- basis of the layout plan in
DESIGN.md section 6.      python scripts/node_order_proxies.py"""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests')
import pbr_b200
from pbr_b200 import host, scenes
from oracle import oracle as O
from oracle import scene as S
import helpers as Hh
sc = scenes.soup(1_000_000, seed=12345)
flat = host.Scene.from_arrays(sc).build_flat()
nodes = np.ascontiguousarray(flat["nodes"], np.float32).reshape(-1, 8); N = len(nodes)
facesV = np.ascontiguousarray(flat["facesV"], np.uint32); facesN = np.ascontiguousarray(flat["facesN"], np.uint32)
v4 = S.pack_float4(sc["vertices"]); n4 = np.zeros((1,4),np.float32)
p=lambda a:a.ctypes.data_as(C.c_void_p)
class P: camera = S.camera(eye=(0.0,0.0,3.5), center=(0.0,0.0,1.0))
L=O.lib(); D=O.make_defines(img_width=1920,img_height=1080,bvh_num_nodes=N)
def hist(rays):
    rays=np.ascontiguousarray(rays,np.float32); c=np.zeros(N,np.uint32)
    L.oracle_visit_histogram_pruned(p(D),p(nodes),p(facesV),p(facesN),p(v4),p(n4),p(rays),C.c_int64(len(rays)),p(c)); c[0]=0; return c.astype(np.float64)
cp=hist(Hh.primary_rays(P,320,180)); cr=hist(Hh.random_rays(40000,1,-1.0,1.0)); cr2=hist(Hh.random_rays(20000,7,-1.0,1.0))
ext=np.maximum(nodes[:,4:7]-nodes[:,0:3],0); sa=2*(ext[:,0]*ext[:,1]+ext[:,1]*ext[:,2]+ext[:,0]*ext[:,2])
inner=nodes[:,3]<0
# parent / depth from the pre-order structure: left child = i+1 for inner; right child = ? use miss link of left child
depth=np.zeros(N,np.int32); parent=np.zeros(N,np.int64)
# stack-based pre-order reconstruction: inner node has 2 children (left may be skipped -> chain); approximate using miss links:
# node j's parent = the last inner node i<j whose subtree end > j; subtree end of i = miss link of i (or N if -1)
end=np.where(inner, nodes[:,7].astype(np.int64), np.arange(N)+1); 
# for inner nodes on the right spine, miss link = -1 -> climb: treat as N
end=np.where(end<=0, N, end)
stack=[]
for j in range(1,N):
    while stack and end[stack[-1]]<=j: stack.pop()
    if stack: parent[j]=stack[-1]; depth[j]=depth[stack[-1]]+1
    if inner[j]: stack.append(j)
psa=sa[parent]; psa[0]=0
def cover(order,c,ks=(2048,8192,32768,131072)):
    cs=np.cumsum(c[order])/c.sum(); return " / ".join("%4.1f"%(100*cs[k-1]) for k in ks)
ideal_p=np.argsort(-cp,kind='stable'); ideal_r=np.argsort(-cr,kind='stable')
ords={"pre-order (today; lines hold 4 neighbours)":np.arange(N),"surface area":np.argsort(-sa,kind='stable'),"parent's surface area":np.argsort(-psa,kind='stable'),
      "depth (BFS)":np.argsort(depth,kind='stable'),"histogram of 20k other random rays":np.argsort(-cr2,kind='stable'),"ideal (own histogram)":None}
print("share of visits (%) covered by the first 2048 / 8192 / 32768 / 131072 nodes of an ordering: primary | random")
for k,o in ords.items():
    print("%-46s %s | %s"%(k, cover(ideal_p if o is None else o,cp), cover(ideal_r if o is None else o,cr)))
