import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import pbr_b200, helpers as Hh
from oracle import oracle as O
t=time.time(); s = pbr_b200.scenes.soup(1000000); b = O.build_bvh(s); print("bvh build s", time.time()-t, b["info"], flush=True)
p = Hh.Prepared(s, 1920, 1080, eye=(0,0,3.5), bvh=b)
dev = pbr_b200.Device(0); print(dev.deviceInfo())
ds = Hh.DeviceScene(dev, p)
rays = Hh.primary_rays(p, 1920, 1080)
n = len(rays)
rb = dev.createBuffer(rays); hb = dev.createEmptyBuffer(n*16)
for i in range(3):
    dev.stats(reset=True)
    dev.traceDevice(ds.bufBVH, ds.bufFacesV, ds.bufVertices, rb, n, hb); dev.finish()
    ms = dev.kernelTimeMs(ds.kernel); st = dev.stats(reset=True)
    print("primary trace: %.3f ms, %.1f Mrays/s, nodes/ray %.1f tris/ray %.2f, algGB/s %.1f" % (ms, n/ms/1e3, st[2]/n, st[3]/n, (32*st[2]+64*st[3]+48*n)/ms/1e6), flush=True)
hits = dev.readBuffer(hb, n*16, np.uint8).view(Hh.O.HIT_DTYPE)
sub = slice(0, n, 64)
want,_ = p.oracle_trace(rays[sub])
print("parity subsample:", np.array_equal(want["hitFace"], hits["hitFace"][sub]), np.array_equal(want["leaf"], hits["leaf"][sub]), np.array_equal(want["t"].view(np.uint32), hits["t"][sub].view(np.uint32)))
# random (incoherent) rays
rr = Hh.random_rays(2000000, 1, -1.0, 1.0); rb2 = dev.createBuffer(rr); hb2 = dev.createEmptyBuffer(len(rr)*16)
for i in range(3):
    dev.stats(reset=True)
    dev.traceDevice(ds.bufBVH, ds.bufFacesV, ds.bufVertices, rb2, len(rr), hb2); dev.finish()
    ms = dev.kernelTimeMs(ds.kernel); st = dev.stats(reset=True)
    print("random trace: %.3f ms, %.1f Mrays/s, nodes/ray %.1f tris/ray %.2f algGB/s %.1f" % (ms, len(rr)/ms/1e3, st[2]/len(rr), st[3]/len(rr), (32*st[2]+64*st[3]+48*len(rr))/ms/1e6), flush=True)
# frames
for mode in (0, 1):
    dev.setPipeline(mode)
    for i in range(4):
        dev.stats(reset=True)
        t0=time.time(); img,_ = ds.frames(1, first=i); wall=time.time()-t0
        ms = dev.kernelTimeMs(ds.kernel); st = dev.stats(reset=True)
        print("pipeline %d frame %d: %.3f ms kernel (%.1f ms wall), rays %d, %.1f Mrays/s, nodes/ray %.1f" % (mode, i, ms, wall*1e3, st[0], st[0]/ms/1e3, st[2]/max(1,st[0])), flush=True)
dev.setPipeline(0)
