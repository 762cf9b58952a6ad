import sys, time
sys.path[:0]=["/root/repo","/root/repo/ultrasonic-communication_b200","/root/repo/tests"]
import torch, usc, bench
dev=torch.device("cuda",0)
h=usc.Handle()
pcm,_=bench.make_device_frames(torch,h,bench.NFRAMES,dev,0)
F=bench.NFRAMES
host=torch.empty((F,2048),dtype=torch.int32).pin_memory(); host.copy_(pcm)
outs=[torch.empty(F,dtype=torch.float32).pin_memory() for _ in range(2)]+[torch.empty(F,dtype=torch.int32).pin_memory() for _ in range(2)]
hb=torch.empty(F,dtype=torch.uint8).pin_memory()
# raw H2D
d=torch.empty_like(pcm)
for _ in range(2): d.copy_(host,non_blocking=True); torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(3): d.copy_(host,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/3
print("raw H2D 1.275 GB: %.2f ms  %.1f GB/s"%(dt*1e3, F*8192/dt/1e9))
for cf in (1024,2048,4096,8192,16384,32768):
    h.host_workspace(cf)
    h.demod_frames_hostbuf(host,usc.PCM_I32,F,outs[0],outs[2],outs[1],outs[3],hb)
    t0=time.perf_counter()
    for _ in range(3): h.demod_frames_hostbuf(host,usc.PCM_I32,F,outs[0],outs[2],outs[1],outs[3],hb)
    dt=(time.perf_counter()-t0)/3
    print("chunk %5d: %.2f ms  %.2f Msym/s  %.1f GB/s"%(cf, dt*1e3, F/dt/1e6, F*8192/dt/1e9))
