import sys, time, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'./ultrasonic-communication_b200'); sys.path.insert(0,'./tests')
import usc, synth
from oracle import pyref as R
h=usc.Handle()
print("geometry",h.geometry())
rx=R.RefReceiver()
for name in ["hann","up","down"]:
    t=h.table(name); o=rx.table({"hann":"hann","up":"up_chirp","down":"down_chirp"}[name])
    print(name,"tables equal:",np.array_equal(t,o))
pcm,bits=synth.make_frames(1024)
t0=time.time(); mu,iu,md,idn=rx.demod_frames(pcm,nthreads=8); print("oracle s",time.time()-t0)
gu,giu,gd,gid,gbit=h.demod_frames_host(pcm)
print("mag_up eq",np.array_equal(mu,gu),"idx_up eq",np.array_equal(iu,giu),"mag_dn eq",np.array_equal(md,gd),"idx_dn eq",np.array_equal(idn,gid))
print("max rel diff", np.abs(mu-gu).max()/mu.max(), np.abs(md-gd).max()/md.max())
print("bit acc vs truth", (gbit==bits).mean(), "first idx", giu[:8], iu[:8])
