import sys, time, threading
sys.path.insert(0,'/root/repo')
import numpy as np, pynvml, synth
from sternheimergw_b200 import Context, select_solver_type
pynvml.nvmlInit(); h=pynvml.nvmlDeviceGetHandleByIndex(0)
samples=[]; stop=False
def poll():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h)/1e3))
        time.sleep(0.01)
syn=synth.preset("si64"); ctx=Context(0); ctx.install_system(syn)
fiu=synth.imag_freqs(32); igu=np.arange(1,1901,dtype=np.int32); cfg=select_solver_type(priority=(1,3),threshold=1e-4)
ctx.set_profiling(True)
ctx.coulomb(cfg,2,1900,8,igu,fiu)
th=threading.Thread(target=poll); th.start()
t0=time.perf_counter()
for s in range(6):
    ctx.coulomb(cfg,2,1900,8,igu,fiu)
    print(s, round(time.perf_counter()-t0,3), {k:(round(v["ms"],1), v["regions"]) for k,v in ctx.profile().items()})
stop=True; th.join()
a=np.array([(t-t0,c,p) for t,c,p in samples])
for lo in np.arange(0,a[-1,0],0.25):
    m=(a[:,0]>=lo)&(a[:,0]<lo+0.25)
    if m.any(): print("t %.2f clk med %d min %d pow med %.0f max %.0f n %d"%(lo,np.median(a[m,1]),a[m,1].min(),np.median(a[m,2]),a[m,2].max(),m.sum()))
