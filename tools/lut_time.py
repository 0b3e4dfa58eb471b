import sys,time,ctypes
sys.path.insert(0,"/root/repo")
import numpy as np, torch
from lut_renderer_b200 import synth, _native as N
from lut_renderer_b200.stage import ColourStage, DeviceContext
ctx=DeviceContext.get(0)
t=synth.look_lut(65).astype(np.float32)
txt=synth.cube_text(t).encode()
for i in range(4):
    t0=time.perf_counter(); lut=ctx.lut_from_text(txt); torch.cuda.synchronize(); t1=time.perf_counter()
    lut.close(); torch.cuda.synchronize(); t2=time.perf_counter()
    print("lut_from_text", round((t1-t0)*1e3,1), "close", round((t2-t1)*1e3,1))
for i in range(3):
    t0=time.perf_counter(); lut=ctx.lut_from_table(t); torch.cuda.synchronize(); t1=time.perf_counter(); lut.close()
    print("lut_from_table", round((t1-t0)*1e3,1))
st=ColourStage(3840,2160,"yuv422p10le")
n=24
pi,po=ctx.pinned_buffer(n*st.in_frame_bytes),ctx.pinned_buffer(n*st.out_frame_bytes)
pi.array[:]=np.tile(synth.natural_frame("yuv422p10le",3840,2160),n)
lut=ctx.lut_from_table(t)
for i in range(4):
    t0=time.perf_counter(); st.run_host(lut,pi,po,n); t1=time.perf_counter()
    print("run_host 24 frames", round((t1-t0)*1e3,1), "ms", round(n/(t1-t0)), "fps")
lut2=ctx.lut_from_text(txt)
for i in range(3):
    t0=time.perf_counter(); st.run_host(lut2 if i%2 else lut,pi,po,n); t1=time.perf_counter()
    print("run_host alternating luts", round((t1-t0)*1e3,1), "ms")
