import sys,time,os,tempfile
sys.path.insert(0,"/root/repo")
import numpy as np, torch
from lut_renderer_b200 import synth
from lut_renderer_b200.stage import ColourStage, DeviceContext
ctx=DeviceContext.get(0)
kinds = {0: dict(pix_fmt="yuv420p", w=1920, h=1080, lut=33), 1: dict(pix_fmt="yuv422p10le", w=3840, h=2160, lut=65)}
tmp=tempfile.mkdtemp(); paths={}
t0=time.perf_counter()
for i in range(8):
    n=kinds[i%2]["lut"]
    table=(0.9*synth.look_lut(n)+0.1*synth.random_lut(n,100+i)).astype(np.float32)
    paths[i]=synth.write_cube(os.path.join(tmp,f"l{i}.cube"),table)
print("write cubes", time.perf_counter()-t0)
fpc=24
stages,pins={},{}
for k,c in kinds.items():
    st=ColourStage(c["w"],c["h"],c["pix_fmt"]); stages[k]=st
    pi,po=ctx.pinned_buffer(fpc*st.in_frame_bytes),ctx.pinned_buffer(fpc*st.out_frame_bytes)
    pi.array[:]=np.tile(synth.natural_frame(c["pix_fmt"],c["w"],c["h"]),fpc); pins[k]=(pi,po)
for rep in range(3):
    for i in range(8):
        t0=time.perf_counter()
        with open(paths[i],"rb") as f: data=f.read()
        t1=time.perf_counter()
        lut=ctx.lut_from_text(data)
        t2=time.perf_counter()
        stages[i%2].run_host(lut,pins[i%2][0],pins[i%2][1],fpc)
        t3=time.perf_counter()
        lut.close()
        t4=time.perf_counter()
        print(rep,i,"read %.1f parse+upload %.1f run_host %.1f close %.1f ms"%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3), len(data))
