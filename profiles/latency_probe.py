import time, numpy as np, torch
from multi_orbslam3_b200 import orbx, synth
W,H=752,480
ex = orbx.ORBextractor(1000,1.2,8,20,7,max_width=W,max_height=H,max_batch=1)
m = orbx.ORBmatcher(0.9, True, max_keypoints=ex.cap, max_batch=1)
fr = synth.rects_stream(W,H,16,seed=3)
hf = torch.from_numpy(fr).pin_memory().numpy()
cap=ex.cap
out = {"kps": torch.empty((1,cap,7),dtype=torch.float32).pin_memory().numpy().view(np.uint8).reshape(1,cap,28).view(orbx.KP_DTYPE).reshape(1,cap),
       "desc": torch.empty((1,cap,32),dtype=torch.uint8).pin_memory().numpy(), "n": torch.empty(1,dtype=torch.int32).pin_memory().numpy(),
       "mono": torch.empty(1,dtype=torch.int32).pin_memory().numpy(), "matches12": torch.empty((1,cap),dtype=torch.int32).pin_memory().numpy(),
       "nmatches": torch.empty(1,dtype=torch.int32).pin_memory().numpy()}
bounds=(0.0,float(W),0.0,float(H))
for i in range(20): orbx.extract_match_batch(ex,m,hf[i%16:i%16+1],(0,0),bounds,100,out)
ex.profile(1)
ts=[]
for i in range(300):
    t=time.perf_counter(); orbx.extract_match_batch(ex,m,hf[i%16:i%16+1],(0,0),bounds,100,out); ts.append(time.perf_counter()-t)
ms, nb = ex.profile(0)
print("p50 %.3f ms"%(np.percentile(ts,50)*1e3), "stages per call (ms):", [round(x/nb,4) for x in ms], "sum %.3f"%(sum(ms)/nb))
# extraction only (class API)
ts=[]
for i in range(200):
    t=time.perf_counter(); ex(hf[i%16], None, (0,0)); ts.append(time.perf_counter()-t)
print("extract only p50 %.3f ms"%(np.percentile(ts,50)*1e3))
