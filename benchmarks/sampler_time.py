import sys, torch, numpy as np
sys.path.insert(0,'/root/repo')
from qldpcsim_b200 import pcmlibrary, simulator
for code in ['LP118_0','T']:
    Hx,Hz=pcmlibrary.by_name(code)
    pipe=simulator.Pipeline(Hx,Hz,0.05,'MS',50,'L')
    pipe.sample_device(1000000,1,0); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5): pipe.sample_device(1000000,1,i*1000000)
    e1.record(); torch.cuda.synchronize()
    print(code,'sample 1M shots:', e0.elapsed_time(e1)/5,'ms')
