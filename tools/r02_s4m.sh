#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none python -c "
import os
for k,v in sorted(os.environ.items()):
    if any(s in k.upper() for s in ('INJECT','NSIGHT','NV_','PROFIL','CUPTI','NCU','LD_PRELOAD','COMPUTE')): print(k,'=',v)
import torch; x=torch.zeros(4,device='cuda')+1; torch.cuda.synchronize()
print(open('/proc/self/maps').read().count('nsight'), [l.split()[-1] for l in open('/proc/self/maps') if 'nsight' in l.lower() or 'inject' in l.lower()][:5])
" > gpurun_out/ncu_env.log 2>&1
