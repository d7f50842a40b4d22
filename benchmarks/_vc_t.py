import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
from activezero_b200 import ops
from benchmarks.kernel_sweep import time_ms
DEV="cuda:0"
L, R = torch.randn(2,32,136,240,device=DEV), torch.randn(2,32,136,240,device=DEV)
w = torch.randn(32,64,3,3,3,device=DEV)*0.05
wp = ops.pack_volume_conv_weight(w)
for dbg in (0,1,4,8,16,12,28):
    os.environ["AZ_VCONV_DBG"]=str(dbg)
    print(json.dumps({"dbg":dbg,"ms":round(time_ms(lambda: ops.volume_conv0(L,R,wp,48), flush=True),4)}))
