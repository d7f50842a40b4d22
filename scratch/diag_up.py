import sys, torch
sys.path.insert(0,'.')
from activezero_b200 import ops
torch.manual_seed(30)
for shape,size,scale in [((2,1,48,16,32),(192,64,128),8.0), ((1,1,48,16,32),(192,64,128),8.0), ((1,1,48,16,32),(192,64,128),1.0), ((1,1,24,16,32),(96,64,128),8.0)]:
    low=(torch.randn(shape)*scale).cuda()
    out=ops.upsample_soft_argmin(low,size)
    up=torch.nn.functional.interpolate(low.double(),size,mode='trilinear',align_corners=False).squeeze(1)
    d=torch.arange(size[0],device='cuda',dtype=torch.float64).view(1,-1,1,1)
    ref=(torch.softmax(up,1)*d).sum(1,keepdim=True)
    err=(out.double()-ref).abs()
    print(shape,size,scale,'max err',float(err.max()),'nan',int(torch.isnan(out).sum()), 'argmax err idx', [int(v) for v in (err==err.max()).nonzero()[0]] if not torch.isnan(err.max()) else None)
