import torch, sys, os
sys.path.insert(0, os.getcwd())
from multi_style_transfer_gan_b200 import ops
torch.manual_seed(0)
for C,N,H,W in [(64,4,256,256),(128,4,128,128),(64,2,64,64)]:
    x = torch.randn(N,H,W,C,device='cuda').bfloat16()
    wq = (torch.randn(3*C*C,device='cuda')*(2.0/C)**0.5).bfloat16()
    wp = (torch.randn(C*C,device='cuda')*(1.0/C)**0.5).bfloat16()
    bq = torch.randn(3*C,device='cuda')*0.1; bp = torch.randn(C,device='cuda')*0.1
    st = ops.instnorm_stats(x)
    outs=[]
    for i in range(4):
        outs.append(ops.la_stage_fwd(x,wq,bq,wp,bp,in_stats=st,in_act=ops.ACT_RELU))
    torch.cuda.synchronize()
    for i in range(1,4):
        d=(outs[i].float()-outs[0].float()).abs()
        print(C,N,H,W,"run",i,"max diff",float(d.max()),"n diff",int((d>0).sum()))
    o1 = ops.la_stage_fwd(x[:1].contiguous(),wq,bq,wp,bp,in_stats=st[:1].contiguous(),in_act=ops.ACT_RELU)
    d=(o1.float()-outs[0][:1].float()).abs()
    print(C,"single image vs batch: max diff",float(d.max()),"n diff",int((d>0).sum()))
    if int((d>0).sum()):
        idx=(d>0).nonzero()
        print(idx[:10].tolist(), idx[-3:].tolist())
