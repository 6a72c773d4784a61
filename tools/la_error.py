import torch, torch.nn.functional as F, sys
sys.path.insert(0,'/root/repo')
from multi_style_transfer_gan_b200 import ops
torch.manual_seed(0)
for C,H in ((32,32),(64,64),(128,32),(256,16)):
    N=2
    qkv=torch.randn(N,3*C,H,H,device='cuda').bfloat16().float()
    qq,kk,vv=qkv.double().chunk(3,dim=1)
    def win(u): return u.reshape(N,C,H//4,4,H//4,4).permute(0,2,4,1,3,5).reshape(-1,C,16)
    qq,kk,vv=win(qq),win(kk),win(vv)
    a=torch.softmax(F.normalize(qq,dim=1)@F.normalize(kk,dim=1).transpose(1,2),dim=-1)
    o=(a@vv).reshape(N,H//4,H//4,C,4,4).permute(0,3,1,4,2,5).reshape(N,C,H,H)
    out=ops.local_attn_fwd(qkv.permute(0,2,3,1).contiguous().bfloat16()).float().permute(0,3,1,2).double()
    err=(out-o)
    print(C, 'max', float(err.abs().max()/o.abs().max()), 'relL2', float(err.norm()/o.norm()), 'bf16-of-exact relL2', float((o.float().bfloat16().double()-o).norm()/o.norm()))
