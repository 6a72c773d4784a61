import torch, sys, os
sys.path.insert(0, os.getcwd())
from multi_style_transfer_gan_b200.enhanced_generator import EnhancedGenerator
from multi_style_transfer_gan_b200.stylize import MultiStyleStylizer
gens=[]
for s in range(4):
    torch.manual_seed(s); gens.append(EnhancedGenerator(16,1).cuda())
w=[0.4,0.3,0.2,0.1]
torch.manual_seed(1234)
size=int(sys.argv[1]) if len(sys.argv)>1 else 1024
x=(torch.rand(2,3,size,size)*2-1).cuda()
ss=sys.argv[2]=="1"; gr=sys.argv[3]=="1"
st=MultiStyleStylizer(gens,precision="bf16",micro_batch=2,style_streams=ss,use_graph=gr)
y=st(x,w)
torch.cuda.synchronize()
with torch.no_grad():
    singles=[g(x) for g in gens]
ref=sum(wi*s for wi,s in zip(w,singles))
y2=st(x,w)
torch.cuda.synchronize()
print("streams",ss,"graph",gr,"first call diff",float((y-ref).abs().max()),"2nd call",float((y2-ref).abs().max()))
if float((y-ref).abs().max())>1e-3:
    d=(y-ref).abs()
    idx=(d>1e-3).nonzero()
    print("n bad",idx.shape[0],"first",idx[0].tolist(),"last",idx[-1].tolist())
    for n in range(2):
        dn=d[n].amax(0)
        rows=(dn>1e-3).any(1).nonzero().flatten(); cols=(dn>1e-3).any(0).nonzero().flatten()
        if rows.numel(): print("img",n,"rows",int(rows.min()),int(rows.max()),"cols",int(cols.min()),int(cols.max()), "count",int((dn>1e-3).sum()))
