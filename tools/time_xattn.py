import os, sys, torch
sys.path.insert(0, '/root/repo')
from reformer_tts_b200 import ops
B,T,S,H=20,1024,256,8; D=H*64
q=torch.randn(B,T,D,device='cuda').bfloat16(); kv=torch.randn(B,S,2*D,device='cuda').bfloat16(); do=torch.randn(B,T,D,device='cuda').bfloat16()
keep=torch.ones(B,S,dtype=torch.uint8,device='cuda'); keep[:,200:]=0
seed=torch.tensor([12345],dtype=torch.int64,device='cuda')
k,v=kv[...,:D],kv[...,D:]
def t(f,n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)*1e3/n
out,lse=ops.xattn_fwd(q,k,v,keep,H,0.125,0.15,seed); delta=ops.lsh_delta(do,out,H)
print('xattn_fwd us', t(lambda: ops.xattn_fwd(q,k,v,keep,H,0.125,0.15,seed)), ' no dropout', t(lambda: ops.xattn_fwd(q,k,v,keep,H,0.125,0.0,None)), ' no mask', t(lambda: ops.xattn_fwd(q,k,v,None,H,0.125,0.0,None)))
print('xattn_bwd us', t(lambda: ops.xattn_bwd(q,k,v,keep,H,0.125,0.15,seed,do,lse,delta)), ' no dropout', t(lambda: ops.xattn_bwd(q,k,v,keep,H,0.125,0.0,None,do,lse,delta)))
import sys as _s
if len(_s.argv) > 1: raise SystemExit
import torch.nn.functional as F
ql=q.view(B,T,H,64).transpose(1,2).detach().requires_grad_(True); kl=k.reshape(B,S,H,64).transpose(1,2).detach().requires_grad_(True); vl=v.reshape(B,S,H,64).transpose(1,2).detach().requires_grad_(True)
km=keep.bool()[:,None,None,:]
def sd():
    return F.scaled_dot_product_attention(ql,kl,vl,attn_mask=km,dropout_p=0.15)
print('sdpa fwd us', t(sd))
o=sd(); g=torch.randn_like(o)
print('sdpa fwd+bwd us', t(lambda: torch.autograd.grad(sd(),(ql,kl,vl),g)))
