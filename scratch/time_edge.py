import sys; sys.path.insert(0,".")
import torch
from oracle import xpainn_oracle as orc
import xequinet_b200 as xb
from xequinet_b200 import ops
nm = int(sys.argv[1]) if len(sys.argv)>1 else 256
cfg=orc.CONFIG_DEFAULT; d=orc.make_aspirin_batch(nm,seed=0,with_edges=False); dev="cuda"
g,_,_=xb.build_graph(d["pos"].to(dev),5.0,ptr=d["ptr"].to(dev),batch=d["batch"].to(dev)); N=g.n_nodes
dims=ops.Dims(cfg.node_dim,*cfg.muls,cfg.num_basis,cfg.cutoff); r=lambda *s: torch.randn(*s,device=dev)
pos=d["pos"].to(dev); s,v,x,V=r(N,dims.H),r(N,dims.D),r(N,dims.node_dim),r(N,dims.D); W,b=0.3*r(dims.H,20),0.3*r(dims.H); freq=(torch.pi*torch.arange(1,21,device=dev)/5.0).float()
gx,gV,a_s,a_v,a_p=r(N,dims.node_dim),r(N,dims.D),r(N,dims.H),r(N,dims.D),r(N,3)
def tm(f,n=20):
    f(); torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b_=torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(n): f()
    b_.record(); torch.cuda.synchronize(); return a.elapsed_time(b_)/n
print("N",N,"tile_mode",g.tile_mode,"E",g.n_edges)
print("fwd ms %.4f"%tm(lambda: ops.edge_message_fwd_raw(g,dims,pos,s,v,x,V,W,b,freq)))
print("bwd ms %.4f"%tm(lambda: ops.edge_message_bwd_raw(g,dims,pos,s,v,W,b,freq,gx,gV,need_w=False)))
print("bwd+w ms %.4f"%tm(lambda: ops.edge_message_bwd_raw(g,dims,pos,s,v,W,b,freq,gx,gV,need_w=True)))
print("bwdbwd ms %.4f"%tm(lambda: ops.edge_message_bwdbwd_raw(g,dims,pos,s,v,W,b,freq,gx,gV,a_s,a_v,a_p)))
print("jvp-only ms %.4f"%tm(lambda: ops.edge_message_bwdbwd_raw(g,dims,pos,s,v,W,b,freq,gx,gV,a_s,a_v,a_p,need_s=False,need_v=False,need_pos=False,need_w=False)))
print("bwdbwd-main ms %.4f"%tm(lambda: ops.edge_message_bwdbwd_raw(g,dims,pos,s,v,W,b,freq,gx,gV,a_s,a_v,a_p,need_g=False,need_w=False)))
