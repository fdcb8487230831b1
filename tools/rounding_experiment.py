import math, sys, torch
sys.path.insert(0,'/root/repo')
from oracle import denoiser as OD
from oracle.weights import mdm_state_dict, text_features
torch.set_num_threads(8)
w = mdm_state_dict(181, seed=0)
B,F,T = 4,181,196
g = torch.Generator().manual_seed(3)
x = torch.randn(B,F,1,T,generator=g)
t = torch.tensor([900, 500, 100, 3])
feat = text_features(["a"]*B)
scale = torch.full((B,),2.5)
bf = lambda v: v.bfloat16().float()

def enc_layer(x, w, pre, nh, cfg):
    S,Bn,d = x.shape; dh = d//nh
    xa = cfg.get('xfmt',bf)(x)
    qkv = xa @ cfg.get('xfmt',bf)(w[pre+"self_attn.in_proj_weight"]).T + w[pre+"self_attn.in_proj_bias"]
    qkv = bf(qkv)
    q,k,v = qkv.split(d,dim=-1)
    heads = lambda t_: t_.reshape(S,Bn,nh,dh).permute(1,2,0,3)
    q,k,v = heads(q),heads(k),heads(v)
    s = (q @ k.transpose(-1,-2))/math.sqrt(dh)
    p = torch.exp(s - s.amax(-1,keepdim=True))
    pb = bf(p)
    o = (pb @ v)/p.sum(-1,keepdim=True)
    o = bf(o.permute(2,0,1,3).reshape(S,Bn,d))
    sa = o @ bf(w[pre+"self_attn.out_proj.weight"]).T + w[pre+"self_attn.out_proj.bias"]
    res = x if cfg['res32'] else (cfg['resfn'](x))
    y = OD.layer_norm(res + sa, w[pre+"norm1.weight"], w[pre+"norm1.bias"])
    ya = cfg.get('xfmt',bf)(y)
    h = bf(OD.gelu(ya @ cfg.get('xfmt',bf)(w[pre+"linear1.weight"]).T + w[pre+"linear1.bias"]))
    ff = h @ bf(w[pre+"linear2.weight"]).T + w[pre+"linear2.bias"]
    res = y if cfg['res32'] else cfg['resfn'](y)
    return OD.layer_norm(res + ff, w[pre+"norm2.weight"], w[pre+"norm2.bias"])

def fwd(w, x, t, feat, uncond, cfg):
    B,F,_,T = x.shape
    emb = OD.time_embedding(w,t)
    ft = torch.zeros(B,512) if uncond else feat
    emb = emb + ft @ w["embed_text.weight"].T + w["embed_text.bias"]
    xs = x.permute(3,0,1,2).reshape(T,B,F)
    xs = bf(xs) @ bf(w["input_process.poseEmbedding.weight"]).T + w["input_process.poseEmbedding.bias"]
    seq = torch.cat([emb[None],xs],0) + w["sequence_pos_encoder.pe"][:T+1]
    if not cfg['res32']: seq = cfg['resfn'](seq)
    for i in range(8):
        seq = enc_layer(seq, w, f"seqTransEncoder.layers.{i}.", 4, cfg)
    out = cfg.get('xfmt',bf)(seq[1:]) @ cfg.get('xfmt',bf)(w["output_process.poseFinal.weight"]).T + w["output_process.poseFinal.bias"]
    return out.reshape(T,B,F,1).permute(1,2,3,0)

def cfgf(cfg):
    c = fwd(w,x,t,feat,False,cfg); u = fwd(w,x,t,feat,True,cfg)
    return u + scale.view(-1,1,1,1)*(c-u), c

with torch.no_grad():
    ref = OD.cfg_forward(w,x,t,feat,scale)
    refc = OD.mdm_forward(w,x,t,feat)
    def hilo(v):
        hi = bf(v); return hi + bf(v-hi)
    for name,cfg in [("bf16 residual",dict(res32=False,resfn=bf)),("fp32 residual",dict(res32=True)),("hi+lo residual",dict(res32=False,resfn=hilo))]:
        got,c = cfgf(cfg)
        e = (got-ref).abs().amax(dim=(1,2,3))/ref.abs().amax(dim=(1,2,3))
        ec = (c-refc).abs().amax(dim=(1,2,3))/refc.abs().amax(dim=(1,2,3))
        l2 = (got-ref).norm()/ref.norm()
        print(f"{name:16s} cfg max-norm relerr per sample {[f'{v:.2e}' for v in e.tolist()]}  global {float((got-ref).abs().max()/ref.abs().max()):.3e}  relL2 {float(l2):.3e} | cond-only {float((c-refc).abs().max()/refc.abs().max()):.3e}")
    f16 = lambda v: v.half().float()
    for name,cfg in [("f16 stream (res+A), bf16 rest",dict(res32=False,resfn=f16,xfmt=f16)),("f16 residual copy, bf16 A",dict(res32=False,resfn=f16))]:
        got,c = cfgf(cfg)
        e = (got-ref).abs().amax(dim=(1,2,3))/ref.abs().amax(dim=(1,2,3))
        l2 = (got-ref).norm()/ref.norm()
        print(f"{name:32s} cfg max-norm relerr per sample {[f'{v:.2e}' for v in e.tolist()]}  global {float((got-ref).abs().max()/ref.abs().max()):.3e}  relL2 {float(l2):.3e} | cond-only {float((c-refc).abs().max()/refc.abs().max()):.3e}")
