"""Which part of StyleTrainStep invalidates a CUDA graph capture?  (diagnostic, GPU)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from types import SimpleNamespace
import torch
from laenerf_b200.style_encoder import LAENeRF, StyleTrainStep
dev = torch.device("cuda")
params = SimpleNamespace(bound=2.0, num_palette_bases=8, style_weight=0.0, weight_loss_uniform=1e-6, weight_loss_non_uniform=1e-6,
                         offset_loss=1e-6, palette_loss_valid=1e-3, palette_loss_distinct=1e-3)
K = 8192
x = (torch.rand(K, 3, device=dev) - 0.5) * 3
d = torch.nn.functional.normalize(torch.randn(K, 3, device=dev), dim=-1)
t = torch.rand(K, 3, device=dev)
style = LAENeRF(params, dir_encoding="sphere_harmonics").to(dev)
st = StyleTrainStep(style, params)
for _ in range(3):
    st(x, d, t)
torch.cuda.synchronize()
m, p = style, params

def fwd():
    return m.forward_train(x=x, d=d)
def fwd_loss():
    pc, pw, po = m.forward_train(x=x, d=d)
    loss = st.loss_fct(input=pc, target=t.half())
    return loss + m.weights_loss(pw, p).half() + m.offset_loss(po, p).half() + m.palet_loss(p).half()
def fwd_bwd():
    st.optimizer.zero_grad()
    st.optimizer.scale(fwd_loss()).backward()
def full():
    st(x, d, t)
def enc_only():
    return m.encoder(x, bound=m.bound)
def enc_bwd():
    e = m.encoder(x, bound=m.bound)
    e.float().sum().backward()
def opt_only():
    fwd_bwd()
    st.optimizer.step()

for name, fn in (("encoder fwd", enc_only), ("encoder fwd+bwd", enc_bwd), ("forward_train", fwd), ("forward + losses", fwd_loss), ("forward + backward", fwd_bwd),
                 ("optimizer.step", opt_only), ("full step", full)):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    try:
        with torch.cuda.stream(side):
            fn()
    except Exception as e:
        print("EAGER FAIL", name, str(e)[:100]); continue
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("OK  ", name)
    except Exception as e:
        print("FAIL", name, "--", str(e).splitlines()[0][:150])
        torch.cuda.synchronize()
