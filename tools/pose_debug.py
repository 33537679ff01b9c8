import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmdyn_b200 import engine
from mmdyn_b200.pytorch.models.models import setup_model
DEV = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(5)
model = setup_model("cnn-mvae", cross_modal=True, condition_dim=0, input_dim=4096, architecture="cnn", conditional=False,
                    categorical_conditions=False, latent_size=256, use_pose=True).to(DEV)
arena, ex = engine.get_execs(model, torch.device(DEV))
pex = ex["pose"]
ws = engine.Workspace(torch.device(DEV))
arena.attach_grads(); arena.grad.zero_()
g = torch.Generator().manual_seed(B)
pose, z = torch.rand(B, 7, generator=g), torch.randn(B, 256, generator=g)
d_heads = torch.randn(B, 512, generator=g)
r = pex.enc_forward(pose.to(DEV), ws, "penc")
def snap(tag):
    torch.cuda.synchronize()
    print(tag, "fc0.w grad norm", pex.p("pose_encoder.fc_net.0.weight", True).norm().item(), "fc0.b", pex.p("pose_encoder.fc_net.0.bias", True).norm().item(), flush=True)
snap("after fwd")
pex.enc_backward(r, d_heads.to(DEV), ws, "penc", 0.5)
snap("after enc bwd")
dh1, h1 = ws.bufs["penc.dh1"], r["h1"]
dya = dh1 * (h1 > 0)
want_w = 0.5 * dya.t() @ pose.to(DEV)
want_b = 0.5 * dya.sum(0)
gw = pex.p("pose_encoder.fc_net.0.weight", True).view(512, 7)
gb = pex.p("pose_encoder.fc_net.0.bias", True)
print("dW0 vs torch on the same GPU buffers:", ((gw - want_w).norm() / want_w.norm()).item(), "db0:", ((gb - want_b).norm() / want_b.norm()).item())
print("scr vs dya:", ((ws.bufs["penc.scr"] - dya).norm() / dya.norm()).item())
d = (gw - want_w).abs()
print("worst rows:", d.sum(1).topk(5))
before = gw.clone()
rd = pex.dec_forward(z.to(DEV), ws, "pdec")
snap("after dec fwd")
d_rec = 100.0 * torch.randn(B, 7, generator=g)
pex.dec_backward(rd, d_rec.to(DEV), ws, "pdec", 0.5)
snap("after dec bwd")
print("fc0.w changed by the decoder passes:", (gw - before).abs().max().item())
# fp64 reference of the first layer only
import torch.nn.functional as F
sd = {k: v.detach().double().cpu() for k, v in model.state_dict().items() if k.startswith("pose_")}
h1r = F.relu(F.linear(pose.double(), sd["pose_encoder.fc_net.0.weight"], sd["pose_encoder.fc_net.0.bias"]))
flips = ((h1r > 0) != (h1.cpu() > 0)).sum().item()
print("ReLU mask flips fp32 GPU vs fp64 reference:", flips, "of", h1r.numel(), " h1 rel err", ((h1.cpu().double() - h1r).norm() / h1r.norm()).item())
idx = ((h1r > 0) != (h1.cpu() > 0)).nonzero()
for r_, u_ in idx.tolist():
    pre = (pose[r_].double() @ sd["pose_encoder.fc_net.0.weight"][u_] + sd["pose_encoder.fc_net.0.bias"][u_]).item()
    print("flip at row", r_, "unit", u_, "fp64 pre-activation", pre, "gpu h1", h1[r_, u_].item(), "dh1", dh1[r_, u_].item(),
          "row-term norm", (0.5 * dh1[r_, u_].item() * pose[r_]).norm().item(), "dW0 norm", gw.norm().item())
