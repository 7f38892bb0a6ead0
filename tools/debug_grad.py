"""Where does the whole-network gradient first deviate from the reference's fp32 autograd?  (test/debug tool, not product)"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from argparse import Namespace
import ref_harness as rh
from airpose_b200 import synthetic
from airpose_b200.copenet_twoview import copenet_twoview
from test_gpu_dropin import _install_bf16_rounding_points

B = 4
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
sd = synthetic.make_network_state(123, dec_gain=0.01)
tsd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
batch = rh.make_batch(B, 123, 321, device="cuda")
for k in list(tsd):
    if tsd[k].dim() == 4: tsd[k] = tsd[k].to(torch.bfloat16).float()
for k in ("im0", "im1"): batch[k] = batch[k].to(torch.bfloat16).float()
rt = rh.import_reference("cuda")
ref = rh.make_module(rt, B, device="cuda", load_weights=False)
ref.model.load_state_dict(tsd, strict=True); ref.train()
_install_bf16_rounding_points(ref.model)
for m in ref.model.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
feats = []
orig = ref.model.forward_feat_ext
def wrapped(x):
    out = orig(x); out.retain_grad(); feats.append(out); return out
ref.model.forward_feat_ext = wrapped
acts = {}
def keep(name):
    def hook(m, a, out):
        out.retain_grad(); acts.setdefault(name, []).append(out)
    return hook
ref.model.layer4[2].register_forward_hook(keep("layer4.2"))
ref.model.layer4[2].bn3.register_forward_hook(keep("layer4.2.bn3"))
ref.model.layer4[2].conv3.register_forward_hook(keep("layer4.2.conv3"))
x2s = []
ref.model.layer4[2].conv3.register_forward_pre_hook(lambda m, a: x2s.append(a[0].detach().to(torch.bfloat16).float()) or None)
res = ref.training_step({k: v.clone() for k, v in batch.items()}, 1)
res["loss"].backward()
g_ref_xf = torch.cat([f.grad for f in feats])
tmp = tempfile.mkdtemp()
mp = synthetic.write_mean_params(os.path.join(tmp, "m.npz")); synthetic.write_smplx_model(tmp, 0)
mod = copenet_twoview(Namespace(smpl_mean_params=mp, smplx_model_dir=tmp, batch_size=B, val_batch_size=B, reg_iters=3, lr=5e-5))
mod.model.load_state_dict(tsd, strict=True); mod = mod.to("cuda").train()
opt = mod.configure_optimizers()
cap = {}
ob = mod.model.backward_feat_ext
def capture(x, tape, g, **kw):
    cap["g"] = g.clone(); return ob(x, tape, g, **kw)
mod.model.backward_feat_ext = capture
of = mod.model._forward_feat_ext_train_pair
def capf(a, b, tape=0):
    out = of(a, b, tape=tape); cap["xf"] = out.clone(); return out
mod.model._forward_feat_ext_train_pair = capf
loss, _ = mod.training_step(batch, opt, mask1=False, mask2=False)
torch.cuda.synchronize()
cos = lambda a, b: float((a.double().flatten() @ b.double().flatten()) / (a.double().norm() * b.double().norm()))
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
xf_ref = torch.cat([f.detach() for f in feats])
print("features: cos %.6f rel %.3e" % (cos(cap["xf"], xf_ref), rel(cap["xf"], xf_ref)))
print("d loss / d features: cos %.6f rel %.3e  |ref| %.3e" % (cos(cap["g"], g_ref_xf), rel(cap["g"], g_ref_xf), float(g_ref_xf.abs().max())))
# gradient at the last block's output / bn3 output / conv3 output in the reference (both views concatenated)
for name in ("layer4.2", "layer4.2.bn3", "layer4.2.conv3"):
    g = torch.cat([t.grad for t in acts[name]])
    print(name, "ref grad shape", tuple(g.shape), "absmax %.3e" % float(g.abs().max()), "fraction nonzero %.3f" % float((g != 0).float().mean()))
lib_names = [n for n, _ in mod.model.named_parameters()]
for n, p in mod.model.named_parameters():
    if n.startswith("layer4.2") or n.startswith("layer4.1.conv3") or n.startswith("layer4.1.bn3"):
        gr = dict(ref.model.named_parameters())[n].grad
        print("%-28s cos %.5f rel %.3e" % (n, cos(p.grad, gr), rel(p.grad, gr)))
# tape access: our dz of layer4.2.conv3 (conv index 52) against the reference's gradient at conv3's output
import ctypes as C
from airpose_b200 import _lib
lib = _lib.load()


# ---- the first weight gradient of the backward pass (layer4.2.conv3, 1x1) recomputed from the REFERENCE's own tensors
gref = ref.model.layer4[2].conv3.weight.grad[:, :, 0, 0]
rb = lambda t: t.to(torch.bfloat16).float()
Wa = sum(torch.einsum("nohw,nihw->oi", acts["layer4.2.conv3"][v].grad, x2s[v]) for v in range(2))
Wb = sum(torch.einsum("nohw,nihw->oi", rb(acts["layer4.2.conv3"][v].grad), x2s[v]) for v in range(2))
ours = dict(mod.model.named_parameters())["layer4.2.conv3.weight"].grad[:, :, 0, 0]
print("wgrad from reference dz (fp32) and its bf16 input: cos vs reference grad %.6f" % cos(Wa, gref))
print("wgrad from reference dz ROUNDED to bf16:            cos vs reference grad %.6f" % cos(Wb, gref))
print("ours vs reference grad %.6f ; ours vs rounded-dz recomputation %.6f" % (cos(ours, gref), cos(ours, Wb)))
dzr = torch.cat([acts["layer4.2.conv3"][v].grad for v in range(2)])
print("reference dz: absmax %.3e, mean |dz| %.3e, per-channel |sum| / sum|.| median %.3e" % (float(dzr.abs().max()), float(dzr.abs().mean()),
      float((dzr.sum(dim=(0, 2, 3)).abs() / dzr.abs().sum(dim=(0, 2, 3))).median())))

# ---- our tape against the reference's activations at the last block
import ctypes as C
from airpose_b200 import _lib
lib = _lib.load()
hdl = mod.model._handle
def tape(i, which, C_, H):
    t_ = torch.empty(2 * B, H, H, C_, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.airpose_debug_tape_get(hdl, 0, i, which, t_.data_ptr(), t_.numel(), _lib.current_stream()), "tape_get")
    torch.cuda.synchronize()
    return t_.float().permute(0, 3, 1, 2)
z3_o, y3_o, y2_o = tape(52, 0, 2048, 7), tape(52, 1, 2048, 7), tape(51, 1, 512, 7)
z3_r = torch.cat([acts["layer4.2.conv3"][v].detach() for v in range(2)])
y3_r = torch.cat([acts["layer4.2"][v].detach() for v in range(2)])
x2_r = torch.cat(x2s)
print("tape vs reference at layer4.2: z3 cos %.6f rel %.3e | block output cos %.6f rel %.3e | conv3 input cos %.6f rel %.3e" %
      (cos(z3_o, z3_r), rel(z3_o, z3_r), cos(y3_o, y3_r), rel(y3_o, y3_r), cos(y2_o, x2_r), rel(y2_o, x2_r)))
mask_o, mask_r = (y3_o > 0), (y3_r > 0)
print("ReLU mask of the block output: %.4f of the elements differ" % float((mask_o != mask_r).float().mean()))
# wgrad recomputed in torch from OUR tape (per view BatchNorm backward with the reference's upstream gradient)
def bn_bwd(dpre, z, gamma, eps=1e-5):
    M = z.numel() / z.shape[1]
    mean = z.mean(dim=(0, 2, 3), keepdim=True); var = z.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    invstd = 1.0 / torch.sqrt(var + eps); xh = (z - mean) * invstd
    db = dpre.sum(dim=(0, 2, 3), keepdim=True); dg = (dpre * xh).sum(dim=(0, 2, 3), keepdim=True)
    return gamma.view(1, -1, 1, 1) * invstd * (dpre - db / M - xh * dg / M)
gam = ref.model.layer4[2].bn3.weight.detach()
G_r = torch.cat([acts["layer4.2"][v].grad for v in range(2)])          # gradient at the block output (after ReLU)
Wt = 0
for v in range(2):
    sl = slice(v * B, (v + 1) * B)
    dpre = G_r[sl] * mask_o[sl]
    dz = bn_bwd(dpre, z3_o[sl], gam)
    Wt = Wt + torch.einsum("nohw,nihw->oi", dz, y2_o[sl])
print("wgrad recomputed in torch from OUR tape + reference upstream gradient: cos vs reference %.6f, vs ours %.6f" % (cos(Wt, gref), cos(Wt, ours)))
